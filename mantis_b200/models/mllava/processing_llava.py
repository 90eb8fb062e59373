"""MLlavaProcessor: interleaved text + multi-image preprocessing (caller of the hot path).

Same behaviour as mantis/models/mllava/processing_llava.py:44-285: balances `<image>` placeholders against the number of
images (prepend missing / drop surplus), rewrites each placeholder to "(image {j}: <Image><image></Image>)", tokenizes,
drops images whose placeholders were truncated away, runs the image processor, returns input_ids / attention_mask /
pixel_values.  `_right_pad_inputs_with_attention_mask` returns the reference's result for one sample (pixel_values stays a
list) and, unlike the reference (which asserts), right-pads real batches.
"""
from typing import Dict, List

import torch
from transformers.feature_extraction_utils import BatchFeature


class MLlavaProcessor:
    attributes = ["image_processor", "tokenizer"]

    def __init__(self, image_processor=None, tokenizer=None):
        self.image_processor = image_processor
        self.tokenizer = tokenizer
        self.image_token_index = None

    @classmethod
    def from_pretrained(cls, pretrained_model_name_or_path, **kwargs):
        from transformers import AutoImageProcessor, AutoTokenizer
        tok = AutoTokenizer.from_pretrained(pretrained_model_name_or_path, **kwargs)
        ip = AutoImageProcessor.from_pretrained(pretrained_model_name_or_path, **kwargs)
        return cls(image_processor=ip, tokenizer=tok)

    def save_pretrained(self, save_directory, **kwargs):
        self.tokenizer.save_pretrained(save_directory, **kwargs)
        self.image_processor.save_pretrained(save_directory, **kwargs)

    @staticmethod
    def _balance(text: str, num_images: int) -> str:
        n_tok = text.count("<image>")
        if n_tok < num_images:
            missing = "<image>" * (num_images - n_tok)
            for tag in ("USER:", "Human:", "HUMAN:"):
                if tag in text:
                    return text.replace(tag, tag + missing, 1)
            return missing + text
        if n_tok > num_images:
            parts = text.split("<image>")
            return "".join(p + ("<image>" if i < num_images else "") for i, p in enumerate(parts))
        return text

    def preprocess_interleaved_images_and_text(self, text, images=None):
        assert text is not None, "text cannot be None."
        if images is None:
            if isinstance(text, str):
                return [text], None
            if isinstance(text, list) and (not text or isinstance(text[0], str)):
                return text, None
            raise ValueError("Invalid input text. text must be a string or a list of strings.")
        is_img = lambda x: hasattr(x, "size") and not isinstance(x, (list, tuple))  # PIL-like
        if is_img(images):
            images = [images]
        if isinstance(images, list) and images and is_img(images[0]):
            if isinstance(text, str):
                images = [images]
            elif isinstance(text, list):
                if len(text) != len(images):
                    raise ValueError("Invalid input text. Number of texts does not match number of images.")
                images = [[im] for im in images]
        if isinstance(text, str):
            texts = [self._balance(text, len(images[0]))]
        elif isinstance(text, list):
            if not isinstance(text[0], str):
                raise ValueError("Invalid input text. Each element of text must be a string.")
            texts = [self._balance(t, len(images[i])) for i, t in enumerate(text)]
        else:
            raise ValueError("Invalid input text. text must be a string or a list of strings.")
        assert all(t.count("<image>") == len(ims) for t, ims in zip(texts, images)), \
            "Number of <image> tokens in text does not match number of images."
        out = []
        for i, t in enumerate(texts):
            for j in range(len(images[i])):
                t = t.replace("<image>", f"(image {j + 1}: <Image><IMAGE></Image>)", 1)
            out.append(t.replace("<IMAGE>", "<image>"))
        return out, images

    def __call__(self, text=None, images=None, padding=False, truncation=None, max_length=None, return_tensors="pt",
                 add_image_ids: bool = True) -> BatchFeature:
        if not self.image_token_index:
            self.image_token_index = self.tokenizer.convert_tokens_to_ids("<image>")
        if add_image_ids:
            text, images = self.preprocess_interleaved_images_and_text(text, images)
        text_inputs = self.tokenizer(text, return_tensors=return_tensors, padding=padding, truncation=truncation,
                                     max_length=max_length)
        pixel_values = None
        if images is not None:
            n_tok = torch.sum(torch.as_tensor(text_inputs["input_ids"]) == self.image_token_index, dim=-1)
            for i, n in enumerate(n_tok):
                if n < len(images[i]):
                    images[i] = images[i][:int(n)]
            flat = [im for per_text in images for im in per_text]
            pixel_values = self.image_processor(flat, return_tensors=return_tensors)["pixel_values"]
        return BatchFeature(data={**text_inputs, "pixel_values": pixel_values})

    def batch_decode(self, *args, **kwargs):
        return self.tokenizer.batch_decode(*args, **kwargs)

    def decode(self, *args, **kwargs):
        return self.tokenizer.decode(*args, **kwargs)

    @property
    def model_input_names(self):
        return list(dict.fromkeys(list(self.tokenizer.model_input_names) + list(self.image_processor.model_input_names)))

    def _right_pad_inputs_with_attention_mask(self, model_inputs: List[Dict]):
        """Batch a list of per-sample processor outputs.  The reference only supports one sample here (it asserts, :279); real
        batches are right-padded like its training collator does (mantis/train/data.py:1392-1527): input_ids with the pad id,
        attention_mask with 0, labels with -100, `pixel_values` stays a list of per-sample tensors (the model concatenates
        it, modeling_llava.py:431-432)."""
        if len(model_inputs) > 1:
            from ...train.data import Collator
            return Collator(processor=None, pad_token_id=getattr(self.tokenizer, "pad_token_id", None))(model_inputs)
        sample = model_inputs[0]
        return {k: ([v] if k == "pixel_values" else v) for k, v in sample.items()}
