"""chat_mllava: single-call multi-image chat on top of generate() (mirror of mantis/models/mllava/utils.py:10-97)."""
from typing import List, Tuple, Union

import torch

from ..conversation import conv_mllava_v1 as default_conv, conv_templates


def chat_mllava(text: str, images: List[Union["PIL.Image.Image", str]], model, processor, max_input_length: int = None,
                history: List[dict] = None, **kwargs) -> Tuple[str, List[dict]]:
    name = getattr(model.language_model, "name_or_path", "") or ""
    if "llama-3" in name.lower():
        conv = conv_templates["llama_3"]
        terminators = [processor.tokenizer.eos_token_id, processor.tokenizer.convert_tokens_to_ids("<|eot_id|>")]
    else:
        conv = default_conv
        terminators = None
    kwargs["eos_token_id"] = terminators
    conv = conv.copy()
    conv.messages = []
    if history is not None:
        for message in history:
            assert message["role"] in conv.roles
            conv.append_message(message["role"], message["text"])
        if text:
            assert conv.messages[-1][0] == conv.roles[1], \
                "The last message in the history should be the assistant, if the given text is not empty"
            conv.append_message(conv.roles[0], text)
            conv.append_message(conv.roles[1], "")
            history.append({"role": conv.roles[0], "text": text})
            history.append({"role": conv.roles[1], "text": ""})
        else:
            if conv.messages[-1][0] == conv.roles[1]:
                assert conv.messages[-1][1] == "", "No user message should be provided"
            else:
                assert conv.messages[-1][0] == conv.roles[0], \
                    "The last message in the history should be the user, if the given text is empty"
                conv.append_message(conv.roles[0], "")
                history.append({"role": conv.roles[0], "text": ""})
    else:
        history = [{"role": conv.roles[0], "text": text}, {"role": conv.roles[1], "text": ""}]
        conv.append_message(conv.roles[0], text)
        conv.append_message(conv.roles[1], "")
    assert conv.messages[-1][0] == conv.roles[1] and conv.messages[-1][1] == "", "Format check"
    prompt = conv.get_prompt()
    if images:
        import PIL.Image
        for i in range(len(images)):
            if isinstance(images[i], str):
                images[i] = PIL.Image.open(images[i]).convert("RGB")
    inputs = processor(images=images, text=prompt, return_tensors="pt", truncation=True, max_length=max_input_length)
    for k, v in list(inputs.items()):
        if v is None:
            continue
        if isinstance(v, torch.Tensor):
            inputs[k] = v.to(model.device)
        elif isinstance(v, list):
            inputs[k] = [x.to(model.device) for x in v]
        else:
            raise ValueError(f"Invalid input type: {type(v)}")
    output_ids = model.generate(**inputs, **kwargs)[0]
    generated_ids = output_ids[inputs["input_ids"].shape[-1]:]
    generated_text = processor.decode(generated_ids, skip_special_tokens=True)
    history[-1]["text"] = generated_text
    return generated_text, history
