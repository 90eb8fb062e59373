"""ChatDataset / assistant_labels (the supervision mask of instruction tuning) against the UNMODIFIED reference
ChatDataset.getitem (mantis/train/data.py:346-491), run on a bare instance with the same stub processor."""
import copy
import random

import pytest
import torch

from mantis_b200.models.conversation import conv_templates
from mantis_b200.train import ChatDataset, assistant_labels


class _Tok:
    """whitespace tokenizer; every template separator is one token"""
    pad_token_id = 0
    specials = {"</s>": 2, "<|eot_id|>": 3, "<end_of_utterance>": 4, "<image>": 7}

    def __init__(self):
        self.vocab = dict(self.specials)

    def convert_tokens_to_ids(self, t):
        return self.vocab.setdefault(t, 10 + len(self.vocab))

    def encode(self, text):
        for s in self.specials:
            text = text.replace(s, f" {s} ")
        return [self.convert_tokens_to_ids(w) for w in text.split()]


class _Proc:
    def __init__(self):
        self.tokenizer = _Tok()

    def __call__(self, text, images=None, return_tensors="pt", truncation=None, max_length=None, **kw):
        ids = self.tokenizer.encode(text)
        if truncation and max_length:
            ids = ids[:max_length]
        t = torch.tensor([ids])
        return {"input_ids": t, "attention_mask": torch.ones_like(t), "pixel_values": None if not images else torch.zeros(len(images), 3, 2, 2)}


def test_assistant_labels_small_example():
    # system </s> user q </s> assistant a1 a2 </s> user q2 </s> assistant a3
    ids = torch.tensor([[11, 2, 12, 13, 2, 14, 15, 16, 2, 12, 17, 2, 14, 18]])
    lab = assistant_labels(ids, 2, "single")
    assert lab.tolist() == [[-100, -100, -100, -100, -100, 14, 15, 16, 2, -100, -100, -100, 14, 18]]
    # idefics without system: separator 0 already closes the user turn
    ids = torch.tensor([[12, 13, 4, 14, 15, 4, 12, 16, 4, 14, 17, 4]])
    assert assistant_labels(ids, 4, "idefics", has_system=False).tolist() == [[-100, -100, -100, 14, 15, 4, -100, -100, -100, 14, 17, 4]]
    with pytest.raises(ValueError):
        assistant_labels(ids, 4, "plain")


def test_chat_dataset_matches_the_live_reference():
    from oracle.ref_shim import find_ref_root
    if find_ref_root() is None:
        pytest.skip("reference tree not available here")
    from oracle.ref_shim import load_reference_train_data
    import sys
    from PIL import Image
    ref = load_reference_train_data()
    ref_conv = sys.modules["mantis.models.conversation"]
    rnd = random.Random(21)
    words = ["alpha", "beta", "<image>", "gamma", "delta", "what", "is", "this", "<image>"]
    n = 0
    for name in ("mllava_v1", "llama_3", "idefics_2", "idefics_3"):
        for trial in range(40):
            turns = []
            for t in range(rnd.randint(1, 4)):
                turns.append({"role": "user", "content": " ".join(rnd.choice(words) for _ in range(rnd.randint(1, 6)))})
                turns.append({"role": "assistant", "content": " ".join(rnd.choice(words[:2] + words[3:8]) for _ in range(rnd.randint(1, 5)))})
            if rnd.random() < 0.2:
                turns = [{"role": "assistant", "content": "dropped"}] + turns
            n_img = sum(t["content"].count("<image>") for t in turns) + rnd.randint(0, 2)
            images = [Image.new("RGB", (20, 20)) for _ in range(n_img)]
            max_len = rnd.choice([None, 12, 30])
            record = {"conversation": turns, "images": images}
            ours = ChatDataset([dict(conversation=copy.deepcopy(turns), images=list(images))], _Proc(), conv_templates[name], max_seq_len=max_len)[0]
            # the reference method on a bare instance (its __init__ loads HF datasets)
            r = object.__new__(ref.ChatDataset)
            r.conv = ref_conv.conv_templates[name].copy()
            r.processor, r.max_seq_len, r.ensure_seq_len_multiple_of = _Proc(), max_len, None
            r.max_image_size, r.image_dir, r.image_key = None, None, "images"
            r.data, r.conversations, r.all_images = [dict(conversation=copy.deepcopy(turns), images=list(images))], [None], [None]
            r.data_path = r.name = r.split = "probe"
            exp = r.getitem(0)
            assert torch.equal(ours["input_ids"], exp["input_ids"]), (name, turns)
            assert torch.equal(ours["labels"], exp["labels"]), (name, turns)
            n += 1
    assert n == 160
