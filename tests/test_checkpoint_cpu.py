"""CPU tests of the HF checkpoint surface: save_pretrained -> from_pretrained round-trips bit-exactly with the REFERENCE key
layout (transformers >= 5 would otherwise rename `language_model.model.*` for model_type "llava" and re-initialise loaded
weights through an unguarded _init_weights)."""
import os
import tempfile

import torch
from safetensors.torch import load_file

from helpers import build_from_meta, load_fixture


def test_llava_save_load_roundtrip_keeps_reference_keys():
    from mantis_b200.models.mllava import LlavaForConditionalGeneration
    fx = load_fixture("llava_siglip_full.pt")
    m = build_from_meta(fx["meta"]); m.load_state_dict(fx["state_dict"])
    d = tempfile.mkdtemp()
    m.save_pretrained(d)
    on_disk = load_file(os.path.join(d, "model.safetensors"))
    assert set(on_disk.keys()) == set(fx["state_dict"].keys())            # == the reference's state-dict layout
    m2 = LlavaForConditionalGeneration.from_pretrained(d)
    for k, v in fx["state_dict"].items():
        assert torch.equal(m2.state_dict()[k], v), k
    assert m2.config.image_token_index == 300 and type(m2.config.text_config).__name__ == "LlamaConfig"
    # the registry of transformers' own renamings is restored afterwards
    from transformers import conversion_mapping as cm
    assert cm.get_checkpoint_conversion_mapping("llava")


def test_idefics2_save_load_roundtrip():
    from transformers import Idefics2Config
    from mantis_b200.models.idefics2 import Idefics2ForConditionalGeneration
    fx = load_fixture("idefics2_full.pt")
    m = Idefics2ForConditionalGeneration(Idefics2Config(**fx["cfg"])); m.load_state_dict(fx["state_dict"])
    d = tempfile.mkdtemp()
    m.save_pretrained(d)
    assert set(load_file(os.path.join(d, "model.safetensors")).keys()) == set(fx["state_dict"].keys())
    m2 = Idefics2ForConditionalGeneration.from_pretrained(d)
    for k, v in fx["state_dict"].items():
        assert torch.equal(m2.state_dict()[k], v), k


def test_idefics3_save_load_roundtrip():
    from mantis_b200.models.idefics3 import Idefics3Config, Idefics3ForConditionalGeneration
    fx = load_fixture("idefics3_full.pt")
    m = Idefics3ForConditionalGeneration(Idefics3Config(**fx["cfg"])); m.load_state_dict(fx["state_dict"])
    d = tempfile.mkdtemp()
    m.save_pretrained(d)
    assert set(load_file(os.path.join(d, "model.safetensors")).keys()) == set(fx["state_dict"].keys())
    m2 = Idefics3ForConditionalGeneration.from_pretrained(d)
    for k, v in fx["state_dict"].items():
        assert torch.equal(m2.state_dict()[k], v), k
    assert m2.config.scale_factor == 2 and type(m2.config.text_config).__name__ == "LlamaConfig"


def test_llava_next_save_load_roundtrip():
    from transformers import CLIPVisionConfig, LlamaConfig
    from mantis_b200.models.mllava_next import LlavaNextConfig, LlavaNextForConditionalGeneration
    fx = load_fixture("llava_next_batch.pt")
    cfg = LlavaNextConfig(vision_config=CLIPVisionConfig(**fx["vision"]), text_config=LlamaConfig(**fx["text"]), **fx["cfg"])
    m = LlavaNextForConditionalGeneration(cfg); m.load_state_dict(fx["state_dict"])
    d = tempfile.mkdtemp()
    m.save_pretrained(d)
    assert set(load_file(os.path.join(d, "model.safetensors")).keys()) == set(fx["state_dict"].keys())
    m2 = LlavaNextForConditionalGeneration.from_pretrained(d)
    for k, v in fx["state_dict"].items():
        assert torch.equal(m2.state_dict()[k], v), k
    assert torch.equal(m2.image_newline, fx["state_dict"]["image_newline"])
