import os

import torch

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load_fixture(name):
    return torch.load(os.path.join(GOLDEN, name), map_location="cpu", weights_only=False)


def build_from_meta(meta, dtype=torch.float32, device="cuda"):
    """Builds the mantis_b200 model described by a golden fixture's meta."""
    from transformers import CLIPVisionConfig, LlamaConfig, SiglipVisionConfig

    from mantis_b200.models.mllava import LlavaConfig, LlavaForConditionalGeneration, MLlavaForConditionalGeneration
    vc = (SiglipVisionConfig if meta["vision_kind"] == "siglip" else CLIPVisionConfig)(**meta["vision_kwargs"])
    tc = LlamaConfig(**meta["text_kwargs"])
    cfg = LlavaConfig(vision_config=vc, text_config=tc, **meta["cfg_kwargs"])
    cls = MLlavaForConditionalGeneration if meta["cls"] == "mllava" else LlavaForConditionalGeneration
    model = cls(cfg)
    return model


def load_model(fx, dtype=torch.float32, device="cuda"):
    model = build_from_meta(fx["meta"])
    missing, unexpected = model.load_state_dict(fx["state_dict"], strict=False)
    assert not unexpected, f"unexpected keys: {unexpected[:5]}"
    assert not missing, f"missing keys: {missing[:5]}"
    return model.to(device=device, dtype=dtype)


def rel_err(a, b):
    a = a.detach().float().cpu(); b = b.detach().float().cpu()
    return ((a - b).norm() / (b.norm() + 1e-12)).item()
