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


def merge_properties(final, mask, pos, inputs_embeds, image_features, input_ids, image_token):
    """Size-independent properties of the image-token merge for UNPADDED batches whose samples all carry the same number of
    images (the bench layout, SURVEY 8a: S = T + n_img * (P - 1)); operands hold small integers so every sum below is exact
    in fp32.  Works on any device (the CPU test feeds it the numpy oracle's output, the GPU test the kernels' at full size).
      1. conservation: column-wise checksum of the output == checksum of the text rows that are not placeholders + checksum
         of all image rows (every source row lands exactly once, nothing else is written)
      2. order: walking a sample left to right, text rows keep their order and image k's P rows are contiguous and ascending
      3. mask all ones, position ids 0..S-1"""
    import torch
    B, T = input_ids.shape
    n_img_total, P, D = image_features.shape
    S = final.shape[1]
    is_img = input_ids == image_token
    assert S == T + int(is_img[0].sum()) * (P - 1)
    text_sum = (inputs_embeds.float() * (~is_img)[..., None]).sum(dim=(0, 1))
    assert torch.equal(final.float().sum(dim=(0, 1)), text_sum + image_features.float().sum(dim=(0, 1)))
    assert bool((mask != 0).all())
    assert torch.equal(pos, torch.arange(S, device=pos.device).expand(B, S))
    # 2. order: rebuild the layout from the reference's position formula (new position of token t = cumsum(P if image else 1)
    #    - 1, modeling_llava.py:309); the slots in between take the image rows in row-major order
    step = torch.where(is_img, torch.full_like(input_ids, P), torch.ones_like(input_ids))
    new_pos = torch.cumsum(step, dim=1) - 1                                   # position of a text token / LAST row of an image
    expected = torch.zeros_like(final)
    taken = torch.zeros((B, S), dtype=torch.bool, device=final.device)
    bi, ti = (~is_img).nonzero(as_tuple=True)
    expected[bi, new_pos[bi, ti]] = inputs_embeds[bi, ti].to(final.dtype)
    taken[bi, new_pos[bi, ti]] = True
    expected[~taken] = image_features.reshape(-1, D).to(final.dtype)
    assert torch.equal(final, expected)
