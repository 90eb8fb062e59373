"""TEST INFRASTRUCTURE ONLY.  Golden fixtures for Idefics2ForConditionalGeneration from the UNMODIFIED reference
(mantis/models/idefics2/modeling_idefics2.py via oracle/ref_shim.py), CPU fp32, eager attention.

  idefics2_full.pt    1 sample, 2 full-resolution images (pixel_attention_mask = None), 8 latents/image
  idefics2_ragged.pt  2 samples x 2 image slots: one smaller image (partial pixel mask -> NaViT position ids + patch padding
                      mask) and one all-zero padding image that the model must drop
Also pins navit position-id KAT-6/7/8 of SURVEY section 8c in tests/test_oracle.py through the product's host function.
"""
import os
import sys

import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from oracle.ref_shim import ref_idefics2_classes  # noqa: E402

from transformers import Idefics2Config  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden")
CFG = dict(
    vision_config=dict(hidden_size=64, intermediate_size=128, num_hidden_layers=2, num_attention_heads=4, image_size=112,
                       patch_size=14),
    perceiver_config=dict(resampler_n_latents=8, resampler_depth=2, resampler_n_heads=4, resampler_head_dim=16,
                          num_key_value_heads=2, hidden_size=64, rms_norm_eps=1e-5),
    text_config=dict(model_type="mistral", hidden_size=64, intermediate_size=128, num_hidden_layers=2,
                     num_attention_heads=4, num_key_value_heads=2, vocab_size=320, pad_token_id=0, rms_norm_eps=1e-5,
                     sliding_window=4096),
    image_token_id=300, tie_word_embeddings=False)
GRAD_KEYS = ["lm_head.weight", "model.text_model.embed_tokens.weight", "model.text_model.layers.0.self_attn.q_proj.weight",
             "model.connector.perceiver_resampler.latents", "model.connector.perceiver_resampler.layers.0.self_attn.k_proj.weight",
             "model.connector.perceiver_resampler.layers.1.mlp.down_proj.weight", "model.connector.perceiver_resampler.norm.weight",
             "model.connector.perceiver_resampler.layers.0.input_context_norm.weight",
             "model.connector.modality_projection.gate_proj.weight", "model.vision_model.post_layernorm.weight",
             "model.vision_model.encoder.layers.0.self_attn.q_proj.weight", "model.vision_model.embeddings.position_embedding.weight",
             "model.vision_model.embeddings.patch_embedding.weight"]


def build(seed, cfg_dict=None):
    Ref = ref_idefics2_classes()
    cfg = Idefics2Config(**(cfg_dict or CFG))
    cfg._attn_implementation = "eager"
    torch.manual_seed(seed)
    model = Ref(cfg)
    g = torch.Generator().manual_seed(seed + 1)
    with torch.no_grad():
        for n, p in model.named_parameters():
            if p.dim() == 1:
                p.add_(0.05 * torch.randn(p.shape, generator=g))
            if n.endswith("perceiver_resampler.latents"):
                p.copy_(torch.randn(p.shape, generator=g))
    return model.train()


def run(model, name, cfg_dict=None, extra=None, **inputs):
    out = model(use_cache=False, **inputs)
    model.zero_grad()
    out.loss.backward()
    grads = {k: p.grad.detach().clone() for k, p in model.named_parameters() if k in GRAD_KEYS and p.grad is not None}
    fx = dict(cfg=cfg_dict or CFG, state_dict={k: v.detach().clone() for k, v in model.state_dict().items()}, inputs=inputs,
              logits=out.logits.detach().clone(), loss=out.loss.detach().clone(), grads=grads)
    fx.update(extra or {})
    torch.save(fx, os.path.join(OUT, name))
    print(name, tuple(out.logits.shape), float(out.loss.detach()), len(grads))


def main():
    g = torch.Generator().manual_seed(11)
    model = build(0)
    ids = torch.randint(1, 299, (1, 40), generator=g)
    ids[0, 2:10] = 300; ids[0, 20:28] = 300
    labels = ids.clone(); labels[0, :5] = 300
    run(model, "idefics2_full.pt", input_ids=ids, attention_mask=torch.ones_like(ids),
        pixel_values=torch.randn(1, 2, 3, 112, 112, generator=g), labels=labels)

    model = build(5)
    ids = torch.randint(1, 299, (2, 36), generator=g)
    ids[0, 1:9] = 300; ids[0, 15:23] = 300          # sample 0: two images
    ids[1, 4:12] = 300; ids[1, 30:] = 0             # sample 1: one image (+ right padding)
    att = (ids != 0).long()
    pv = torch.randn(2, 2, 3, 112, 112, generator=g)
    pam = torch.ones(2, 2, 112, 112, dtype=torch.bool)
    pam[0, 1, 70:, :] = False; pam[0, 1, :, 84:] = False       # 5 x 6 patches valid
    pv[0, 1][:, 70:, :] = 0; pv[0, 1][:, :, 84:] = 0
    pv[1, 1] = 0; pam[1, 1] = False                             # padding image
    labels = ids.clone(); labels[ids == 0] = 300
    run(model, "idefics2_ragged.pt", input_ids=ids, attention_mask=att, pixel_values=pv, pixel_attention_mask=pam,
        labels=labels)

    # Mistral sliding window SMALLER than the sequence (transformers mistral/modeling_mistral.py: kv_idx > q_idx - window):
    # forward/backward at S = 40 > window = 12, plus a cache-free greedy continuation that keeps sliding
    import copy
    cfg_sw = copy.deepcopy(CFG)
    cfg_sw["text_config"]["sliding_window"] = 12
    model = build(9, cfg_sw)
    with torch.no_grad():
        for n, p in model.named_parameters():
            if p.dim() >= 2 and "vision_model" not in n:
                p.mul_(6.0)
    ids = torch.randint(1, 299, (1, 40), generator=g)
    ids[0, 3:11] = 300
    pv = torch.randn(1, 1, 3, 112, 112, generator=g)
    model.eval()
    seq = ids.clone()
    with torch.no_grad():
        for _ in range(8):
            lg = model(input_ids=seq, attention_mask=torch.ones_like(seq), pixel_values=pv, use_cache=False).logits
            seq = torch.cat([seq, lg[:, -1].argmax(-1, keepdim=True)], 1)
    model.train()
    run(model, "idefics2_sliding.pt", cfg_dict=cfg_sw, extra={"generated": seq}, input_ids=ids,
        attention_mask=torch.ones_like(ids), pixel_values=pv, labels=ids.clone())


if __name__ == "__main__":
    main()
