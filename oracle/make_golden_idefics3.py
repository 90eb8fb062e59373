"""TEST INFRASTRUCTURE ONLY.  Golden fixtures for Idefics3ForConditionalGeneration from the UNMODIFIED reference
(mantis/models/idefics3/modeling_idefics3.py via oracle/ref_shim.py), CPU fp32, eager attention.

  idefics3_full.pt    1 sample, 2 full-resolution images (pixel_attention_mask = None); 8x8 patches, scale_factor 2
                      -> 16 image tokens per image
  idefics3_ragged.pt  2 samples x 2 image slots: one image with a partial pixel mask (NaViT position ids + patch padding
                      mask in the tower), one all-zero padding image the model must drop, right-padded text,
                      labels with -100 (the ignore index of this family's loss, modeling_idefics3.py:1166-1180)
"""
import os
import sys

import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from oracle.ref_shim import ref_idefics3_classes  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden")
CFG = dict(
    vision_config=dict(hidden_size=64, intermediate_size=128, num_hidden_layers=2, num_attention_heads=4, image_size=112,
                       patch_size=14),
    text_config=dict(model_type="llama", hidden_size=64, intermediate_size=128, num_hidden_layers=2,
                     num_attention_heads=4, num_key_value_heads=2, vocab_size=320, pad_token_id=0, rms_norm_eps=1e-5,
                     rope_theta=500000.0),
    image_token_id=300, scale_factor=2, tie_word_embeddings=False)
GRAD_KEYS = ["lm_head.weight", "model.text_model.embed_tokens.weight", "model.text_model.layers.0.self_attn.q_proj.weight",
             "model.text_model.layers.1.mlp.down_proj.weight", "model.connector.modality_projection.proj.weight",
             "model.vision_model.post_layernorm.weight", "model.vision_model.encoder.layers.0.self_attn.q_proj.weight",
             "model.vision_model.encoder.layers.1.mlp.fc2.bias", "model.vision_model.embeddings.position_embedding.weight",
             "model.vision_model.embeddings.patch_embedding.weight"]


def build(seed):
    Cfg, _, Ref = ref_idefics3_classes()
    cfg = Cfg(**CFG)
    for c in (cfg, cfg.vision_config, cfg.text_config):
        c._attn_implementation = "eager"
    torch.manual_seed(seed)
    model = Ref(cfg)
    g = torch.Generator().manual_seed(seed + 1)
    with torch.no_grad():
        for n, p in model.named_parameters():
            if p.dim() == 1:
                p.add_(0.05 * torch.randn(p.shape, generator=g))
    return model.train()


def run(model, name, **inputs):
    out = model(use_cache=False, **inputs)
    model.zero_grad()
    out.loss.backward()
    grads = {k: p.grad.detach().clone() for k, p in model.named_parameters() if k in GRAD_KEYS and p.grad is not None}
    fx = dict(cfg=CFG, state_dict={k: v.detach().clone() for k, v in model.state_dict().items()}, inputs=inputs,
              logits=out.logits.detach().clone(), loss=out.loss.detach().clone(), grads=grads)
    torch.save(fx, os.path.join(OUT, name))
    print(name, tuple(out.logits.shape), float(out.loss.detach()), len(grads))


def main():
    g = torch.Generator().manual_seed(13)
    model = build(0)
    ids = torch.randint(1, 299, (1, 56), generator=g)
    ids[0, 2:18] = 300; ids[0, 30:46] = 300
    labels = ids.clone(); labels[0, :5] = -100; labels[ids == 300] = -100
    run(model, "idefics3_full.pt", input_ids=ids, attention_mask=torch.ones_like(ids),
        pixel_values=torch.randn(1, 2, 3, 112, 112, generator=g), labels=labels)

    model = build(5)
    ids = torch.randint(1, 299, (2, 52), generator=g)
    ids[0, 1:17] = 300; ids[0, 25:41] = 300         # sample 0: two images
    ids[1, 4:20] = 300; ids[1, 44:] = 0             # sample 1: one image (+ right padding)
    att = (ids != 0).long()
    pv = torch.randn(2, 2, 3, 112, 112, generator=g)
    pam = torch.ones(2, 2, 112, 112, dtype=torch.bool)
    pam[0, 1, 84:, :] = False; pam[0, 1, :, 56:] = False       # 6 x 4 patches valid
    pv[0, 1][:, 84:, :] = 0; pv[0, 1][:, :, 56:] = 0
    pv[1, 1] = 0; pam[1, 1] = False                             # padding image
    labels = ids.clone(); labels[ids == 0] = -100; labels[ids == 300] = -100
    run(model, "idefics3_ragged.pt", input_ids=ids, attention_mask=att, pixel_values=pv, pixel_attention_mask=pam,
        labels=labels)


if __name__ == "__main__":
    main()
