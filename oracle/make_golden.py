"""TEST INFRASTRUCTURE ONLY.  Generates tests/golden/*.pt by running the UNMODIFIED reference
(/root/reference, via oracle/ref_shim.py) on CPU in fp32.  Run in the build container only:

    python oracle/make_golden.py

Fixtures (all fp32, CPU):
  llava_siglip_full.pt   BASELINE config 1: 2-layer SigLIP (224 px, "full") + 2-layer LLaMA, 2 images + 64 text tokens
  llava_clip_default.pt  CLIP tower (class token, pre_layrnorm, quick_gelu) + "default" strategy
  llava_batch_pad.pt     batch of 2 with uneven image counts + right padding (zero padding_idx embedding row)
  mllava_clip.pt         MLlavaForConditionalGeneration (image_type_embeddings + CLIPEncoder xatten layers)
  merge_kat.pt           _merge_input_ids_with_image_features known-answer vectors incl. SURVEY 8c KAT-1..5
  greedy_llava.pt        cache-free greedy decode (argmax of reference forward on the growing sequence)
Each model fixture holds: config kwargs, state_dict, inputs, logits, loss and a few parameter gradients.
"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from oracle.ref_shim import ref_llava_classes  # noqa: E402

from transformers import CLIPVisionConfig, LlamaConfig, SiglipVisionConfig  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden")
GRAD_KEYS = [
    "multi_modal_projector.linear_1.weight", "multi_modal_projector.linear_2.bias",
    "language_model.lm_head.weight", "language_model.model.embed_tokens.weight",
    "language_model.model.layers.0.self_attn.q_proj.weight", "language_model.model.layers.0.self_attn.k_proj.weight",
    "language_model.model.layers.1.mlp.down_proj.weight", "language_model.model.layers.1.input_layernorm.weight",
    "language_model.model.norm.weight",
    "vision_tower.vision_model.encoder.layers.0.self_attn.q_proj.weight",
    "vision_tower.vision_model.encoder.layers.0.mlp.fc1.bias",
    "vision_tower.vision_model.embeddings.patch_embedding.weight",
    "vision_tower.vision_model.embeddings.position_embedding.weight",
]


def vision_cfg(kind, image_size):
    kw = dict(hidden_size=64, intermediate_size=128, num_hidden_layers=2, num_attention_heads=4,
              image_size=image_size, patch_size=14)
    if kind == "siglip":
        return "siglip_vision_model", kw
    return "clip_vision_model", dict(kw, projection_dim=32)


def text_cfg(pad=None):
    return dict(hidden_size=64, intermediate_size=128, num_hidden_layers=2, num_attention_heads=4,
                num_key_value_heads=2, vocab_size=320, pad_token_id=pad, rms_norm_eps=1e-5, rope_theta=500000.0,
                max_position_embeddings=4096, tie_word_embeddings=False)


def build(kind, image_size, strategy, cls, pad_in_text=None, seed=0):
    LlavaConfig, RefLlava, RefMLlava = ref_llava_classes()
    mt, vkw = vision_cfg(kind, image_size)
    vc = (SiglipVisionConfig if kind == "siglip" else CLIPVisionConfig)(**vkw)
    tc = LlamaConfig(**text_cfg(pad_in_text))
    cfg_kwargs = dict(image_token_index=300, pad_token_id=301, vision_feature_select_strategy=strategy,
                      vision_feature_layer=-2, vocab_size=320, projector_hidden_act="gelu")
    cfg = LlavaConfig(vision_config=vc, text_config=tc, **cfg_kwargs)
    torch.manual_seed(seed)
    model = (RefMLlava if cls == "mllava" else RefLlava)(cfg)
    # make sure every parameter is non-trivial (norm weights / biases are ones/zeros after init)
    g = torch.Generator().manual_seed(seed + 1)
    with torch.no_grad():
        for n, p in model.named_parameters():
            if p.dim() == 1:
                p.add_(0.05 * torch.randn(p.shape, generator=g))
        if pad_in_text is not None:
            model.language_model.model.embed_tokens.weight[pad_in_text].zero_()
    meta = dict(vision_kind=kind, vision_model_type=mt, vision_kwargs=vkw, text_kwargs=text_cfg(pad_in_text),
                cfg_kwargs=cfg_kwargs, cls=cls)
    return model, meta


def run_case(model, meta, input_ids, pixel_values, attention_mask, labels, name):
    model.train()
    out = model(input_ids=input_ids, pixel_values=pixel_values, attention_mask=attention_mask, labels=labels)
    model.zero_grad()
    out.loss.backward()
    grads = {k: p.grad.detach().clone() for k, p in model.named_parameters() if k in GRAD_KEYS and p.grad is not None}
    extra = [k for k, _ in model.named_parameters() if k.startswith(("image_type_embeddings", "vision_xatten_layers.layers.0.self_attn.q_proj.weight"))]
    for k, p in model.named_parameters():
        if k in extra and p.grad is not None:
            grads[k] = p.grad.detach().clone()
    fx = dict(meta=meta, state_dict={k: v.detach().clone() for k, v in model.state_dict().items()},
              input_ids=input_ids, pixel_values=pixel_values, attention_mask=attention_mask, labels=labels,
              logits=out.logits.detach().clone(), loss=out.loss.detach().clone(), grads=grads)
    torch.save(fx, os.path.join(OUT, name))
    print(name, "logits", tuple(out.logits.shape), "loss", float(out.loss), "grads", len(grads))
    return fx


def main():
    os.makedirs(OUT, exist_ok=True)
    torch.manual_seed(0)
    # ---- config 1 (SigLIP, "full") -------------------------------------------------------------------
    model, meta = build("siglip", 224, "full", "llava")
    g = torch.Generator().manual_seed(123)
    ids = torch.randint(0, 299, (1, 66), generator=g); ids[0, 3] = 300; ids[0, 40] = 300
    pv = torch.randn(2, 3, 224, 224, generator=g)
    labels = ids.clone(); labels[ids == 300] = -100; labels[0, :10] = -100
    run_case(model, meta, ids, pv, torch.ones_like(ids), labels, "llava_siglip_full.pt")

    # greedy decode oracle (cache-free loop over the reference forward); weights scaled up so that the argmax
    # actually depends on the context (a 0.02-std random model emits one constant token)
    model, meta = build("siglip", 224, "full", "llava", seed=5)
    with torch.no_grad():
        for n, p in model.named_parameters():
            if p.dim() >= 2 and "vision_tower" not in n:
                p.mul_(8.0)
        # head row j = embedding row j-1: the residual path makes token t favour t+1, perturbed by the context
        E = model.language_model.model.embed_tokens.weight
        model.language_model.lm_head.weight.copy_(torch.roll(E, 1, 0) * 0.5 + model.language_model.lm_head.weight * 0.5)
    model.eval()
    with torch.no_grad():
        seq = ids.clone()
        for _ in range(16):
            lg = model(input_ids=seq, pixel_values=pv, attention_mask=torch.ones_like(seq)).logits
            seq = torch.cat([seq, lg[:, -1].argmax(-1, keepdim=True)], 1)
    torch.save(dict(meta=meta, state_dict={k: v.detach().clone() for k, v in model.state_dict().items()},
                    input_ids=ids, pixel_values=pv, generated=seq), os.path.join(OUT, "greedy_llava.pt"))
    print("greedy", seq[0, 66:].tolist())

    # ---- CLIP tower, "default" ------------------------------------------------------------------------
    model, meta = build("clip", 112, "default", "llava", seed=1)
    ids = torch.randint(0, 299, (1, 40), generator=g); ids[0, 0] = 300; ids[0, 20] = 300; ids[0, 21] = 300
    pv = torch.randn(3, 3, 112, 112, generator=g)
    labels = ids.clone(); labels[ids == 300] = -100
    run_case(model, meta, ids, pv, torch.ones_like(ids), labels, "llava_clip_default.pt")

    # ---- batch of 2, uneven images, right padding ------------------------------------------------------
    model, meta = build("siglip", 112, "full", "llava", pad_in_text=None, seed=2)   # non-zero pad embedding row
    ids = torch.randint(0, 299, (2, 30), generator=g)
    ids[0, 2] = 300; ids[0, 17] = 300
    ids[1, 5] = 300; ids[1, 22:] = 301
    att = (ids != 301).long()
    pv = torch.randn(3, 3, 112, 112, generator=g)
    labels = ids.clone(); labels[ids == 300] = -100; labels[ids == 301] = -100
    try:
        run_case(model, meta, ids, pv, att, labels, "llava_batch_pad.pt")
    except Exception as e:  # the reference's merge has known failure modes with padding; record it
        print("llava_batch_pad reference failed:", type(e).__name__, e)

    # ---- MLlava ---------------------------------------------------------------------------------------
    model, meta = build("clip", 112, "default", "mllava", seed=3)
    ids = torch.randint(0, 299, (1, 24), generator=g); ids[0, 1] = 300; ids[0, 12] = 300
    pv = torch.randn(2, 3, 112, 112, generator=g)
    labels = ids.clone(); labels[ids == 300] = -100
    run_case(model, meta, ids, pv, torch.ones_like(ids), labels, "mllava_clip.pt")

    # ---- merge KATs -----------------------------------------------------------------------------------
    from oracle.ref_shim import ref_llava_classes as _r
    LlavaConfig, RefLlava, _ = _r()
    vc = SiglipVisionConfig(hidden_size=16, intermediate_size=32, num_hidden_layers=1, num_attention_heads=2,
                            image_size=28, patch_size=14)
    tc = LlamaConfig(hidden_size=16, intermediate_size=32, num_hidden_layers=1, num_attention_heads=2,
                     num_key_value_heads=2, vocab_size=32)
    m = RefLlava(LlavaConfig(vision_config=vc, text_config=tc, image_token_index=9, pad_token_id=0, vocab_size=32))
    cases = []

    def add(ids, P, n_images, labels=True, zero_pad=False, D=8):
        ids = torch.tensor(ids)
        emb = (ids[..., None].float() + 0.5).expand(*ids.shape, D).clone()
        if zero_pad:
            emb[ids == 0] = 0
        att = (ids != 0).long()
        feats = (1000 + torch.arange(n_images * P).float()).view(n_images, P, 1).expand(n_images, P, D).clone()
        lab = ids.clone() if labels else None
        rec = dict(input_ids=ids, inputs_embeds=emb, attention_mask=att, labels=lab, image_features=feats, P=P)
        try:
            r = m._merge_input_ids_with_image_features(feats, emb, ids, att, lab)
            rec.update(final_embedding=r[0], final_attention_mask=r[1], final_labels=r[2], position_ids=r[3], error=None)
        except Exception as e:
            rec.update(error=type(e).__name__)
        cases.append(rec)

    add([[5, 9, 6, 9, 7]], 3, 2)                               # KAT-1
    add([[5, 9, 6, 9, 7], [9, 4, 3, 0, 0]], 3, 3)              # KAT-2 right pad, uneven images
    add([[5, 9, 6, 9, 7], [0, 0, 9, 4, 3]], 3, 3)              # KAT-3 left pad
    add([[9, 1, 2]], 2, 1, labels=False)                       # KAT-4
    add([[5, 9, 6]], 3, 2)                                     # KAT-5 -> ValueError
    add([[5, 9, 6, 9, 7], [9, 4, 3, 0, 0]], 3, 3, zero_pad=True)
    add([[5, 9, 6, 9, 7], [0, 0, 9, 4, 3]], 3, 3, zero_pad=True)
    add([[1, 2, 3, 9], [9, 9, 9, 9]], 4, 5)
    rng = np.random.default_rng(7)
    for _ in range(40):
        B = int(rng.integers(1, 4)); T = int(rng.integers(2, 10)); P = int(rng.integers(1, 5))
        ids = rng.integers(1, 12, size=(B, T))
        mode = int(rng.integers(0, 3))
        for b in range(B):
            npad = int(rng.integers(0, T)) if mode else 0
            if mode == 1 and npad:
                ids[b, T - npad:] = 0
            if mode == 2 and npad:
                ids[b, :npad] = 0
        n = int((ids == 9).sum())
        if n == 0:
            continue
        add(ids.tolist(), P, n, labels=bool(rng.integers(0, 2)), zero_pad=bool(rng.integers(0, 2)))
    torch.save(cases, os.path.join(OUT, "merge_kat.pt"))
    print("merge cases", len(cases), "errors", sum(c["error"] is not None for c in cases))


if __name__ == "__main__":
    main()
