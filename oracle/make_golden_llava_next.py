"""TEST INFRASTRUCTURE ONLY.  Golden fixture for LlavaNextForConditionalGeneration from the UNMODIFIED reference
(mantis/models/mllava_next/modeling_llava_next.py via oracle/ref_shim.py), CPU fp32, eager attention.

  llava_next_batch.pt   2 samples (1 and 2 images, the first right-padded with pad tokens so that the pad-row zeroing
                        of ref:455-461 is exercised), per-image crop stacks of different depth (3 / 1 / 2 crops: only the
                        base crop may influence the result), CLIP tower with the "default" strategy -> 64 + 1 rows/image
"""
import os
import sys

import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from oracle.ref_shim import ref_llava_next_classes  # noqa: E402

from transformers import CLIPVisionConfig, LlamaConfig  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden")
VISION = dict(hidden_size=64, intermediate_size=128, num_hidden_layers=2, num_attention_heads=4, image_size=112, patch_size=14)
TEXT = dict(hidden_size=64, intermediate_size=128, num_hidden_layers=2, num_attention_heads=4, num_key_value_heads=2,
            vocab_size=320, rms_norm_eps=1e-5, rope_theta=500000.0)
CFG = dict(image_token_index=300, pad_token_id=301, vocab_size=320, ignore_index=-100,
           image_grid_pinpoints=[[224, 224], [112, 224]], vision_feature_select_strategy="default", vision_feature_layer=-2)
GRAD_KEYS = ["image_newline", "multi_modal_projector.linear_1.weight", "multi_modal_projector.linear_2.bias",
             "language_model.lm_head.weight", "language_model.model.embed_tokens.weight",
             "language_model.model.layers.0.self_attn.q_proj.weight", "language_model.model.layers.1.mlp.down_proj.weight",
             "vision_tower.vision_model.encoder.layers.0.self_attn.q_proj.weight",
             "vision_tower.vision_model.embeddings.patch_embedding.weight"]


def main():
    Cfg, Ref = ref_llava_next_classes()
    cfg = Cfg(vision_config=CLIPVisionConfig(**VISION), text_config=LlamaConfig(**TEXT), **CFG)
    cfg._attn_implementation = "eager"
    torch.manual_seed(3)
    model = Ref(cfg)
    g = torch.Generator().manual_seed(4)
    with torch.no_grad():
        model.image_newline.copy_(0.05 * torch.randn(64, generator=g))
        for n, p in model.named_parameters():
            if p.dim() == 1 and n != "image_newline":
                p.add_(0.05 * torch.randn(p.shape, generator=g))
    model.train()
    ids = torch.randint(0, 299, (2, 24), generator=g)
    ids[0, 3] = 300; ids[0, 17:] = 301                       # sample 0: one image, right padding
    ids[1, 5] = 300; ids[1, 12] = 300                        # sample 1: two images
    att = (ids != 301).long()
    labels = ids.clone(); labels[ids == 301] = -100; labels[ids == 300] = -100; labels[:, :2] = -100
    pv = [torch.randn(3, 3, 112, 112, generator=g), torch.randn(1, 3, 112, 112, generator=g),
          torch.randn(2, 3, 112, 112, generator=g)]
    sizes = torch.tensor([[224, 224], [112, 112], [112, 224]])
    out = model(input_ids=ids, pixel_values=pv, image_sizes=sizes, attention_mask=att, labels=labels, use_cache=False)
    model.zero_grad()
    out.loss.backward()
    grads = {k: p.grad.detach().clone() for k, p in model.named_parameters() if k in GRAD_KEYS and p.grad is not None}
    fx = dict(vision=VISION, text=TEXT, cfg=CFG, state_dict={k: v.detach().clone() for k, v in model.state_dict().items()},
              input_ids=ids, attention_mask=att, labels=labels, pixel_values=pv, image_sizes=sizes,
              logits=out.logits.detach().clone(), loss=out.loss.detach().clone(), grads=grads)
    torch.save(fx, os.path.join(OUT, "llava_next_batch.pt"))
    print("llava_next_batch.pt", tuple(out.logits.shape), float(out.loss.detach()), sorted(grads))


if __name__ == "__main__":
    main()
