"""TEST INFRASTRUCTURE ONLY.  Batched greedy decoding oracle for the decode-time branch of the reference
(mantis/models/mllava/modeling_llava.py:477-508: with uneven image counts the merged prompts of a batch contain slots that
must stay masked while new tokens are appended).

  greedy_llava_batch.pt   2 left-padded prompts (2 images / 1 image, SigLIP "full" -> 64 rows per image), 8 greedy steps by
                          the cache-free loop: every step re-runs the UNMODIFIED reference forward on the whole growing
                          batch (ids, attention_mask extended with ones) and takes argmax(logits[:, -1]).
A cached implementation has to reproduce these tokens, which requires the pad-slot bookkeeping of :477-508.
"""
import os
import sys

import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from oracle.make_golden import OUT, build  # noqa: E402


def main():
    g = torch.Generator().manual_seed(77)
    model, meta = build("siglip", 112, "full", "llava", seed=9)
    with torch.no_grad():
        for n, p in model.named_parameters():
            if p.dim() >= 2 and "vision_tower" not in n:
                p.mul_(8.0)
        E = model.language_model.model.embed_tokens.weight
        model.language_model.lm_head.weight.copy_(torch.roll(E, 1, 0) * 0.5 + model.language_model.lm_head.weight * 0.5)
    model.eval()
    ids = torch.randint(0, 299, (2, 20), generator=g)
    ids[0, 4] = 300; ids[0, 11] = 300                     # sample 0: two images, no padding
    ids[1, :6] = 301; ids[1, 9] = 300                      # sample 1: left padding, one image
    att = (ids != 301).long()
    pv = torch.randn(3, 3, 112, 112, generator=g)
    seq, mask = ids.clone(), att.clone()
    with torch.no_grad():
        for _ in range(8):
            lg = model(input_ids=seq, pixel_values=pv, attention_mask=mask).logits
            nxt = lg[:, -1].argmax(-1, keepdim=True)
            seq = torch.cat([seq, nxt], 1)
            mask = torch.cat([mask, torch.ones_like(nxt)], 1)
    torch.save(dict(meta=meta, state_dict={k: v.detach().clone() for k, v in model.state_dict().items()},
                    input_ids=ids, attention_mask=att, pixel_values=pv, generated=seq),
               os.path.join(OUT, "greedy_llava_batch.pt"))
    print("greedy batch", seq[:, 20:].tolist())


if __name__ == "__main__":
    main()
