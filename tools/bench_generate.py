#!/usr/bin/env python
"""generate() benchmark (BASELINE config 5): Mantis-8B-SigLIP random init bf16, 8 images + 256-token prompt (prefill
S = 6072) -> N greedy new tokens, bs = 1 and 16.  Reports prefill tok/s (bs * 6072 / t_prefill) and decode tok/s
(bs * N / t_decode) plus the HBM roofline of the decode step (weights 15.01 GB + 131,072 B x ctx per sequence)."""
import argparse
import json
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--bs", type=int, nargs="+", default=[1, 16])
    ap.add_argument("--new-tokens", type=int, default=128)
    ap.add_argument("--text-layers", type=int, default=32)
    a = ap.parse_args()
    from mantis_b200 import ops
    from mantis_b200.models.kv_cache import B200KVCache
    from mantis_b200.models.mllava import LlavaForConditionalGeneration, mantis_8b_siglip_llama3_config
    dev = torch.device("cuda")
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        peaks = {"hbm_gbs": 6650.0}
    cfg = mantis_8b_siglip_llama3_config(num_text_layers=a.text_layers)
    torch.manual_seed(0)
    torch.set_default_dtype(torch.bfloat16)
    with torch.device(dev):
        model = LlavaForConditionalGeneration(cfg)
    torch.set_default_dtype(torch.float32)
    model.eval()
    g = torch.Generator().manual_seed(5)
    for bs in a.bs:
        ids = torch.randint(0, 128000, (bs, 256), generator=g)
        for j in range(8):
            ids[:, j * 32 + 4] = 128256
        pv = torch.randn(bs * 8, 3, 384, 384, generator=g).bfloat16()
        ids = ids.to(dev); pv = pv.to(dev)
        att = torch.ones_like(ids)
        with torch.no_grad():
            for rep in range(2):                      # rep 0 = warm-up
                cache = B200KVCache()
                torch.cuda.synchronize()
                e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True); e2 = torch.cuda.Event(enable_timing=True)
                e0.record()
                out = model(input_ids=ids, pixel_values=pv, attention_mask=att, past_key_values=cache, use_cache=True,
                            logits_to_keep=1)
                e1.record()
                am = att
                l0 = ops.launch_count
                t_host = time.perf_counter()
                n_new = a.new_tokens if rep else 8
                for step in range(n_new):
                    nxt = out.logits[:, -1, :].argmax(-1)
                    am = torch.cat([am, torch.ones_like(nxt[:, None])], dim=1)
                    out = model(input_ids=nxt[:, None], pixel_values=pv, attention_mask=am, past_key_values=cache,
                                use_cache=True, logits_to_keep=1)
                e2.record()
                torch.cuda.synchronize()
                t_host = time.perf_counter() - t_host
        S = cache.get_seq_length() - n_new
        t_pre = e0.elapsed_time(e1) * 1e-3; t_dec = e1.elapsed_time(e2) * 1e-3
        ctx_avg = S + n_new / 2
        bytes_step = 15.01e9 * a.text_layers / 32 + 131072.0 * a.text_layers / 32 * ctx_avg * bs
        res = {"bs": bs, "prefill_len": S, "new_tokens": n_new, "prefill_tok_s": bs * S / t_pre,
               "decode_tok_s": bs * n_new / t_dec, "decode_ms_per_step": t_dec / n_new * 1e3,
               "decode_hbm_gbs": bytes_step / (t_dec / n_new) / 1e9,
               "decode_frac_of_hbm_peak": bytes_step / (t_dec / n_new) / 1e9 / peaks["hbm_gbs"],
               "host_ms_per_step": t_host / n_new * 1e3, "launches_per_step": (ops.launch_count - l0) / n_new,
               "text_layers": a.text_layers}
        print(json.dumps(res), flush=True)


if __name__ == "__main__":
    main()
