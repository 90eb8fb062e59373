#!/usr/bin/env python
"""Per-kernel micro-benchmarks at the Mantis-8B-SigLIP shapes (CUDA events on the launching stream, operand sets
cycled so every launch reads cold-in-L2 data).  Prints one JSON object per kernel with achieved TFLOP/s or GB/s and
the fraction of the measured peak (MEASURED_PEAKS.json)."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from mantis_b200 import ops  # noqa: E402

dev = torch.device("cuda")
try:
    PEAKS = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
except Exception:
    PEAKS = {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}


def timeit(fn, reps=10, warm=3):
    for i in range(warm):
        fn(i)
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(reps):
        fn(i)
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def report(name, ms, flops=None, bytes_=None, **kw):
    d = {"kernel": name, "ms": round(ms, 4)}
    if flops:
        d["tflops"] = round(flops / ms / 1e9, 1); d["frac_of_bf16_peak"] = round(d["tflops"] / PEAKS["bf16_tflops"], 3)
    if bytes_:
        d["gbs"] = round(bytes_ / ms / 1e6, 1); d["frac_of_hbm_peak"] = round(d["gbs"] / PEAKS["hbm_gbs"], 3)
    d.update(kw)
    print(json.dumps(d), flush=True)


def bench_gemm():
    M = 7864
    for (N, K, tag) in [(4096, 4096, "q/o_proj"), (1024, 4096, "k/v_proj"), (14336, 4096, "gate/up"), (4096, 14336, "down"),
                        (128258, 4096, "lm_head(M=4096)")]:
        m = 4096 if N > 100000 else M
        nset = 3
        xs = [torch.randn(m, K, device=dev).bfloat16() for _ in range(nset)]
        ws = [(torch.randn(N, K, device=dev) * 0.02).bfloat16() for _ in range(nset)]
        ldc = (N + 7) // 8 * 8
        outs = [torch.empty(m, ldc, device=dev, dtype=torch.bfloat16) for _ in range(nset)]
        ms = timeit(lambda i: ops.gemm(xs[i % nset], ws[i % nset], out=outs[i % nset][:, :N]))
        report(f"gemm fwd {tag} M={m} N={N} K={K}", ms, flops=2.0 * m * N * K)
        if N < 100000:
            gs = [torch.randn(m, N, device=dev).bfloat16() for _ in range(nset)]
            ms = timeit(lambda i: ops.gemm(gs[i % nset], ws[i % nset], trans_a=False, trans_b=False))
            report(f"gemm dgrad {tag}", ms, flops=2.0 * m * N * K)
            ms = timeit(lambda i: ops.gemm(gs[i % nset], xs[i % nset], trans_a=True, trans_b=False))
            report(f"gemm wgrad {tag}", ms, flops=2.0 * m * N * K)
            # cuBLAS on the same operands (the incumbent)
            ms = timeit(lambda i: torch.matmul(xs[i % nset], ws[i % nset].t()))
            report(f"cuBLAS fwd {tag} (incumbent)", ms, flops=2.0 * m * N * K)
        del xs, ws, outs


def bench_attention():
    B, S, H, Hkv, hd = 1, 7864, 32, 8, 128
    nset = 2
    qs = [torch.randn(B, S, H, hd, device=dev).bfloat16() for _ in range(nset)]
    ks = [torch.randn(B, S, Hkv, hd, device=dev).bfloat16() for _ in range(nset)]
    vs = [torch.randn(B, S, Hkv, hd, device=dev).bfloat16() for _ in range(nset)]
    scale = hd ** -0.5
    fl = 4.0 * S * S * hd * H / 2            # causal: half the square
    ms = timeit(lambda i: ops.attention_fwd(qs[i % nset], ks[i % nset], vs[i % nset], True, None, scale), reps=5)
    report("attn fwd causal S=7864 32q/8kv hd128", ms, flops=fl)
    ops.ATTN_FWD2 = not ops.ATTN_FWD2
    ms = timeit(lambda i: ops.attention_fwd(qs[i % nset], ks[i % nset], vs[i % nset], True, None, scale), reps=5)
    report(f"attn fwd causal S=7864 (ATTN_FWD2={ops.ATTN_FWD2})", ms, flops=fl)
    ops.ATTN_FWD2 = not ops.ATTN_FWD2
    o, lse = ops.attention_fwd(qs[0], ks[0], vs[0], True, None, scale)
    do = torch.randn_like(o)
    ms = timeit(lambda i: ops.attention_bwd(qs[0], ks[0], vs[0], o, do, lse, True, None, scale, fast=True), reps=5)
    report("attn bwd causal S=7864 (dKV + dQ kernels)", ms, flops=2.5 * fl, executed_flops_ratio=3.5 / 2.5)
    try:
        from flash_attn import flash_attn_func
        ms = timeit(lambda i: flash_attn_func(qs[i % nset], ks[i % nset], vs[i % nset], causal=True), reps=5)
        report("flash_attn 2.8 fwd (incumbent, HMMA)", ms, flops=fl)
        q = qs[0].clone().requires_grad_(True); k = ks[0].clone().requires_grad_(True); v = vs[0].clone().requires_grad_(True)
        out = flash_attn_func(q, k, v, causal=True)
        ms = timeit(lambda i: torch.autograd.grad(out, (q, k, v), do, retain_graph=True), reps=5)
        report("flash_attn 2.8 bwd (incumbent, HMMA)", ms, flops=2.5 * fl)
    except Exception as e:  # noqa
        print(json.dumps({"kernel": "flash_attn incumbent", "error": str(e)[:200]}))


def bench_rowwise():
    M, D = 7864, 4096
    x = torch.randn(M, D, device=dev).bfloat16(); w = torch.ones(D, device=dev).bfloat16()
    ms = timeit(lambda i: ops.rms_norm(x, w, 1e-5))
    report("rmsnorm fwd [7864,4096]", ms, bytes_=2 * M * D * 2)
    xg = x.clone().requires_grad_(True); wg = w.clone().requires_grad_(True)
    yy, rr = ops.rms_norm_res(xg, wg, 1e-5)
    gy = torch.randn_like(yy); gr = torch.randn_like(rr)
    ms = timeit(lambda i: torch.autograd.grad((yy, rr), (xg, wg), (gy, gr), retain_graph=True))
    report("rmsnorm bwd (dx + dw, residual gradient folded in) [7864,4096]", ms, bytes_=4 * M * D * 2,
           note="algorithmic = x, dy, dres read + dx written")
    g = torch.randn(M, 14336, device=dev).bfloat16(); u = torch.randn(M, 14336, device=dev).bfloat16()
    ms = timeit(lambda i: ops.swiglu(g, u))
    report("swiglu fwd [7864,14336]", ms, bytes_=3 * M * 14336 * 2)
    # merge (scatter): config-2 sized, B = 4
    B, T, P = 4, 2048, 728
    ids = torch.randint(0, 128000, (B, T), device=dev)
    for j in range(8):
        ids[:, j * 256 + 16] = 128256
    emb = torch.randn(B, T, D, device=dev).bfloat16(); feats = torch.randn(32, P, D, device=dev).bfloat16()
    att = torch.ones_like(ids)
    ms = timeit(lambda i: ops.merge_input_ids_with_image_features(feats, emb, ids, att, ids, 128256, 128257), reps=10)
    S = T + 8 * (P - 1)
    report("merge (plan+index+rows, incl. 1 host sync) B=4", ms, bytes_=2 * B * S * D * 2)
    srcmap = torch.empty((B, S), dtype=torch.int32, device=dev)
    ws, hdr = ops.merge_plan(ids, emb, P, 128256, 128257)
    om = torch.empty((B, S), dtype=torch.int64, device=dev); op = torch.empty_like(om); ol = torch.empty_like(om)
    ops._call("mb200_merge_index", ops._p(ids), ops._p(att), ops._p(ids), ops._p(ws), B, T, P, S, int(hdr[1]), 128256, -100,
              ops._p(srcmap), ops._p(om), ops._p(ol), ops._p(op), ops._st())
    outs = [torch.empty((B, S, D), dtype=torch.bfloat16, device=dev) for _ in range(2)]
    f2 = feats.reshape(-1, D)
    ms = timeit(lambda i: ops._call("mb200_merge_rows", ops._p(srcmap), ops._p(emb), ops._p(f2), ops._p(outs[i % 2]), B, S, T,
                                    D * 2, f2.shape[0], ops._st()), reps=20)
    report("merge_rows_kernel alone (row scatter) B=4", ms, bytes_=2 * B * S * D * 2)
    # CE on one LM-head chunk
    n, V = 4096, 128258
    ld = (V + 7) // 8 * 8
    lg = torch.randn(n, ld, device=dev).bfloat16(); lab = torch.randint(0, V, (n,), device=dev)
    lr = torch.empty(n, device=dev); inv = torch.ones(1, device=dev)
    ms = timeit(lambda i: ops._call("mb200_ce_fwd_bwd", ops._p(lg), ops._p(lab), ops._p(lr), None, ops._p(lg), n, V, ld,
                                    ops._p(inv), 1.0, 1, ops._st()), reps=5)
    report("ce_fwd_bwd chunk [4096,128258]", ms, bytes_=2 * n * V * 2, note="algorithmic = 1 read + 1 write")


def bench_decode():
    """decode-step kernels at bs = 1 and 16 (weights cycled so each launch streams from HBM)"""
    import ctypes
    from mantis_b200 import _lib
    L = _lib.lib()
    for B in (1, 16):
        nset = 4
        x = torch.randn(B, 4096, device=dev).bfloat16(); xi = torch.randn(B, 14336, device=dev).bfloat16()
        wq = [(torch.randn(4096, 4096, device=dev) * 0.02).bfloat16() for _ in range(nset)]
        wk = [(torch.randn(1024, 4096, device=dev) * 0.02).bfloat16() for _ in range(nset)]
        wv = [(torch.randn(1024, 4096, device=dev) * 0.02).bfloat16() for _ in range(nset)]
        wg = [(torch.randn(14336, 4096, device=dev) * 0.02).bfloat16() for _ in range(nset)]
        wu = [(torch.randn(14336, 4096, device=dev) * 0.02).bfloat16() for _ in range(nset)]
        wd = [(torch.randn(4096, 14336, device=dev) * 0.02).bfloat16() for _ in range(nset)]
        q = torch.empty(B, 4096, device=dev, dtype=torch.bfloat16); k = torch.empty(B, 1024, device=dev, dtype=torch.bfloat16)
        v = torch.empty_like(k); act = torch.empty(B, 14336, device=dev, dtype=torch.bfloat16)
        st = ops._st
        ms = timeit(lambda i: ops._call("mb200_skinny_gemm3_bf16", ops._p(x), ops._p(wq[i % nset]), ops._p(wk[i % nset]), ops._p(wv[i % nset]),
                                        ops._p(q), ops._p(k), ops._p(v), B, 4096, 1024, 1024, 4096, 4096, 4096, st()), reps=20)
        report(f"decode qkv skinny gemm bs={B}", ms, bytes_=6144 * 4096 * 2)
        ms = timeit(lambda i: ops._call("mb200_skinny_swiglu_bf16", ops._p(x), ops._p(wg[i % nset]), ops._p(wu[i % nset]), ops._p(act),
                                        B, 14336, 4096, 4096, 4096, 14336, st()), reps=20)
        report(f"decode gate/up + swiglu bs={B}", ms, bytes_=2 * 14336 * 4096 * 2)
        ms = timeit(lambda i: ops._call("mb200_skinny_gemm_bf16", ops._p(xi), ops._p(wd[i % nset]), ops._p(q), None, ops._p(q), B, 4096, 14336,
                                        14336, 14336, 4096, 4096, st()), reps=20)
        report(f"decode down proj (+residual) bs={B}", ms, bytes_=4096 * 14336 * 2)
        del wq, wk, wv, wg, wu, wd
        wl = (torch.randn(128258, 4096, device=dev) * 0.02).bfloat16()
        lg = torch.empty(B, 128264, device=dev, dtype=torch.bfloat16)
        ms = timeit(lambda i: ops._call("mb200_skinny_gemm_bf16", ops._p(x), ops._p(wl), ops._p(lg), None, None, B, 128258, 4096, 4096, 4096,
                                        128264, 0, st()), reps=5)
        report(f"decode lm_head bs={B}", ms, bytes_=128258 * 4096 * 2)
        del wl
        ctx = 6200
        qd = torch.randn(B, 1, 32, 128, device=dev).bfloat16()
        kc = torch.randn(B, ctx, 8, 128, device=dev).bfloat16(); vc = torch.randn(B, ctx, 8, 128, device=dev).bfloat16()
        ms = timeit(lambda i: ops.decode_attention(qd, kc, vc, ctx, None, 128 ** -0.5), reps=20)
        report(f"decode attention ctx={ctx} bs={B}", ms, bytes_=2.0 * B * ctx * 8 * 128 * 2)
        # same kernel over a head-major cache ([B, Hkv, ctx, hd] storage, passed as a strided view): a tile is one contiguous
        # 8 KB piece instead of 32 x 256 B at a 2 KB stride -- isolates the DRAM-locality cost of the token-major layout
        kh = torch.randn(B, 8, ctx, 128, device=dev).bfloat16().permute(0, 2, 1, 3)
        vh = torch.randn(B, 8, ctx, 128, device=dev).bfloat16().permute(0, 2, 1, 3)
        ms = timeit(lambda i: ops.decode_attention(qd, kh, vh, ctx, None, 128 ** -0.5), reps=20)
        report(f"decode attention ctx={ctx} bs={B}, head-major cache view", ms, bytes_=2.0 * B * ctx * 8 * 128 * 2)
        del kh, vh
        xr = torch.randn(B, 4096, device=dev).bfloat16(); w = torch.ones(4096, device=dev).bfloat16()
        ms = timeit(lambda i: ops.rms_norm(xr, w, 1e-5), reps=50)
        report(f"decode rmsnorm bs={B} (launch-latency bound)", ms, bytes_=2 * B * 4096 * 2)


if __name__ == "__main__":
    which = sys.argv[1:] or ["gemm", "attn", "row"]
    if "gemm" in which:
        bench_gemm()
    if "attn" in which:
        bench_attention()
    if "row" in which:
        bench_rowwise()
    if "decode" in which:
        bench_decode()
