#!/usr/bin/env python
"""Launches each hot kernel a handful of times at the bench shapes so that `ncu --set full -k regex:<name>` captures stay
short:  python tools/ncu_targets.py gemm|gemm2cta|attn|merge|rmsnorm|ce"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from mantis_b200 import ops  # noqa: E402

dev = torch.device("cuda")
which = sys.argv[1] if len(sys.argv) > 1 else "gemm"
M = 7864
if which in ("gemm", "gemm2cta"):
    ops.GEMM_2CTA = which == "gemm2cta"
    x = torch.randn(M, 4096, device=dev).bfloat16(); w = (torch.randn(14336, 4096, device=dev) * 0.02).bfloat16()
    g = torch.randn(M, 14336, device=dev).bfloat16()
    for _ in range(4):
        ops.gemm(x, w)                                   # fwd  gate/up
        ops.gemm(g, w, trans_a=False, trans_b=False)     # dgrad
        ops.gemm(g, x, trans_a=True, trans_b=False)      # wgrad
elif which == "swiglu":                                  # the MLP with SwiGLU fused into the GEMM epilogues, fwd + bwd
    x = (torch.randn(1, M, 4096, device=dev) * 0.5).bfloat16().requires_grad_(True)
    wg = (torch.randn(14336, 4096, device=dev) * 0.02).bfloat16().requires_grad_(True)
    wu = (torch.randn(14336, 4096, device=dev) * 0.02).bfloat16().requires_grad_(True)
    wd = (torch.randn(4096, 14336, device=dev) * 0.02).bfloat16().requires_grad_(True)
    for _ in range(2):
        y = ops.swiglu_mlp(x, wg, wu, wd, None)
        y.backward(torch.randn_like(y))
elif which == "cublas":                                  # the incumbent at the same shape, for a side-by-side ncu capture
    x = torch.randn(M, 4096, device=dev).bfloat16(); w = (torch.randn(14336, 4096, device=dev) * 0.02).bfloat16()
    for _ in range(4):
        torch.matmul(x, w.t())
elif which == "attn":
    q = torch.randn(1, M, 32, 128, device=dev).bfloat16(); k = torch.randn(1, M, 8, 128, device=dev).bfloat16()
    v = torch.randn(1, M, 8, 128, device=dev).bfloat16()
    for _ in range(3):
        o, lse = ops.attention_fwd(q, k, v, True, None, 128 ** -0.5)
        ops.attention_bwd(q, k, v, o, torch.randn_like(o), lse, True, None, 128 ** -0.5, fast=True)
elif which in ("decode1", "decode16"):                   # decode-step kernels at bs 1 / 16 (bench_kernels.py decode shapes)
    B = 1 if which == "decode1" else 16
    x = torch.randn(B, 4096, device=dev).bfloat16(); xi = torch.randn(B, 14336, device=dev).bfloat16()
    wq = (torch.randn(4096, 4096, device=dev) * 0.02).bfloat16(); wk = (torch.randn(1024, 4096, device=dev) * 0.02).bfloat16()
    wv = (torch.randn(1024, 4096, device=dev) * 0.02).bfloat16()
    wg = (torch.randn(14336, 4096, device=dev) * 0.02).bfloat16(); wu = (torch.randn(14336, 4096, device=dev) * 0.02).bfloat16()
    wd = (torch.randn(4096, 14336, device=dev) * 0.02).bfloat16()
    q = torch.empty(B, 4096, device=dev, dtype=torch.bfloat16); k = torch.empty(B, 1024, device=dev, dtype=torch.bfloat16)
    v = torch.empty_like(k); act = torch.empty(B, 14336, device=dev, dtype=torch.bfloat16)
    ctx = 6200
    qd = torch.randn(B, 1, 32, 128, device=dev).bfloat16()
    kc = torch.randn(B, ctx, 8, 128, device=dev).bfloat16(); vc = torch.randn(B, ctx, 8, 128, device=dev).bfloat16()
    for _ in range(2):
        ops._call("mb200_skinny_gemm3_bf16", ops._p(x), ops._p(wq), ops._p(wk), ops._p(wv), ops._p(q), ops._p(k), ops._p(v), B, 4096,
                  1024, 1024, 4096, 4096, 4096, ops._st())
        ops._call("mb200_skinny_swiglu_bf16", ops._p(x), ops._p(wg), ops._p(wu), ops._p(act), B, 14336, 4096, 4096, 4096, 14336, ops._st())
        ops._call("mb200_skinny_gemm_bf16", ops._p(xi), ops._p(wd), ops._p(q), None, ops._p(q), B, 4096, 14336, 14336, 14336, 4096,
                  4096, ops._st())
        ops.decode_attention(qd, kc, vc, ctx, None, 128 ** -0.5)
elif which == "merge":
    B, T, P, D = 4, 2048, 728, 4096
    ids = torch.randint(0, 128000, (B, T), device=dev)
    for j in range(8):
        ids[:, j * 256 + 16] = 128256
    emb = torch.randn(B, T, D, device=dev).bfloat16(); feats = torch.randn(32, P, D, device=dev).bfloat16()
    for _ in range(4):
        ops.merge_input_ids_with_image_features(feats, emb, ids, torch.ones_like(ids), ids, 128256, 128257)
elif which == "rmsnorm":
    x = torch.randn(M, 4096, device=dev).bfloat16().requires_grad_(True); w = torch.ones(4096, device=dev).bfloat16().requires_grad_(True)
    for _ in range(3):
        y = ops.rms_norm(x, w, 1e-5); y.backward(torch.randn_like(y))
elif which == "ce":
    n, V = 4096, 128258
    ld = (V + 7) // 8 * 8
    lg = torch.randn(n, ld, device=dev).bfloat16(); lab = torch.randint(0, V, (n,), device=dev)
    lr = torch.empty(n, device=dev); inv = torch.ones(1, device=dev)
    for _ in range(3):
        ops._call("mb200_ce_fwd_bwd", ops._p(lg), ops._p(lab), ops._p(lr), None, ops._p(lg), n, V, ld, ops._p(inv), 1.0, 1, ops._st())
torch.cuda.synchronize()
