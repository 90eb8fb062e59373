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
