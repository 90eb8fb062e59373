"""Kernel parity at the EXACT shapes bench.py runs (BASELINE config 2: S = 7864 merged tokens, LLaMA-3-8B widths, vocab
128,258) against fp32 torch math -- the small-shape tests in test_kernels_gpu.py stop at S = 1000 / N = 4096."""
import pytest
import torch

pytestmark = pytest.mark.gpu

S_BENCH = 7864


def _rel(a, b):
    a = a.float(); b = b.float()
    return ((a - b).norm() / (b.norm() + 1e-12)).item()


def _exact():
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.set_float32_matmul_precision("highest")


def _attn_ref_heads(q, k, v, scale, go=None):
    """causal GQA attention in fp32, one head at a time (scores of one head at S = 7864 are 247 MB).
    q [S,H,hd], k/v [S,Hkv,hd] (bf16) -> o [S,H,hd] fp32 and, if go is given, (dq, dk, dv) fp32."""
    S, H, hd = q.shape
    Hkv = k.shape[1]
    rep = H // Hkv
    o = torch.empty((S, H, hd), dtype=torch.float32, device=q.device)
    dq = torch.zeros_like(o) if go is not None else None
    dk = torch.zeros((S, Hkv, hd), dtype=torch.float32, device=q.device) if go is not None else None
    dv = torch.zeros_like(dk) if go is not None else None
    tri = torch.ones((S, S), dtype=torch.bool, device=q.device).tril_()
    for h in range(H):
        g = h // rep
        qh, kh, vh = q[:, h].float(), k[:, g].float(), v[:, g].float()
        s = (qh @ kh.t()) * scale
        s.masked_fill_(~tri, float("-inf"))
        p = torch.softmax(s, dim=-1)
        del s
        oh = p @ vh
        o[:, h] = oh
        if go is not None:
            gh = go[:, h].float()
            dv[:, g] += p.t() @ gh
            dp = gh @ vh.t()
            delta = (gh * oh).sum(-1, keepdim=True)
            ds = p * (dp - delta) * scale
            del dp
            dq[:, h] = ds @ kh
            dk[:, g] += ds.t() @ qh
            del ds
        del p
    return o, dq, dk, dv


@pytest.mark.parametrize("fwd2", [True, False])
def test_attention_fwd_bench_shape(ops, cuda, fwd2):
    import mantis_b200.ops as om
    _exact()
    torch.manual_seed(40)
    H, Hkv, hd = 32, 8, 128
    q = torch.randn(1, S_BENCH, H, hd, device=cuda).bfloat16()
    k = torch.randn(1, S_BENCH, Hkv, hd, device=cuda).bfloat16()
    v = torch.randn(1, S_BENCH, Hkv, hd, device=cuda).bfloat16()
    old = om.ATTN_FWD2; om.ATTN_FWD2 = fwd2
    try:
        o, lse = ops.attention_fwd(q, k, v, True, None, hd ** -0.5)
    finally:
        om.ATTN_FWD2 = old
    ref, _, _, _ = _attn_ref_heads(q[0], k[0], v[0], hd ** -0.5)
    assert _rel(o[0], ref) < 6e-3, _rel(o[0], ref)
    # the rows that see the most keys (62 KV tiles) are the ones a tile-count bug would break
    assert _rel(o[0, -256:], ref[-256:]) < 8e-3
    assert (o[0].float() - ref).abs().max().item() < 3e-2


def test_attention_bwd_bench_shape(ops, cuda):
    _exact()
    torch.manual_seed(41)
    H, Hkv, hd = 32, 8, 128
    q = torch.randn(1, S_BENCH, H, hd, device=cuda).bfloat16().requires_grad_(True)
    k = torch.randn(1, S_BENCH, Hkv, hd, device=cuda).bfloat16().requires_grad_(True)
    v = torch.randn(1, S_BENCH, Hkv, hd, device=cuda).bfloat16().requires_grad_(True)
    go = torch.randn(1, S_BENCH, H, hd, device=cuda).bfloat16()
    o = ops.attention(q, k, v, causal=True, kmask=None, scale=hd ** -0.5)
    o.backward(go)
    _, dq, dk, dv = _attn_ref_heads(q.detach()[0], k.detach()[0], v.detach()[0], hd ** -0.5, go[0])
    assert _rel(q.grad[0], dq) < 1.2e-2, _rel(q.grad[0], dq)
    assert _rel(k.grad[0], dk) < 1.2e-2, _rel(k.grad[0], dk)
    assert _rel(v.grad[0], dv) < 1.2e-2, _rel(v.grad[0], dv)


@pytest.mark.parametrize("what", ["fwd", "dgrad", "wgrad", "wgrad_accumulate"])
@pytest.mark.parametrize("N,K", [(14336, 4096), (4096, 14336), (4096, 4096), (1024, 4096)])
def test_gemm_bench_shapes(ops, cuda, what, N, K):
    """gate/up (N 14336), down (K 14336), q/o (4096^2), k/v (N 1024) at M = 7864 in all three operand majornesses"""
    _exact()
    torch.manual_seed(42)
    M = S_BENCH
    x = (torch.randn(M, K, device=cuda) * 0.5).bfloat16()
    w = (torch.randn(N, K, device=cuda) * 0.05).bfloat16()
    dy = (torch.randn(M, N, device=cuda) * 0.5).bfloat16()
    if what == "fwd":
        got, ref = ops.gemm(x, w), x.float() @ w.float().t()
    elif what == "dgrad":
        got, ref = ops.gemm(dy, w, trans_a=False, trans_b=False), dy.float() @ w.float()
    elif what == "wgrad":
        got, ref = ops.gemm(dy, x, trans_a=True, trans_b=False), dy.float().t() @ x.float()
    else:
        acc = (torch.randn(N, K, device=cuda) * 20).bfloat16()
        ref = acc.float() + dy.float().t() @ x.float()
        got = ops.gemm(dy, x, trans_a=True, trans_b=False, addend=acc, out=acc)
    assert _rel(got, ref) < 4e-3, _rel(got, ref)
    assert (got.float() - ref).abs().max().item() <= 1.6e-2 * ref.abs().max().item() + 1e-3


def test_lm_head_gemm_bench_shape(ops, cuda):
    """LM-head chunk: (4096 rows, 128258 vocab, 4096) forward into a padded-ld buffer, d(hidden) and dW from it"""
    _exact()
    torch.manual_seed(43)
    M, V, D = 4096, 128258, 4096
    h = (torch.randn(M, D, device=cuda) * 0.5).bfloat16()
    w = (torch.randn(V, D, device=cuda) * 0.05).bfloat16()
    ld = (V + 7) // 8 * 8
    buf = torch.empty((M, ld), dtype=torch.bfloat16, device=cuda)
    logits = buf[:, :V]
    ops.gemm(h, w, out=logits)
    ref = h.float() @ w.float().t()
    assert _rel(logits, ref) < 4e-3
    assert (logits.float() - ref).abs().max().item() <= 1.6e-2 * ref.abs().max().item()
    del ref
    dl = (torch.randn(M, V, device=cuda) * 0.01).bfloat16()
    logits.copy_(dl)
    dh = ops.gemm(logits, w, trans_a=False, trans_b=False)
    assert _rel(dh, dl.float() @ w.float()) < 4e-3
    dw = ops.gemm(logits, h, trans_a=True, trans_b=False)
    assert _rel(dw, dl.float().t() @ h.float()) < 4e-3


def test_fused_lm_head_ce_bench_shape(ops, cuda):
    """fused chunked LM-head + CE at vocab 128,258 with 74 % ignored rows (config-2 label density) vs fp32 torch"""
    _exact()
    torch.manual_seed(44)
    n, V, D = 6000, 128258, 4096
    h = (torch.randn(1, n, D, device=cuda) * 0.5).bfloat16().requires_grad_(True)
    w = (torch.randn(V, D, device=cuda) * 0.03).bfloat16().requires_grad_(True)
    lab = torch.randint(0, V, (1, n), device=cuda)
    lab[torch.rand(1, n, device=cuda) < 0.74] = -100
    count = (lab >= 0).sum().float().reshape(1)
    loss = ops.lm_head_ce(h, w, lab, count)
    loss.backward()
    hr = h.detach().float().requires_grad_(True); wr = w.detach().float().requires_grad_(True)
    keep = (lab[0] >= 0).nonzero().squeeze(1)
    lr = torch.nn.functional.cross_entropy(hr[0, keep] @ wr.t(), lab[0, keep])
    lr.backward()
    assert abs(loss.item() - lr.item()) < 2e-3 * abs(lr.item())
    assert _rel(h.grad, hr.grad) < 1.2e-2, _rel(h.grad, hr.grad)
    assert _rel(w.grad, wr.grad) < 1.2e-2, _rel(w.grad, wr.grad)
