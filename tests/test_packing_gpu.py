"""Sequence packing (ref: mantis/train/data.py:1546-1671): block-diagonal causal attention over packed rows.
The kernel-level oracle is a dense fp32 softmax with the reference's block-diagonal mask; the model-level property is
packed forward/backward == the samples run one by one."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _rel(a, b):
    a = a.float(); b = b.float()
    return ((a - b).norm() / (b.norm() + 1e-12)).item()


def _dense_block_ref(q, k, v, segs, kmask, scale):
    B, S, H, hd = q.shape
    G = H // k.shape[2]
    kk = k.repeat_interleave(G, dim=2); vv = v.repeat_interleave(G, dim=2)
    s = torch.einsum("bqhd,bkhd->bhqk", q, kk) * scale
    allow = torch.zeros(B, S, S, dtype=torch.bool, device=q.device)
    for (b, s0, s1) in segs:
        blk = torch.ones(s1 - s0, s1 - s0, dtype=torch.bool, device=q.device).tril()
        if kmask is not None:
            blk = blk & (kmask[b, s0:s1] != 0)[None, :]
        allow[b, s0:s1, s0:s1] = blk
    s = s.masked_fill(~allow[:, None], float("-inf"))
    p = torch.softmax(s, -1).nan_to_num(0.0)                  # rows with no visible key -> 0, like the kernels
    return torch.einsum("bhqk,bkhd->bqhd", p, vv)


@pytest.mark.parametrize("dtype,hd,H,Hkv,lens", [
    (torch.bfloat16, 128, 8, 2, [[300, 129, 70, 513]]),                 # tcgen05 kernels (Sq >= 64)
    (torch.bfloat16, 128, 4, 4, [[40, 200], [130, 110]]),                # short segment -> generic kernel, B = 2
    (torch.float32, 16, 4, 2, [[5, 17, 1, 30]]),
])
def test_varlen_attention_matches_dense_block_mask(ops, cuda, dtype, hd, H, Hkv, lens):
    torch.manual_seed(40)
    B, S = len(lens), sum(lens[0])
    assert all(sum(l) == S for l in lens)
    segs = []
    for b, l in enumerate(lens):
        acc = 0
        for n in l:
            segs.append((b, acc, acc + n)); acc += n
    q = torch.randn(B, S, H, hd, device=cuda).to(dtype).requires_grad_(True)
    k = torch.randn(B, S, Hkv, hd, device=cuda).to(dtype).requires_grad_(True)
    v = torch.randn(B, S, Hkv, hd, device=cuda).to(dtype).requires_grad_(True)
    kmask = torch.ones(B, S, dtype=torch.int64, device=cuda)
    kmask[0, 3] = 0; kmask[B - 1, S - 2] = 0
    scale = hd ** -0.5
    for km in (None, kmask):
        o = ops.attention_varlen(q, k, v, segs, kmask=km, scale=scale)
        go = torch.randn_like(o)
        dq, dk, dv = torch.autograd.grad(o, (q, k, v), go)
        qr, kr, vr = (t.detach().float().requires_grad_(True) for t in (q, k, v))
        ref = _dense_block_ref(qr, kr, vr, segs, km, scale)
        rq, rk, rv = torch.autograd.grad(ref, (qr, kr, vr), go.float())
        tol = 2e-5 if dtype == torch.float32 else 2e-2
        assert _rel(o, ref) < tol, _rel(o, ref)
        assert _rel(dq, rq) < tol and _rel(dk, rk) < tol and _rel(dv, rv) < tol, (_rel(dq, rq), _rel(dk, rk), _rel(dv, rv))


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_packed_forward_backward_equals_separate_samples(cuda, dtype):
    """LLaMA decoder + LM head: one packed row (4-D block-diagonal mask + restarting position ids, as PackingDataset emits
    them) gives the same logits and the same summed-loss gradients as the samples run one at a time."""
    from transformers import LlamaConfig
    from mantis_b200.models.llama import B200CausalLM
    from mantis_b200.train import PackingDataset
    torch.manual_seed(41)
    hd = 128 if dtype == torch.bfloat16 else 16
    cfg = LlamaConfig(hidden_size=4 * hd, intermediate_size=256, num_hidden_layers=2, num_attention_heads=4,
                      num_key_value_heads=2, vocab_size=500, rms_norm_eps=1e-5, rope_theta=10000.0, head_dim=hd)
    model = B200CausalLM(cfg).to(cuda).to(dtype)
    lens = [90, 140, 33, 200]
    items = []
    for i, n in enumerate(lens):
        ids = torch.randint(0, 500, (1, n))
        items.append({"input_ids": ids, "attention_mask": torch.ones(1, n, dtype=torch.long), "labels": ids.clone(),
                      "pixel_values": None})
    packed = PackingDataset(items, max_self_attn_len=10_000).pack_batch(items)
    assert packed["attention_mask"].shape == (1, 1, sum(lens), sum(lens))
    out = model(input_ids=packed["input_ids"].to(cuda), attention_mask=packed["attention_mask"].to(cuda),
                position_ids=packed["position_ids"].to(cuda))
    # the same row through the sync-free entry (segment table from the dataset, 2-D key mask)
    out2 = model(input_ids=packed["input_ids"].to(cuda), position_ids=packed["position_ids"].to(cuda),
                 cu_segments=packed["cu_segments"])
    assert torch.equal(out.logits, out2.logits)
    sep = torch.cat([model(input_ids=it["input_ids"].to(cuda)).logits for it in items], dim=1)
    tol = 1e-4 if dtype == torch.float32 else 2e-2
    assert _rel(out.logits, sep) < tol, _rel(out.logits, sep)
    # gradients of sum-of-logits-weighted loss
    w = torch.randn_like(out.logits.float())
    model.zero_grad(); (out.logits.float() * w).sum().backward()
    g_packed = {n: p.grad.detach().float().clone() for n, p in model.named_parameters()}
    model.zero_grad()
    acc = 0
    for it in items:
        n = it["input_ids"].shape[1]
        lg = model(input_ids=it["input_ids"].to(cuda)).logits.float()
        (lg * w[:, acc:acc + n]).sum().backward()
        acc += n
    for n_, p in model.named_parameters():
        assert _rel(g_packed[n_], p.grad) < (1e-3 if dtype == torch.float32 else 5e-2), (n_, _rel(g_packed[n_], p.grad))
