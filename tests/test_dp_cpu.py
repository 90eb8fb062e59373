"""World-size-2 gloo test (CPU) of the data-parallel plumbing: flat gradient buffer, autograd accumulating into its views
across micro-batches, and the single all-reduce.  (The AdamW update itself is a CUDA kernel and is tested on the GPU.)"""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from mantis_b200.train.engine import B200Trainer
    torch.manual_seed(0)
    model = torch.nn.Sequential(torch.nn.Linear(16, 32), torch.nn.GELU(), torch.nn.Linear(32, 4))
    # plumbing only (no CUDA kernels on this box): flat fp32 state, gradients folded into the flat main gradient by hooks
    tr = B200Trainer(model, grad_accum=2, freeze_vision=False, fused_wgrad_accum=False, overlap_allreduce=False)
    assert tr.world == world
    g = torch.Generator().manual_seed(100 + rank)
    xs = [torch.randn(8, 16, generator=g) for _ in range(2)]
    for x in xs:                                     # two micro-batches accumulate into the flat buffer
        (model(x).pow(2).mean() / tr.grad_accum).backward()
    assert all(p.grad is None for p in tr.params)                      # folded into the flat buffer and released
    assert tr.params[0]._b200_unfused_main_grad.data_ptr() == tr.flat_grad.data_ptr()
    local = tr.flat_grad.clone()
    scale = tr.reduce_gradients()
    torch.save((rank, local, tr.flat_grad.clone() * scale), os.path.join(q, f"r{rank}.pt"))
    dist.destroy_process_group()


def test_flat_buffer_allreduce_world2():
    import tempfile
    ctx = mp.get_context("spawn")
    q = tempfile.mkdtemp()                     # results come back through files (robust against queue teardown races)
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=180)
        assert p.exitcode == 0
    res = sorted([torch.load(os.path.join(q, f"r{r}.pt")) for r in range(2)], key=lambda t: t[0])
    mean = (res[0][1] + res[1][1]) / 2
    assert torch.allclose(res[0][2], mean, atol=1e-6) and torch.allclose(res[1][2], mean, atol=1e-6)
    assert not torch.allclose(res[0][1], res[1][1])          # ranks really saw different data


def test_flat_state_views_are_aligned():
    from mantis_b200.train.engine import FlatState
    ps = [torch.nn.Parameter(torch.randn(n)) for n in (3, 17, 2064, 5)]
    vals = [p.detach().clone() for p in ps]
    st = FlatState(ps)
    assert st.total == 1024 * 6
    for p, v, o in zip(ps, vals, st.offsets):
        assert o % 1024 == 0 and p.data_ptr() == st.P.data_ptr() + 4 * o
        assert torch.equal(p.detach(), v) and p._b200_main_grad.shape == p.shape


def _overlap_worker(rank, world, port, q):
    import torch.nn as nn
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from mantis_b200.train.engine import B200Trainer

    class Layer(nn.Module):
        def __init__(self):
            super().__init__(); self.a = nn.Linear(8, 8); self.b = nn.Linear(8, 8)

        def forward(self, x):
            return x + self.b(torch.tanh(self.a(x)))

    class Net(nn.Module):
        def __init__(self):
            super().__init__()
            self.language_model = nn.Module(); self.language_model.model = nn.Module()
            self.language_model.model.embed = nn.Linear(8, 8)
            self.language_model.model.layers = nn.ModuleList([Layer() for _ in range(4)])
            self.language_model.head = nn.Linear(8, 3)

        def forward(self, x):
            x = self.language_model.model.embed(x)
            for l in self.language_model.model.layers:
                x = l(x)
            return self.language_model.head(x)

    torch.manual_seed(0)
    model = Net()
    # MLlava registers trainable modules AFTER language_model (image_type_embeddings, vision_xatten_layers): their gradients
    # are only final at the end of backward, so they must not sit in the tail slice the first overlapped all-reduce takes
    model.image_type_embeddings = nn.Embedding(4, 8)
    inner = model.forward
    model.forward = lambda x: inner(x + model.image_type_embeddings.weight[:1])
    tr = B200Trainer(model, grad_accum=2, freeze_vision=False, fused_wgrad_accum=False, overlap_allreduce=True)
    assert getattr(tr, "_has_hooks", False)
    idx = lambda t: [i for i, p in enumerate(tr.params) if p is t][0]
    late = idx(model.image_type_embeddings.weight)
    assert tr.state.offsets[late] < tr._layer_starts[0], "late parameters must precede the decoder layers in the flat buffer"
    g = torch.Generator().manual_seed(10 + rank)
    xs = [torch.randn(5, 8, generator=g) for _ in range(2)]
    # reference: plain accumulation, then one all-reduce
    for x in xs:
        (model(x).pow(2).mean() / 2).backward()
    ref = tr.flat_grad.clone(); dist.all_reduce(ref); ref /= world
    tr.flat_grad.zero_()
    # overlapped path: hooks active only during the last micro-batch
    for i, x in enumerate(xs):
        if i == 1:
            tr._overlap = True; tr._reduced_from = tr.flat_grad.numel()
        (model(x).pow(2).mean() / 2).backward()
        tr._overlap = False
    n_async = len(tr._works)
    scale = tr.reduce_gradients()
    ok = bool(torch.allclose(tr.flat_grad * scale, ref, atol=1e-6)) and ref[tr.state.offsets[late]:].abs().sum() > 0
    torch.save((rank, n_async, ok), os.path.join(q, f"r{rank}.pt"))
    dist.destroy_process_group()


def test_overlapped_allreduce_matches_single_allreduce():
    import tempfile
    ctx = mp.get_context("spawn")
    q = tempfile.mkdtemp()
    port = _free_port()
    procs = [ctx.Process(target=_overlap_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=180)
        assert p.exitcode == 0
    res = [torch.load(os.path.join(q, f"r{r}.pt")) for r in range(2)]
    for rank, n_async, ok in res:
        assert ok and n_async >= 3
