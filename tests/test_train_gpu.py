"""GPU tests of the training engine: fused wgrad accumulation == autograd accumulation; a few AdamW steps reduce the loss."""
import pytest
import torch

from helpers import load_fixture, load_model, rel_err

pytestmark = pytest.mark.gpu


def _batch(fx, cuda, dtype):
    return dict(input_ids=fx["input_ids"].to(cuda), pixel_values=fx["pixel_values"].to(cuda).to(dtype),
                attention_mask=fx["attention_mask"].to(cuda), labels=fx["labels"].to(cuda))


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_fused_wgrad_accumulation_matches_autograd(cuda, dtype):
    from mantis_b200.train import B200Trainer
    fx = load_fixture("llava_siglip_full.pt")
    grads = []
    for fused in (False, True):
        model = load_model(fx, dtype, cuda).train()
        tr = B200Trainer(model, grad_accum=2, fused_wgrad_accum=fused)
        b = _batch(fx, cuda, dtype)
        tr.micro_step(b); tr.micro_step(b)
        grads.append(tr.flat_grad.clone())
    tol = 1e-5 if dtype == torch.float32 else 2e-2
    assert grads[0].dtype == torch.float32                    # the main gradient accumulates in fp32 whatever the model dtype
    assert rel_err(grads[1], grads[0]) < tol
    assert grads[0].abs().sum() > 0


def test_bf16_training_at_reference_lr_moves_the_weights(cuda):
    """bf16 model, lr 1e-5 (the reference recipe): after 10 steps through B200Trainer every trainable tensor's fp32 master has
    moved, the bf16 weights are the rounding of the masters, and the optimizer checkpoint carries the masters."""
    from mantis_b200.train import B200Trainer
    fx = load_fixture("llava_siglip_full.pt")
    model = load_model(fx, torch.bfloat16, cuda).train()
    tr = B200Trainer(model, lr=1e-5, grad_accum=1, max_grad_norm=1.0)
    before = [tr.state.master(i).clone() for i in range(len(tr.params))]
    b = _batch(fx, cuda, torch.bfloat16)
    for _ in range(10):
        loss = tr.train_step([b])
    assert torch.isfinite(loss)
    moved = []
    for i, p in enumerate(tr.params):
        now = tr.state.master(i)
        got_grad = tr.state.view(tr.state.M, i) != 0          # e.g. embedding rows of tokens that never occur receive none
        assert got_grad.any(), i
        moved.append((now != before[i])[got_grad].float().mean().item())
        assert torch.equal(now[~got_grad], before[i][~got_grad])
        assert torch.equal(p.detach().view(torch.int16), tr.state.view(tr.state.P, i).view(torch.int16))
        # bf16 weight == round-to-nearest of its master (exact ties aside)
        assert (p.detach().view(torch.int16) == now.bfloat16().view(torch.int16)).float().mean().item() > 0.999
    assert min(moved) > 0.99, moved
    # the main gradient is never zero-filled: after a step every parameter is marked "fresh" and the next gradient overwrites
    assert all(p._b200_grad_fresh for p in tr.params)
    g_before = tr.flat_grad.clone()
    tr.micro_step(b)
    assert not any(p._b200_grad_fresh for p in tr.params)
    tr2_model = load_model(fx, torch.bfloat16, cuda).train()
    tr2_model.load_state_dict(model.state_dict())
    tr2 = B200Trainer(tr2_model, lr=1e-5, grad_accum=1, max_grad_norm=1.0)
    tr2.micro_step(b)                                            # same weights, gradient buffer that started from zeros
    assert torch.equal(tr.flat_grad, tr2.flat_grad) and not torch.equal(tr.flat_grad, g_before)
    sd = tr.state_dict()
    assert all(m.dtype == torch.float32 for m in sd["master_params"])
    assert 0 < tr.grad_norm() < 1e4


def test_training_reduces_loss(cuda):
    from mantis_b200.train import B200Trainer
    fx = load_fixture("llava_siglip_full.pt")
    model = load_model(fx, torch.float32, cuda).train()
    tr = B200Trainer(model, lr=2e-3, grad_accum=1, max_grad_norm=1.0)
    b = _batch(fx, cuda, torch.float32)
    losses = [tr.train_step([b]).item() for _ in range(8)]
    assert losses[-1] < losses[0] - 0.5, losses
    for n, p in model.named_parameters():
        if "vision_tower" in n:
            assert not p.requires_grad
