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


@pytest.mark.parametrize("name", ["llava_siglip_full.pt", "llava_batch_pad.pt"])
def test_collator_hints_make_the_step_sync_free_with_identical_results(cuda, name):
    """train.Collator's host-side hints (merged length, padding side, supervised-row count) replace the two device read-backs of a
    micro-batch (merge plan header, LM-head row compaction size); loss and every gradient are bit-identical, and a wrong hint
    is reported by ops.check_deferred()"""
    from mantis_b200 import ops
    from mantis_b200.train import Collator
    fx = load_fixture(name)
    model = load_model(fx, torch.float32, cuda).train()
    b = _batch(fx, cuda, torch.float32)
    pad = fx["meta"]["cfg_kwargs"]["pad_token_id"]
    col = Collator(pad_token_id=pad, image_token_index=fx["meta"]["cfg_kwargs"]["image_token_index"])
    hint = col.merge_hint(fx["input_ids"], pad, fx["labels"], fx["attention_mask"])
    assert set(hint) == {"max_image_tokens", "left_padding", "valid_rows"}

    def run(**extra):
        model.zero_grad(set_to_none=True)
        out = model(**b, **extra)
        out.loss.backward()
        return out.loss.detach().clone(), {k: p.grad.detach().clone() for k, p in model.named_parameters() if p.grad is not None}

    l0, g0 = run()
    l1, g1 = run(merge_hint=hint)
    ops.check_deferred()                              # nothing to report
    assert torch.equal(l0, l1) and abs(l1.item() - fx["loss"].item()) < 1e-4
    assert g0.keys() == g1.keys()
    for k in g0:
        if "embed" in k:              # scatter-adds with repeated rows (token / position embeddings): fp32 atomics, the last
            assert torch.allclose(g0[k], g1[k], rtol=1e-5, atol=1e-7), k      # bit depends on the order of arrival
        else:
            assert torch.equal(g0[k], g1[k]), k
    bad = dict(hint, valid_rows=hint["valid_rows"] - 1)
    model.zero_grad(set_to_none=True)
    model(**b, merge_hint=bad)
    with pytest.raises(ValueError, match="valid_rows"):
        ops.check_deferred()
