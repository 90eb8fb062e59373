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
    assert rel_err(grads[1], grads[0]) < tol
    assert grads[0].abs().sum() > 0


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
