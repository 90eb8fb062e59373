"""Idefics3ForConditionalGeneration on the CUDA kernels vs golden outputs of the unmodified reference
(mantis/models/idefics3/modeling_idefics3.py, CPU fp32; fixtures from oracle/make_golden_idefics3.py)."""
import pytest
import torch

from helpers import load_fixture, rel_err

pytestmark = pytest.mark.gpu


def _build(fx, dtype, device):
    from mantis_b200.models.idefics3 import Idefics3Config, Idefics3ForConditionalGeneration
    model = Idefics3ForConditionalGeneration(Idefics3Config(**fx["cfg"]))
    missing, unexpected = model.load_state_dict(fx["state_dict"], strict=False)
    assert not missing and not unexpected, (missing[:4], unexpected[:4])
    return model.to(device=device, dtype=dtype)


@pytest.mark.parametrize("name", ["idefics3_full.pt", "idefics3_ragged.pt"])
@pytest.mark.parametrize("fused", [False, True])
def test_idefics3_fp32(cuda, name, fused):
    fx = load_fixture(name)
    model = _build(fx, torch.float32, cuda).train()
    model.materialize_logits_in_training = not fused
    inputs = {k: (v.to(cuda) if torch.is_tensor(v) else v) for k, v in fx["inputs"].items()}
    out = model(**inputs)
    if not fused:
        assert out.logits.dtype == torch.float32 and out.logits.shape == fx["logits"].shape
        valid = fx["inputs"]["attention_mask"].bool()
        err = (out.logits.detach().cpu() - fx["logits"]).abs().amax(-1)[valid].max().item()
        assert err <= 1e-3 * fx["logits"].abs().max().item(), err
    else:
        assert out.logits is None
    assert abs(out.loss.item() - fx["loss"].item()) <= 1e-4
    out.loss.backward()
    params = dict(model.named_parameters())
    for k, g in fx["grads"].items():
        assert params[k].grad is not None, k
        assert rel_err(params[k].grad, g) <= 2e-3, (k, rel_err(params[k].grad, g))


@pytest.mark.parametrize("name", ["idefics3_full.pt", "idefics3_ragged.pt"])
def test_idefics3_bf16(cuda, name):
    fx = load_fixture(name)
    model = _build(fx, torch.bfloat16, cuda).eval()
    inputs = {k: (v.to(cuda) if torch.is_tensor(v) else v) for k, v in fx["inputs"].items()}
    inputs["pixel_values"] = inputs["pixel_values"].bfloat16()
    with torch.no_grad():
        out = model(**inputs)
    valid = fx["inputs"]["attention_mask"].bool()
    assert rel_err(out.logits.float().cpu()[valid], fx["logits"][valid]) <= 3e-2
    assert abs(out.loss.item() - fx["loss"].item()) <= 3e-2


def test_idefics3_generate_caches_image_states(cuda):
    fx = load_fixture("idefics3_full.pt")
    model = _build(fx, torch.float32, cuda).eval()
    ids = fx["inputs"]["input_ids"].to(cuda); pv = fx["inputs"]["pixel_values"].to(cuda)
    out = model.generate(input_ids=ids, attention_mask=torch.ones_like(ids), pixel_values=pv, max_new_tokens=6,
                         do_sample=False, num_beams=1)
    seq = ids.clone()
    with torch.no_grad():
        for _ in range(6):
            lg = model(input_ids=seq, attention_mask=torch.ones_like(seq), pixel_values=pv).logits
            seq = torch.cat([seq, lg[:, -1].argmax(-1, keepdim=True)], 1)
    assert out.cpu().tolist() == seq.cpu().tolist()
