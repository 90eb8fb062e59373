"""LlavaNextForConditionalGeneration on the CUDA kernels vs the golden output of the unmodified reference
(mantis/models/mllava_next/modeling_llava_next.py, CPU fp32; fixture from oracle/make_golden_llava_next.py)."""
import pytest
import torch

from helpers import load_fixture, rel_err

pytestmark = pytest.mark.gpu


def _build(fx, dtype, device):
    from transformers import CLIPVisionConfig, LlamaConfig
    from mantis_b200.models.mllava_next import LlavaNextConfig, LlavaNextForConditionalGeneration
    cfg = LlavaNextConfig(vision_config=CLIPVisionConfig(**fx["vision"]), text_config=LlamaConfig(**fx["text"]), **fx["cfg"])
    model = LlavaNextForConditionalGeneration(cfg)
    missing, unexpected = model.load_state_dict(fx["state_dict"], strict=False)
    assert not missing and not unexpected, (missing[:4], unexpected[:4])
    return model.to(device=device, dtype=dtype)


def _inputs(fx, device, dtype=torch.float32):
    return dict(input_ids=fx["input_ids"].to(device), attention_mask=fx["attention_mask"].to(device),
                labels=fx["labels"].to(device), pixel_values=[p.to(device=device, dtype=dtype) for p in fx["pixel_values"]],
                image_sizes=fx["image_sizes"].to(device))


@pytest.mark.parametrize("fused", [False, True])
def test_llava_next_fp32(cuda, fused):
    fx = load_fixture("llava_next_batch.pt")
    model = _build(fx, torch.float32, cuda).train()
    model.materialize_logits_in_training = not fused
    out = model(**_inputs(fx, cuda))
    if not fused:
        assert out.logits.shape == fx["logits"].shape
        # compare where the merged attention mask is 1: every row of sample 1; sample 0 = 17 tokens of which one is an image
        # placeholder -> 16 + 65 = 81 rows, then the pad rows and the zero fill
        valid = torch.ones(fx["logits"].shape[:2], dtype=torch.bool)
        valid[0, 81:] = False
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


def test_llava_next_only_base_crop_matters_and_tensor_input(cuda):
    """the reference discards every crop but the first: changing the others must not change anything, and a 5-D tensor of
    equally deep stacks is accepted like the list form"""
    fx = load_fixture("llava_next_batch.pt")
    model = _build(fx, torch.float32, cuda).eval()
    inp = _inputs(fx, cuda)
    with torch.no_grad():
        a = model(**inp)
        inp2 = dict(inp)
        inp2["pixel_values"] = torch.stack([torch.cat([p[:1], torch.randn(1, *p.shape[1:], device=cuda)]) for p in inp["pixel_values"]])
        b = model(**inp2)
    assert torch.equal(a.logits, b.logits) and torch.equal(a.loss, b.loss)
    assert abs(a.loss.item() - fx["loss"].item()) <= 1e-4


def test_llava_next_bf16_and_generate(cuda):
    fx = load_fixture("llava_next_batch.pt")
    model = _build(fx, torch.bfloat16, cuda).eval()
    inp = _inputs(fx, cuda, torch.bfloat16)
    with torch.no_grad():
        out = model(**inp)
    assert abs(out.loss.item() - fx["loss"].item()) <= 3e-2
    # cached greedy decoding == cache-free re-forward, token for token (fp32 to keep argmax ties out of the picture)
    model = _build(fx, torch.float32, cuda).eval()
    ids = fx["input_ids"][1:2].to(cuda); pv = [p.to(cuda) for p in fx["pixel_values"][1:]]
    gen = model.generate(input_ids=ids, attention_mask=torch.ones_like(ids), pixel_values=pv,
                         image_sizes=fx["image_sizes"][1:].to(cuda), max_new_tokens=5, do_sample=False, num_beams=1)
    seq = ids.clone()
    with torch.no_grad():
        for _ in range(5):
            lg = model(input_ids=seq, attention_mask=torch.ones_like(seq), pixel_values=pv).logits
            seq = torch.cat([seq, lg[:, -1].argmax(-1, keepdim=True)], 1)
    assert gen.cpu().tolist() == seq.cpu().tolist()
