"""Idefics2ForConditionalGeneration on the CUDA kernels vs golden outputs of the unmodified reference (CPU fp32)."""
import pytest
import torch

from helpers import load_fixture, rel_err

pytestmark = pytest.mark.gpu


def _build(fx, dtype, device):
    from transformers import Idefics2Config
    from mantis_b200.models.idefics2 import Idefics2ForConditionalGeneration
    model = Idefics2ForConditionalGeneration(Idefics2Config(**fx["cfg"]))
    missing, unexpected = model.load_state_dict(fx["state_dict"], strict=False)
    assert not missing and not unexpected, (missing[:4], unexpected[:4])
    return model.to(device=device, dtype=dtype)


@pytest.mark.parametrize("name", ["idefics2_full.pt", "idefics2_ragged.pt"])
@pytest.mark.parametrize("fused", [False, True])
def test_idefics2_fp32(cuda, name, fused):
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


@pytest.mark.parametrize("name", ["idefics2_full.pt", "idefics2_ragged.pt"])
def test_idefics2_bf16(cuda, name):
    fx = load_fixture(name)
    model = _build(fx, torch.bfloat16, cuda).eval()
    inputs = {k: (v.to(cuda) if torch.is_tensor(v) else v) for k, v in fx["inputs"].items()}
    inputs["pixel_values"] = inputs["pixel_values"].bfloat16()
    with torch.no_grad():
        out = model(**inputs)
    valid = fx["inputs"]["attention_mask"].bool()
    assert rel_err(out.logits.float().cpu()[valid], fx["logits"][valid]) <= 3e-2
    assert abs(out.loss.item() - fx["loss"].item()) <= 3e-2


def test_idefics2_generate_caches_image_states(cuda):
    fx = load_fixture("idefics2_full.pt")
    model = _build(fx, torch.float32, cuda).eval()
    ids = fx["inputs"]["input_ids"].to(cuda); pv = fx["inputs"]["pixel_values"].to(cuda)
    out = model.generate(input_ids=ids, attention_mask=torch.ones_like(ids), pixel_values=pv, max_new_tokens=6,
                         do_sample=False, num_beams=1)
    # cache-free greedy loop over our own forward must agree token for token
    seq = ids.clone()
    with torch.no_grad():
        for _ in range(6):
            lg = model(input_ids=seq, attention_mask=torch.ones_like(seq), pixel_values=pv).logits
            seq = torch.cat([seq, lg[:, -1].argmax(-1, keepdim=True)], 1)
    assert out.cpu().tolist() == seq.cpu().tolist()


def test_idefics2_decode_uses_native_engine(cuda):
    """bf16, head_dim 128: single-token steps of Idefics2 run through the C++ decode engine (one call per token) and
    reproduce the Python-path logits step after step"""
    from transformers import Idefics2Config
    import mantis_b200.ops as om
    from mantis_b200.models.idefics2 import Idefics2ForConditionalGeneration
    from mantis_b200.models.kv_cache import B200KVCache
    cfg = Idefics2Config(
        vision_config=dict(hidden_size=64, intermediate_size=128, num_hidden_layers=2, num_attention_heads=4, image_size=112,
                           patch_size=14),
        perceiver_config=dict(resampler_n_latents=8, resampler_depth=1, resampler_n_heads=2, resampler_head_dim=32,
                              num_key_value_heads=1, hidden_size=256, rms_norm_eps=1e-5),
        text_config=dict(model_type="mistral", hidden_size=256, intermediate_size=512, num_hidden_layers=2,
                         num_attention_heads=2, num_key_value_heads=1, vocab_size=1000, pad_token_id=0, rms_norm_eps=1e-5,
                         sliding_window=4096),
        image_token_id=990, tie_word_embeddings=False)
    torch.manual_seed(0)
    model = Idefics2ForConditionalGeneration(cfg).to(cuda).to(torch.bfloat16).eval()
    with torch.no_grad():
        for p in model.parameters():
            if p.dim() >= 2:
                p.mul_(3.0)
    ids = torch.randint(1, 980, (2, 24), device=cuda); ids[:, 3:11] = 990
    pv = torch.randn(2, 1, 3, 112, 112, device=cuda).bfloat16()

    def run(native):
        old = om.FORCE_GENERIC
        logits = []
        with torch.no_grad():
            cache = B200KVCache()
            am = torch.ones_like(ids)
            out = model(input_ids=ids, pixel_values=pv, attention_mask=am, past_key_values=cache, use_cache=True)
            for step in range(4):
                nxt = torch.tensor([7, 11], device=cuda) + step
                am = torch.cat([am, torch.ones_like(nxt[:, None])], 1)
                om.FORCE_GENERIC = not native
                try:
                    out = model(input_ids=nxt[:, None], attention_mask=am, past_key_values=cache, use_cache=True,
                                image_hidden_states=out.image_hidden_states)
                finally:
                    om.FORCE_GENERIC = old
                logits.append(out.logits[:, -1].float().clone())
        return logits, hasattr(cache, "_engine")

    a, used_a = run(True)
    b, used_b = run(False)
    assert used_a and not used_b
    for x, y in zip(a, b):
        assert rel_err(x, y) < 3e-2, rel_err(x, y)


def test_idefics2_sliding_window_shorter_than_the_sequence(cuda):
    """Mistral sliding window (12) < S (40): forward/backward through the window-aware attention and a cached greedy decode that
    keeps sliding, against the unmodified reference (fixture idefics2_sliding.pt: transformers' MistralModel applies
    kv_idx > q_idx - sliding_window)"""
    fx = load_fixture("idefics2_sliding.pt")
    assert fx["cfg"]["text_config"]["sliding_window"] == 12
    model = _build(fx, torch.float32, cuda).train()
    model.materialize_logits_in_training = True
    inputs = {k: (v.to(cuda) if torch.is_tensor(v) else v) for k, v in fx["inputs"].items()}
    out = model(**inputs)
    err = (out.logits.detach().cpu() - fx["logits"]).abs().max().item()
    assert err <= 1e-3 * fx["logits"].abs().max().item(), err
    assert abs(out.loss.item() - fx["loss"].item()) <= 1e-4 * max(1.0, abs(fx["loss"].item()))
    out.loss.backward()
    params = dict(model.named_parameters())
    for k, g in fx["grads"].items():
        assert rel_err(params[k].grad, g) <= 2e-3, (k, rel_err(params[k].grad, g))
    # without the window the logits are measurably different (the fixture really exercises it)
    model.eval()
    for layer in model.model.text_model.layers:
        layer.self_attn.sliding_window = None
    with torch.no_grad():
        nowin = model(**{k: v for k, v in inputs.items() if k != "labels"}).logits
    assert (nowin.cpu() - fx["logits"]).abs().max().item() > 50 * max(err, 1e-6)
    for layer in model.model.text_model.layers:
        layer.self_attn.sliding_window = 12
    ids = inputs["input_ids"]; n_new = fx["generated"].shape[1] - ids.shape[1]
    gen = model.generate(input_ids=ids, attention_mask=torch.ones_like(ids), pixel_values=inputs["pixel_values"],
                         max_new_tokens=n_new, do_sample=False, num_beams=1)
    assert gen.cpu().tolist() == fx["generated"].tolist()
