"""HF-API surface on the GPU: gradient checkpointing, resize_token_embeddings, the 3-argument constructor used by the
pretrain driver (mantis/train/train_mllava.py:171), pixel_values passed as a python list (modeling_llava.py:431-432)."""
import pytest
import torch

from helpers import load_fixture, load_model, rel_err

pytestmark = pytest.mark.gpu


def _inputs(fx, cuda):
    return dict(input_ids=fx["input_ids"].to(cuda), pixel_values=fx["pixel_values"].to(cuda),
                attention_mask=fx["attention_mask"].to(cuda), labels=fx["labels"].to(cuda))


def test_gradient_checkpointing_matches(cuda):
    fx = load_fixture("llava_siglip_full.pt")
    grads = []
    for ckpt in (False, True):
        model = load_model(fx, torch.float32, cuda).train()
        if ckpt:
            model.gradient_checkpointing_enable()
        out = model(**_inputs(fx, cuda))
        out.loss.backward()
        assert abs(out.loss.item() - fx["loss"].item()) < 1e-4
        grads.append({k: p.grad.clone() for k, p in model.named_parameters() if p.grad is not None})
    for k in fx["grads"]:
        assert rel_err(grads[1][k], grads[0][k]) < 1e-5, k


def test_resize_token_embeddings_and_list_pixel_values(cuda):
    fx = load_fixture("llava_siglip_full.pt")
    model = load_model(fx, torch.float32, cuda).eval()
    emb = model.resize_token_embeddings(328)
    assert emb.num_embeddings == 328 and model.config.vocab_size == 328 and model.vocab_size == 328
    inp = _inputs(fx, cuda)
    pv = inp["pixel_values"]
    inp["pixel_values"] = [pv[:1], None, pv[1:]]                  # list of per-sample tensors (None entries skipped)
    with torch.no_grad():
        out = model(**inp)
    assert out.logits.shape[-1] == 328
    ref = fx["logits"]
    assert (out.logits[..., :320].cpu() - ref).abs().max().item() <= 1e-3 * ref.abs().max().item()


def test_three_argument_constructor(cuda):
    from mantis_b200.models.mllava import LlavaForConditionalGeneration
    fx = load_fixture("llava_siglip_full.pt")
    full = load_model(fx, torch.float32, cuda)
    model = LlavaForConditionalGeneration(full.config, full.vision_tower, full.language_model).to(cuda)
    model.multi_modal_projector.load_state_dict(full.multi_modal_projector.state_dict())
    model.eval()
    with torch.no_grad():
        out = model(**_inputs(fx, cuda))
    assert abs(out.loss.item() - fx["loss"].item()) < 1e-4
    assert model.language_model is full.language_model


def test_value_error_on_image_count_mismatch(cuda):
    fx = load_fixture("llava_siglip_full.pt")
    model = load_model(fx, torch.float32, cuda).eval()
    inp = _inputs(fx, cuda)
    inp["pixel_values"] = inp["pixel_values"][:1]                 # 2 placeholders, 1 image
    with pytest.raises(ValueError):
        model(**inp)
    with pytest.raises(ValueError):
        model(vision_feature_select_strategy="bogus", **_inputs(fx, cuda))


def test_sync_free_merge_with_collator_hint(cuda):
    """the Collator's merge_hint replaces the merge's host read-back; same outputs, and a wrong hint / image
    count still raises the reference's ValueError -- deferred to ops.check_deferred()."""
    from mantis_b200 import ops
    from mantis_b200.train import Collator
    fx = load_fixture("llava_batch_pad.pt")
    model = load_model(fx, torch.float32, cuda).eval()
    cfgk = fx["meta"]["cfg_kwargs"]
    col = Collator(pad_token_id=cfgk["pad_token_id"], image_token_index=cfgk["image_token_index"])
    hint = col.merge_hint(fx["input_ids"], cfgk["pad_token_id"])
    kw = _inputs(fx, cuda)
    with torch.no_grad():
        ref = model(**kw)
        got = model(**kw, merge_hint=hint)
    ops.check_deferred()
    assert torch.equal(ref.logits, got.logits) and torch.equal(ref.loss, got.loss)
    assert rel_err(got.logits, fx["logits"]) < 1e-3
    bad = dict(hint, max_image_tokens=hint["max_image_tokens"] + 1)
    with torch.no_grad():
        model(**kw, merge_hint=bad)
    with pytest.raises(ValueError):
        ops.check_deferred()
    ops.check_deferred()                                              # the queue is drained
