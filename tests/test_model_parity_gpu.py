"""Model-level parity on the GPU: mantis_b200 (CUDA kernels through the C ABI) vs golden outputs produced by the
unmodified reference on CPU fp32 (oracle/make_golden.py).  Tolerances:
  fp32 run : logits max-relative-to-scale <= 1e-3 (north_star), in practice ~1e-5; loss |d| <= 1e-4; grads rel-L2 <= 1e-3
  bf16 run : rel-L2(logits) <= 3e-2 against the fp32 reference (bf16 has 2^-8 relative spacing; a 2+2 layer
             stack accumulates a few ulps), loss |d| <= 3e-2
"""
import pytest
import torch

from helpers import load_fixture, load_model, rel_err

pytestmark = pytest.mark.gpu

CASES = ["llava_siglip_full.pt", "llava_clip_default.pt", "llava_batch_pad.pt", "mllava_clip.pt"]


@pytest.mark.parametrize("name", CASES)
def test_forward_backward_fp32(cuda, name):
    fx = load_fixture(name)
    model = load_model(fx, torch.float32, cuda)
    model.train()
    model.materialize_logits_in_training = True
    dev = lambda t: t.to(cuda) if t is not None else None
    out = model(input_ids=dev(fx["input_ids"]), pixel_values=dev(fx["pixel_values"]),
                attention_mask=dev(fx["attention_mask"]), labels=dev(fx["labels"]))
    ref_logits = fx["logits"]
    assert out.logits.shape == ref_logits.shape
    # padded positions are don't-care (the reference fills them with an unspecified average); compare where mask==1
    got = out.logits.detach().float().cpu()
    if name == "llava_batch_pad.pt":
        from oracle.merge_oracle import merge_oracle  # noqa: F401  (mask recomputed below from the fixture)
    scale = ref_logits.abs().max().item()
    valid = torch.ones(ref_logits.shape[:2], dtype=torch.bool)
    if name == "llava_batch_pad.pt":
        # final attention mask from our own merge (bit-exactness of that is tested separately)
        valid = _final_mask(model, fx, cuda).bool().cpu()
    err = ((got - ref_logits).abs().amax(-1))[valid].max().item()
    assert err <= 1e-3 * scale, f"logits max err {err} (scale {scale})"
    assert abs(out.loss.item() - fx["loss"].item()) <= 1e-4
    out.loss.backward()
    params = dict(model.named_parameters())
    for k, g in fx["grads"].items():
        assert params[k].grad is not None, k
        e = rel_err(params[k].grad, g)
        assert e <= 2e-3, f"grad {k}: rel err {e}"


def _final_mask(model, fx, cuda):
    with torch.no_grad():
        emb = model.get_input_embeddings()(fx["input_ids"].to(cuda))
        P = {"full": 64, "default": 63}[model.config.vision_feature_select_strategy] if False else None
        feats = model._image_features(fx["pixel_values"].to(cuda), model.config.vision_feature_layer,
                                      model.config.vision_feature_select_strategy)
        _, mask, _, _ = model._merge_input_ids_with_image_features(feats, emb, fx["input_ids"].to(cuda),
                                                                   fx["attention_mask"].to(cuda), None)
    return mask


@pytest.mark.parametrize("name", CASES)
def test_fused_loss_path_fp32(cuda, name):
    """training default: fused LM-head + CE (no logits) must give the same loss and gradients"""
    fx = load_fixture(name)
    model = load_model(fx, torch.float32, cuda)
    model.train()
    dev = lambda t: t.to(cuda) if t is not None else None
    out = model(input_ids=dev(fx["input_ids"]), pixel_values=dev(fx["pixel_values"]),
                attention_mask=dev(fx["attention_mask"]), labels=dev(fx["labels"]))
    assert out.logits is None
    assert abs(out.loss.item() - fx["loss"].item()) <= 1e-4
    out.loss.backward()
    params = dict(model.named_parameters())
    for k, g in fx["grads"].items():
        assert rel_err(params[k].grad, g) <= 2e-3, k


@pytest.mark.parametrize("name", CASES)
def test_forward_bf16(cuda, name):
    fx = load_fixture(name)
    model = load_model(fx, torch.bfloat16, cuda)
    model.eval()
    dev = lambda t: t.to(cuda) if t is not None else None
    with torch.no_grad():
        out = model(input_ids=dev(fx["input_ids"]), pixel_values=dev(fx["pixel_values"]).bfloat16(),
                    attention_mask=dev(fx["attention_mask"]), labels=dev(fx["labels"]))
    valid = torch.ones(fx["logits"].shape[:2], dtype=torch.bool)
    if name == "llava_batch_pad.pt":
        valid = _final_mask(model, fx, cuda).bool().cpu()
    got = out.logits.float().cpu()[valid]; ref = fx["logits"][valid]
    assert rel_err(got, ref) <= 3e-2, rel_err(got, ref)
    assert abs(out.loss.item() - fx["loss"].item()) <= 3e-2


def test_greedy_generate_matches_reference(cuda):
    fx = load_fixture("greedy_llava.pt")
    model = load_model(fx, torch.float32, cuda).eval()
    ids = fx["input_ids"].to(cuda)
    n_new = fx["generated"].shape[1] - ids.shape[1]
    seq = model.greedy_generate(ids, pixel_values=fx["pixel_values"].to(cuda), max_new_tokens=n_new)
    assert seq.cpu().tolist() == fx["generated"].tolist()


def test_hf_generate_api(cuda):
    """model.generate(...) through transformers' GenerationMixin (what chat_mllava calls) == our greedy loop"""
    fx = load_fixture("greedy_llava.pt")
    model = load_model(fx, torch.float32, cuda).eval()
    ids = fx["input_ids"].to(cuda)
    n_new = fx["generated"].shape[1] - ids.shape[1]
    out = model.generate(input_ids=ids, pixel_values=fx["pixel_values"].to(cuda), attention_mask=torch.ones_like(ids),
                         max_new_tokens=n_new, do_sample=False, num_beams=1)
    assert out[0, ids.shape[1]:].cpu().tolist() == fx["generated"][0, ids.shape[1]:].tolist()


def test_native_decode_engine_matches_python_path(cuda):
    """bf16, head_dim 128: the C++ decode step (skinny GEMMs + split-KV attention, one call per token) must reproduce
    the Python-path decode (tensor-core / SIMT kernels) logits, step after step."""
    from transformers import LlamaConfig, SiglipVisionConfig
    import mantis_b200.ops as om
    from mantis_b200.models.kv_cache import B200KVCache
    from mantis_b200.models.mllava import LlavaConfig, LlavaForConditionalGeneration
    vc = SiglipVisionConfig(hidden_size=64, intermediate_size=128, num_hidden_layers=2, num_attention_heads=4,
                            image_size=112, patch_size=14)
    tc = LlamaConfig(hidden_size=256, intermediate_size=512, num_hidden_layers=2, num_attention_heads=2,
                     num_key_value_heads=1, vocab_size=1000, rms_norm_eps=1e-5, rope_theta=500000.0)
    cfg = LlavaConfig(vision_config=vc, text_config=tc, image_token_index=990, pad_token_id=991, vocab_size=1000,
                      vision_feature_select_strategy="full")
    torch.manual_seed(0)
    model = LlavaForConditionalGeneration(cfg).to(cuda).to(torch.bfloat16).eval()
    with torch.no_grad():
        for p in model.parameters():
            if p.dim() >= 2:
                p.mul_(3.0)
    ids = torch.randint(0, 980, (2, 20), device=cuda); ids[:, 3] = 990
    pv = torch.randn(2, 3, 112, 112, device=cuda).bfloat16()

    def run(native):
        old = om.FORCE_GENERIC
        logits = []
        with torch.no_grad():
            cache = B200KVCache()
            am = torch.ones_like(ids)
            out = model(input_ids=ids, pixel_values=pv, attention_mask=am, past_key_values=cache, use_cache=True, logits_to_keep=1)
            nxt = out.logits[:, -1].argmax(-1)
            for _ in range(5):
                am = torch.cat([am, torch.ones_like(nxt[:, None])], 1)
                om.FORCE_GENERIC = not native
                try:
                    out = model(input_ids=nxt[:, None], pixel_values=pv, attention_mask=am, past_key_values=cache,
                                use_cache=True, logits_to_keep=1)
                finally:
                    om.FORCE_GENERIC = old
                logits.append(out.logits[:, -1].float().clone())
                nxt = torch.tensor([7, 11], device=cuda) + len(logits)          # fixed tokens: identical inputs both ways
        return logits, hasattr(cache, "_engine")

    a, used_a = run(True)
    b, used_b = run(False)
    assert used_a and not used_b
    for x, y in zip(a, b):
        assert rel_err(x, y) < 3e-2, rel_err(x, y)


def test_fullwidth_bf16_tensor_core_path_vs_fp32_simt_path(cuda):
    """Full-width (d=4096, 32q/8kv x 128, ff 14336, SigLIP d=1152 / head_dim 72) one-layer model: the bf16 run goes through
    the tcgen05 GEMM (1- and 2-CTA), the tcgen05 attention fwd+bwd, the padded-head ViT path and the fused LM-head/CE;
    the fp32 run of the SAME weights goes through the SIMT kernels.  They must agree to bf16 rounding."""
    from transformers import LlamaConfig, SiglipVisionConfig
    from mantis_b200.models.mllava import LlavaConfig, LlavaForConditionalGeneration
    vc = SiglipVisionConfig(hidden_size=1152, intermediate_size=4304, num_hidden_layers=2, num_attention_heads=16,
                            image_size=112, patch_size=14, hidden_act="gelu_pytorch_tanh", layer_norm_eps=1e-6)
    tc = LlamaConfig(hidden_size=4096, intermediate_size=14336, num_hidden_layers=1, num_attention_heads=32,
                     num_key_value_heads=8, vocab_size=2048, rms_norm_eps=1e-5, rope_theta=500000.0)
    cfg = LlavaConfig(vision_config=vc, text_config=tc, image_token_index=2040, pad_token_id=2041, vocab_size=2048,
                      vision_feature_select_strategy="full")
    torch.manual_seed(0)
    m16 = LlavaForConditionalGeneration(cfg).to(cuda).to(torch.bfloat16)
    m32 = LlavaForConditionalGeneration(cfg).to(cuda)
    m32.load_state_dict({k: v.float() for k, v in m16.state_dict().items()})
    for m in (m16, m32):
        for n, p in m.named_parameters():
            if "vision_tower" in n:
                p.requires_grad_(False)
        m.train()
    ids = torch.randint(0, 2000, (1, 700), device=cuda)
    ids[0, 5] = 2040; ids[0, 300] = 2040; ids[0, 301] = 2040
    labels = ids.clone(); labels[ids == 2040] = -100
    pv = torch.randn(3, 3, 112, 112, device=cuda)
    outs = []
    for m, dt in ((m16, torch.bfloat16), (m32, torch.float32)):
        out = m(input_ids=ids, pixel_values=pv.to(dt), attention_mask=torch.ones_like(ids), labels=labels)
        out.loss.backward()
        outs.append(out)
    assert abs(outs[0].loss.item() - outs[1].loss.item()) < 3e-2, (outs[0].loss.item(), outs[1].loss.item())
    p16, p32 = dict(m16.named_parameters()), dict(m32.named_parameters())
    for k in ["language_model.model.layers.0.self_attn.q_proj.weight", "language_model.model.layers.0.self_attn.v_proj.weight",
              "language_model.model.layers.0.mlp.down_proj.weight", "language_model.lm_head.weight",
              "multi_modal_projector.linear_1.weight", "language_model.model.layers.0.input_layernorm.weight"]:
        e = rel_err(p16[k].grad, p32[k].grad)
        assert e < 5e-2, (k, e)
    m16.eval(); m32.eval()
    with torch.no_grad():
        l16 = m16(input_ids=ids, pixel_values=pv.bfloat16(), attention_mask=torch.ones_like(ids)).logits
        l32 = m32(input_ids=ids, pixel_values=pv, attention_mask=torch.ones_like(ids)).logits
    assert rel_err(l16, l32) < 3e-2, rel_err(l16, l32)


def test_batched_greedy_with_left_padding_and_uneven_images(cuda):
    """decode-time branch (ref :477-508): a left-padded batch whose samples carry different numbers of images keeps masked
    slots inside the merged prompt; the cached decode must reproduce the reference's cache-free loop token for token
    (fixture: oracle/make_golden_batch_greedy.py)"""
    fx = load_fixture("greedy_llava_batch.pt")
    model = load_model(fx, torch.float32, cuda).eval()
    ids = fx["input_ids"].to(cuda); att = fx["attention_mask"].to(cuda); pv = fx["pixel_values"].to(cuda)
    n_new = fx["generated"].shape[1] - ids.shape[1]
    seq = model.greedy_generate(ids, pixel_values=pv, attention_mask=att, max_new_tokens=n_new)
    assert seq.cpu().tolist() == fx["generated"].tolist()
    out = model.generate(input_ids=ids, pixel_values=pv, attention_mask=att, max_new_tokens=n_new, do_sample=False,
                         num_beams=1, pad_token_id=301)
    assert out.cpu().tolist() == fx["generated"].tolist()


def test_decode_engine_bitexact_with_and_without_pdl(cuda, tmp_path):
    """programmatic dependent launch only changes WHEN the kernels of a decode step become resident, never what they read:
    12 native decode steps give bit-identical logits with MB200_PDL=1 (default) and MB200_PDL=0."""
    import os
    import subprocess
    import sys
    probe = os.path.join(os.path.dirname(os.path.abspath(__file__)), "pdl_probe.py")
    outs = []
    for flag in ("1", "0"):
        path = str(tmp_path / f"logits_pdl{flag}.pt")
        env = dict(os.environ, MB200_PDL=flag)
        r = subprocess.run([sys.executable, probe, path], env=env, capture_output=True, text=True, timeout=600)
        assert r.returncode == 0, r.stderr[-2000:]
        outs.append(torch.load(path))
    assert outs[0].shape == (12, 2, 2000) and torch.isfinite(outs[0]).all()
    assert torch.equal(outs[0], outs[1])


def test_labels_none_gives_nan_loss_like_the_reference(cuda):
    """a7 (ref modeling_llava.py:475-476, 523-537): with pixel_values and NO labels the reference substitutes all-ignore labels,
    so its `loss` is CrossEntropyLoss over zero rows == NaN (not None); without images and without labels the loss is None"""
    fx = load_fixture("llava_siglip_full.pt")
    model = load_model(fx, torch.float32, cuda).eval()
    ids = fx["input_ids"].to(cuda)
    with torch.no_grad():
        out = model(input_ids=ids, pixel_values=fx["pixel_values"].to(cuda), attention_mask=fx["attention_mask"].to(cuda))
    assert out.loss is not None and torch.isnan(out.loss) and out.loss.dtype == torch.float32
    assert (out.logits.cpu() - fx["logits"]).abs().max().item() <= 1e-3 * fx["logits"].abs().max().item()
    text_only = ids[:, 45:]                                   # no <image> placeholder in this span
    assert not bool((text_only == model.config.image_token_index).any())
    with torch.no_grad():
        out = model(input_ids=text_only, attention_mask=torch.ones_like(text_only))
    assert out.loss is None


def test_generate_on_a_worker_thread_matches_the_main_thread(cuda):
    """chat_mllava_stream (ref utils.py:100-186) runs model.generate() on a second Python thread: the bf16 path (tcgen05 prefill,
    native decode engine with its call-scoped launch mode) must give the same tokens from a worker thread as from the main one"""
    import threading
    from transformers import LlamaConfig, SiglipVisionConfig
    from mantis_b200.models import decode_engine
    from mantis_b200.models.mllava import LlavaConfig, LlavaForConditionalGeneration
    vc = SiglipVisionConfig(hidden_size=64, intermediate_size=128, num_hidden_layers=2, num_attention_heads=4,
                            image_size=112, patch_size=14)
    tc = LlamaConfig(hidden_size=256, intermediate_size=512, num_hidden_layers=2, num_attention_heads=2,
                     num_key_value_heads=1, vocab_size=1000, rms_norm_eps=1e-5, rope_theta=500000.0)
    cfg = LlavaConfig(vision_config=vc, text_config=tc, image_token_index=990, pad_token_id=991, vocab_size=1000,
                      vision_feature_select_strategy="full")
    torch.manual_seed(3)
    model = LlavaForConditionalGeneration(cfg).to(cuda).to(torch.bfloat16).eval()
    with torch.no_grad():
        for p in model.parameters():
            if p.dim() >= 2:
                p.mul_(4.0)
    ids = torch.randint(0, 980, (2, 200), device=cuda); ids[:, 3] = 990
    pv = torch.randn(2, 3, 112, 112, device=cuda).bfloat16()
    kw = dict(input_ids=ids, pixel_values=pv, attention_mask=torch.ones_like(ids), max_new_tokens=12, do_sample=False,
              num_beams=1, pad_token_id=991)
    main = model.generate(**kw)
    got, errs = [], []

    def work():
        try:
            torch.cuda.set_device(cuda)
            got.append(model.generate(**kw))
            torch.cuda.synchronize()
        except BaseException as e:  # noqa
            errs.append(e)

    before = decode_engine.native_steps
    ts = [threading.Thread(target=work) for _ in range(2)]
    for t in ts:
        t.start(); t.join()
    assert not errs, errs
    assert decode_engine.native_steps - before >= 2 * 11
    for g in got:
        assert torch.equal(g, main)


def _tiny_bf16_llava(cuda, seed=3, scale=4.0):
    from transformers import LlamaConfig, SiglipVisionConfig
    from mantis_b200.models.mllava import LlavaConfig, LlavaForConditionalGeneration
    vc = SiglipVisionConfig(hidden_size=64, intermediate_size=128, num_hidden_layers=2, num_attention_heads=4,
                            image_size=112, patch_size=14)
    tc = LlamaConfig(hidden_size=256, intermediate_size=512, num_hidden_layers=2, num_attention_heads=2,
                     num_key_value_heads=1, vocab_size=1000, rms_norm_eps=1e-5, rope_theta=500000.0)
    cfg = LlavaConfig(vision_config=vc, text_config=tc, image_token_index=990, pad_token_id=991, vocab_size=1000,
                      vision_feature_select_strategy="full")
    torch.manual_seed(seed)
    model = LlavaForConditionalGeneration(cfg).to(cuda).to(torch.bfloat16).eval()
    with torch.no_grad():
        for p in model.parameters():
            if p.dim() >= 2:
                p.mul_(scale)
    return model


@pytest.mark.parametrize("case", ["plain", "left_padded_uneven_images", "eos"])
def test_fast_greedy_generate_equals_generation_mixin(cuda, case, monkeypatch):
    """model.generate() takes the sync-free greedy loop for plain greedy decoding; its output (tokens, padding after EOS, length)
    must equal what transformers' GenerationMixin.generate returns for the same call (MB200_FAST_GENERATE=0)."""
    from mantis_b200.models import decode_engine
    model = _tiny_bf16_llava(cuda)
    g = torch.Generator().manual_seed(7)
    ids = torch.randint(0, 980, (3, 60), generator=g)
    ids[:, 5] = 990
    am = torch.ones_like(ids)
    n_img = 3
    if case == "left_padded_uneven_images":
        ids[0, :9] = 991; am[0, :9] = 0; ids[0, 5] = 991          # row 0: left padding, image moved inside the body
        ids[0, 20] = 990
        ids[1, 30] = 990; n_img = 4                               # row 1 carries two images
    pv = torch.randn(n_img, 3, 112, 112, generator=g).bfloat16()
    kw = dict(input_ids=ids.to(cuda), pixel_values=pv.to(cuda), attention_mask=am.to(cuda), max_new_tokens=40, do_sample=False,
              num_beams=1, pad_token_id=991)
    monkeypatch.setenv("MB200_FAST_GENERATE", "0")
    ref = model.generate(**kw)
    if case == "eos":                                             # an EOS id that the greedy continuation actually emits
        new = ref[:, ids.shape[1]:]
        kw["eos_token_id"] = [int(new[0, 7]), int(new[1, 19])]
        ref = model.generate(**kw)
        assert ref.shape[1] < ids.shape[1] + 40 or bool((ref[:, ids.shape[1]:] == 991).any())
    monkeypatch.setenv("MB200_FAST_GENERATE", "1")
    before = decode_engine.native_steps
    out = model.generate(**kw)
    assert decode_engine.native_steps > before
    assert out.shape == ref.shape and torch.equal(out, ref), (out[:, 60:].tolist(), ref[:, 60:].tolist())
    # anything beyond plain greedy decoding keeps going through GenerationMixin (here: sampling)
    if case == "eos":                                             # min_new_tokens: fast when no EOS shows up early, else GenerationMixin
        kw2 = dict(kw, min_new_tokens=12)
        monkeypatch.setenv("MB200_FAST_GENERATE", "0"); ref2 = model.generate(**kw2)
        monkeypatch.setenv("MB200_FAST_GENERATE", "1"); out2 = model.generate(**kw2)
        assert torch.equal(out2, ref2)
    assert model._fast_greedy_plan(None, None, dict(kw, do_sample=True)) is None
    assert model._fast_greedy_plan(None, None, dict(kw, streamer=object())) is None
