"""CPU tests: the C-ABI library builds, loads without a GPU and exports every symbol include/mantis_b200.h declares."""
import ctypes
import os
import subprocess

import pytest
import torch

from mantis_b200 import _lib


def test_library_loads_and_exports_every_declared_symbol():
    if not os.path.isfile(_lib.LIB_PATH):
        import __graft_entry__ as g
        g.build()
    lib = ctypes.CDLL(_lib.LIB_PATH)
    protos = _lib.parse_header()
    assert len(protos) >= 30
    for name in protos:
        assert hasattr(lib, name), f"{name} declared in include/mantis_b200.h but not exported"
    assert _lib.lib().mb200_version() >= 100


def test_no_undeclared_public_symbols():
    out = subprocess.run(["nm", "-D", "--defined-only", _lib.LIB_PATH], capture_output=True, text=True).stdout
    exported = {l.split()[-1] for l in out.splitlines() if " T mb200_" in l}
    declared = set(_lib.parse_header()) | {"mb200_set_last_error"}
    assert exported <= declared, f"exported but not declared: {sorted(exported - declared)}"


def test_ops_fail_loudly_without_cuda():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from mantis_b200 import ops
    with pytest.raises(_lib.MantisB200Error):
        ops.gemm(torch.zeros(8, 8), torch.zeros(8, 8))


def test_model_classes_importable_with_reference_names():
    from mantis_b200.models.mllava import LlavaConfig, LlavaForConditionalGeneration, MLlavaForConditionalGeneration  # noqa
    from helpers import load_fixture, build_from_meta
    fx = load_fixture("llava_siglip_full.pt")
    model = build_from_meta(fx["meta"])
    missing, unexpected = model.load_state_dict(fx["state_dict"], strict=False)
    assert not missing and not unexpected        # state-dict layout == reference's
    fx = load_fixture("mllava_clip.pt")
    model = build_from_meta(fx["meta"])
    missing, unexpected = model.load_state_dict(fx["state_dict"], strict=False)
    assert not missing and not unexpected


def test_paged_kv_allocator_host_logic():
    """page allocator of the paged KV cache (no kernels involved): unique page addresses inside the slabs, growth that
    only appends block-table entries, crop returning pages to the free list"""
    import torch
    from mantis_b200.models.kv_cache import B200KVCache
    c = B200KVCache(n_layers=2, slab_tokens=256)
    with pytest.raises(ValueError):
        B200KVCache().write(torch.zeros(1, 1, 2, 8), torch.zeros(1, 1, 2, 8), 0)       # n_layers unknown
    c._configure(3, 2, 8, torch.float32, torch.device("cpu"))
    assert len(c) == 2 and c.capacity() == 0
    c.ensure(130)                                      # 2 pages per sequence
    assert [len(b) for b in c.blocks] == [2, 2, 2] and c.capacity() == 256
    before = [list(b) for b in c.blocks]
    tab = c.device_table()
    assert tab.shape == (3, 8) and tab[:, :2].tolist() == before and int(tab[:, 2:].abs().sum()) == 0
    c.ensure(1500)                                     # needs more slabs
    assert len(c.slabs) >= 2 and all(b[:2] == f for b, f in zip(c.blocks, before))
    pages = [p for b in c.blocks for p in b]
    assert len(set(pages)) == len(pages) and not set(pages) & set(c.free)
    page_bytes = c.page_elems * 4
    for p in pages:
        assert any(s.data_ptr() <= p and p + page_bytes <= s.data_ptr() + s.numel() * 4 and (p - s.data_ptr()) % page_bytes == 0
                   for s in c.slabs)
    assert c.device_table().shape[1] == 16
    n_free = len(c.free)
    c.lengths = [300, 300]
    c.crop(129)
    assert [len(b) for b in c.blocks] == [2, 2, 2] and len(c.free) == n_free + 3 * 10 and c.get_seq_length() == 129
    assert c.device_table()[:, :2].tolist() == before


def test_idefics3_shell_has_the_reference_state_dict_layout():
    """same class names through the `mantis.models.idefics3` alias, same state-dict keys as the reference"""
    from helpers import load_fixture
    from mantis.models.idefics3 import Idefics3Config, Idefics3ForConditionalGeneration
    fx = load_fixture("idefics3_full.pt")
    model = Idefics3ForConditionalGeneration(Idefics3Config(**fx["cfg"]))
    missing, unexpected = model.load_state_dict(fx["state_dict"], strict=False)
    assert not missing and not unexpected
    assert model.model.image_seq_len == 16


def test_llava_next_shell_has_the_reference_state_dict_layout():
    from helpers import load_fixture
    from transformers import CLIPVisionConfig, LlamaConfig
    from mantis.models.mllava_next import LlavaNextConfig, LlavaNextForConditionalGeneration
    fx = load_fixture("llava_next_batch.pt")
    cfg = LlavaNextConfig(vision_config=CLIPVisionConfig(**fx["vision"]), text_config=LlamaConfig(**fx["text"]), **fx["cfg"])
    model = LlavaNextForConditionalGeneration(cfg)
    missing, unexpected = model.load_state_dict(fx["state_dict"], strict=False)
    assert not missing and not unexpected
    assert "image_newline" in dict(model.named_parameters())
    crops = [torch.zeros(3, 3, 8, 8) + 1, torch.zeros(1, 3, 8, 8) + 2]
    assert LlavaNextForConditionalGeneration._base_crops(crops)[:, 0, 0, 0].tolist() == [1.0, 2.0]
