"""CPU tests: the C-ABI library builds, loads without a GPU and exports every symbol include/mantis_b200.h declares."""
import ctypes
import os
import subprocess

import pytest

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
