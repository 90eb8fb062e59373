"""CPU tests over the committed evidence: the ncu numbers bench.py quotes belong to the kernels that are in the library NOW,
and the hot kernels really are tcgen05 / TMA code (static SASS check, no GPU needed)."""
import glob
import json
import os
import shutil
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))

pytestmark = pytest.mark.skipif(shutil.which("cuobjdump") is None and not os.path.isfile("/usr/local/cuda/bin/cuobjdump"),
                                reason="cuobjdump not available")


def _ensure_lib():
    from mantis_b200 import _lib
    if not os.path.isfile(_lib.LIB_PATH):
        import __graft_entry__ as g
        g.build()
    return _lib.LIB_PATH


def test_committed_ncu_numbers_match_the_built_kernels():
    """bench.py refuses `roofline.traffic` from a capture whose kernel hash differs from the built kernel's SASS hash
    (bench.ncu_evidence); a kernel change without a re-capture must fail HERE, not as a silently missing number."""
    _ensure_lib()
    from kernel_hash import kernel_hash
    for pattern, kernel in (("ncu_gemm2cta_r*.json", "gemm_sm100_2cta_kernel"), ("ncu_merge_rows_r*.json", "merge_rows_kernel")):
        files = sorted(glob.glob(os.path.join(ROOT, "profiles", pattern)), key=os.path.getmtime)
        assert files, pattern
        rec = json.load(open(files[-1]))
        assert rec["kernel_hash"] == kernel_hash(kernel), (files[-1], rec["kernel_hash"], kernel_hash(kernel))
        assert rec["launches"] and all(l["traffic_bytes"] > 0 for l in rec["launches"])


def test_hot_kernels_are_tcgen05_and_tma_code():
    lib = _ensure_lib()
    exe = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    sass = subprocess.run([exe, "-sass", lib], capture_output=True, text=True, timeout=600, check=True).stdout
    per_fn, cur = {}, None
    for line in sass.splitlines():
        if "Function :" in line:
            cur = line.split("Function :")[1].strip()
            per_fn[cur] = {"UTCHMMA": 0, "UTMALDG": 0, "UTMASTG": 0, "UTMAREDG": 0, "LDTM": 0, "UBLKCP": 0}
        elif cur:
            for k in per_fn[cur]:
                if k in line:
                    per_fn[cur][k] += 1

    def total(name_part, op):
        return sum(v[op] for k, v in per_fn.items() if name_part in k)
    # GEMM (1- and 2-CTA), attention forward and both backward kernels: tensor-core MMA from TMA-fed shared memory into TMEM
    for fn in ("gemm_sm100_2cta_kernel", "gemm_sm100_kernel", "attn_fwd2_sm100_kernel", "attn_bwd_dkv_sm100_kernel",
               "attn_bwd_dq2_sm100_kernel"):
        assert total(fn, "UTCHMMA") > 0 and total(fn, "UTMALDG") > 0 and total(fn, "LDTM") > 0, fn
    # TMA-store epilogue of the 2-CTA GEMM (bf16 store and fp32 reduce-add) and the dS^T store of the single-pass backward
    assert total("gemm_sm100_2cta_kernel", "UTMASTG") > 0 and total("gemm_sm100_2cta_kernel", "UTMAREDG") > 0
    assert total("attn_bwd_dkv_sm100_kernel", "UTMASTG") > 0
    # decode weight stream: bulk copies into the shared-memory ring
    assert total("skinny_rows_kernel", "UBLKCP") > 0 and total("skinny_ring_kernel", "UBLKCP") > 0
