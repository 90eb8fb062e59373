"""sha256 over the source files a kernel is built from: ties an ncu capture under profiles/ to the kernel that is actually in
the tree (bench.py refuses numbers whose hash no longer matches)."""
import hashlib
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "mantis_b200", "csrc")
KERNEL_SOURCES = {
    "gemm_sm100_2cta_kernel": ["gemm_sm100_2cta.cu", "gemm_epi.cuh", "sm100_ptx.cuh", "tmap.cuh", "common.cuh"],
    "merge_rows_kernel": ["merge.cu", "common.cuh"],
    "attn_fwd2_sm100_kernel": ["attn_fwd2_sm100.cu", "sm100_ptx.cuh", "tmap.cuh", "common.cuh"],
    "attn_bwd_sm100": ["attn_bwd_sm100.cu", "sm100_ptx.cuh", "tmap.cuh", "common.cuh"],
    "decode": ["decode.cu", "decode_engine.cu", "common.cuh"],
}


def kernel_hash(kernel):
    h = hashlib.sha256()
    for name in KERNEL_SOURCES[kernel]:
        with open(os.path.join(CSRC, name), "rb") as f:
            h.update(name.encode()); h.update(f.read())
    return h.hexdigest()[:16]
