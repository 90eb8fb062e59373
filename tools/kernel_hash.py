"""Identity of a kernel AS BUILT: sha256 over the SASS of every function in libmantis_b200.so whose name contains the kernel's
name (`cuobjdump -sass`, instruction text only).  It ties an ncu capture under profiles/ to the kernel that is actually in the
library -- bench.py refuses numbers whose hash no longer matches -- and does not move when an unrelated kernel or a shared
header changes.  Falls back to a hash of the kernel's source files when cuobjdump is not available."""
import functools
import hashlib
import os
import re
import shutil
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "mantis_b200", "csrc")
LIB = os.path.join(ROOT, "mantis_b200", "lib", "libmantis_b200.so")
KERNEL_SOURCES = {
    "gemm_sm100_2cta_kernel": ["gemm_sm100_2cta.cu", "gemm_epi.cuh"],
    "merge_rows_kernel": ["merge.cu"],
    "attn_fwd2_sm100_kernel": ["attn_fwd2_sm100.cu"],
    "attn_bwd": ["attn_bwd_sm100.cu"],
    "decode": ["decode.cu", "decode_engine.cu"],
}


@functools.lru_cache(maxsize=1)
def _sass_by_function():
    exe = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.isfile(exe) or not os.path.isfile(LIB):
        return None
    try:
        out = subprocess.run([exe, "-sass", LIB], capture_output=True, text=True, timeout=300, check=True).stdout
    except Exception:  # noqa
        return None
    funcs, cur = {}, None
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1)
            funcs[cur] = []
            continue
        if cur is None:
            continue
        m = re.search(r"/\*[0-9a-f]{4,}\*/\s+(.*?);", line)
        if m:
            funcs[cur].append(m.group(1).strip())
    return funcs


def kernel_hash(kernel):
    funcs = _sass_by_function()
    h = hashlib.sha256()
    if funcs:
        # instruction text only, functions ordered by their own digest: anonymous-namespace symbol names carry a hash of the
        # source PATH and must not leak into the identity
        digests = sorted(hashlib.sha256("\n".join(funcs[n]).encode()).hexdigest() for n in funcs if kernel in n)
        if digests:
            h.update("".join(digests).encode())
            return "sass:" + h.hexdigest()[:16]
    for name in KERNEL_SOURCES.get(kernel, []):
        with open(os.path.join(CSRC, name), "rb") as f:
            h.update(name.encode()); h.update(f.read())
    return "src:" + h.hexdigest()[:16]


if __name__ == "__main__":
    import sys
    for k in sys.argv[1:] or list(KERNEL_SOURCES):
        print(k, kernel_hash(k))
