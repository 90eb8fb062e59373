#!/usr/bin/env python
"""Per-kernel SASS evidence for the Blackwell paths: counts of the tcgen05 / TMEM / TMA mnemonics in libmantis_b200.so
(`cuobjdump -sass`; runs without a GPU).  UTCHMMA = tcgen05.mma (bf16), LDTM/STTM = tcgen05.ld/st (TMEM), UTMALDG/UTMASTG =
TMA tensor load/store, UBLKCP = cp.async.bulk, UTCBAR = tcgen05.commit, HMMA = mma.sync (legacy tensor path), SYNCS =
mbarrier ops.     python tools/sass_summary.py > profiles/sass_summary_r02.txt"""
import collections
import hashlib
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "mantis_b200", "lib", "libmantis_b200.so")
MNEMONICS = ["UTCHMMA", "UTCBAR", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UBLKCP", "HMMA", "SYNCS", "MUFU.EX2", "FFMA"]


def demangle(names):
    try:
        out = subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True).stdout.splitlines()
        return dict(zip(names, out))
    except OSError:
        return {n: n for n in names}


def main():
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    counts, cur = collections.OrderedDict(), None
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1)
            counts[cur] = collections.Counter()
            continue
        if cur is None:
            continue
        m = re.search(r"/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if not m:
            continue
        op = m.group(1)
        counts[cur]["_total"] += 1
        for mn in MNEMONICS:
            if op == mn or op.startswith(mn + "."):
                counts[cur][mn] += 1
    dm = demangle(list(counts))
    sha = hashlib.sha256(open(LIB, "rb").read()).hexdigest()[:16]
    print(f"# cuobjdump -sass {os.path.relpath(LIB, ROOT)} (sha256 {sha}), sm_100a; instruction counts per kernel")
    print(f"{'kernel':70s} " + " ".join(f"{m:>8s}" for m in MNEMONICS) + f" {'total':>8s}")
    for fn, c in sorted(counts.items(), key=lambda kv: (-(kv[1]["UTCHMMA"] + kv[1]["UTMALDG"]), dm[kv[0]])):
        name = re.sub(r"\(anonymous namespace\)::", "", dm[fn])
        name = re.sub(r"\(.*", "", name).replace("void ", "")
        print(f"{name[:70]:70s} " + " ".join(f"{c[m]:8d}" for m in MNEMONICS) + f" {c['_total']:8d}")


if __name__ == "__main__":
    sys.exit(main())
