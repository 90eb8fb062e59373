#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel-family totals and shares for the LAST
optimizer step in the file (a step ends with its adamw launch; with fewer than two of them in the capture the whole file --
model construction included -- is summarised, so capture at least `--warmup 1 --steps 1` completely)."""
import csv
import re
import sys
from collections import defaultdict


def short(name):
    name = re.sub(r"\(anonymous namespace\)::", "", name)
    m = re.search(r"([A-Za-z_0-9:]+)\s*(<|\()", name.replace("void ", ""))
    base = m.group(1) if m else name[:60]
    base = base.split("::")[-1]
    if "gemm_sm100" in name:
        t = re.search(r"gemm_sm100\w*<([^>]*)>", name)
        base += "<" + (t.group(1).replace(" ", "") if t else "") + ">"
    return base


def main(path):
    rows = []
    with open(path, newline="") as f:
        lines = [l for l in f if not l.startswith("==")]
    for r in csv.DictReader(lines):
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(r["Metric Value"].replace(",", ""))
        unit = r["Metric Unit"]
        ns = v * {"ns": 1, "us": 1e3, "ms": 1e6, "s": 1e9}.get(unit, 1)
        rows.append((short(r["Kernel Name"]), ns, r["Kernel Name"]))
    # step boundaries: indices where an adamw run ends
    ends = [i for i in range(len(rows)) if rows[i][0].startswith("adamw") and (i + 1 == len(rows) or not rows[i + 1][0].startswith("adamw"))]
    if len(ends) >= 2:
        lo, hi = ends[-2] + 1, ends[-1] + 1
    elif ends:
        lo, hi = 0, ends[-1] + 1
    else:
        lo, hi = 0, len(rows)
    agg = defaultdict(lambda: [0.0, 0])
    for k, ns, _ in rows[lo:hi]:
        agg[k][0] += ns; agg[k][1] += 1
    total = sum(v[0] for v in agg.values())
    print(f"# {path}: launches {hi - lo} in the last step, serialized kernel time {total / 1e6:.1f} ms")
    print(f"{'kernel':58s} {'launches':>8s} {'ms':>10s} {'share':>7s} {'avg_us':>9s}")
    for k, (ns, n) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:40]:
        print(f"{k[:58]:58s} {n:8d} {ns / 1e6:10.2f} {100 * ns / total:6.1f}% {ns / n / 1e3:9.1f}")


if __name__ == "__main__":
    main(sys.argv[1])
