#!/usr/bin/env python
"""`ncu --set full` capture (.ncu-rep) -> the per-launch numbers bench.py quotes (roofline.traffic, tensor-pipe %), stamped
with the hash of the kernel's sources so a stale capture cannot be quoted for a changed kernel.

    python tools/ncu_to_json.py gpurun_out/ncu_gemm2cta.ncu-rep gemm_sm100_2cta_kernel profiles/ncu_gemm2cta_r02.json
"""
import csv
import io
import json
import os
import subprocess
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from kernel_hash import kernel_hash  # noqa: E402

UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12, "ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}
WANT = {"dram__bytes_read.sum": "dram_read_bytes", "dram__bytes_write.sum": "dram_write_bytes",
        "gpu__time_duration.sum": "duration_us",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed": "tensor_pipe_active_pct",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active": "xu_pipe_pct",
        "smsp__issue_active.avg.pct_of_peak_sustained_active": "issue_active_pct",
        "sm__cycles_elapsed.max.per_second": "sm_ghz",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed": "dram_throughput_pct"}


def main(rep, kernel, out):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    col = {h: i for i, h in enumerate(hdr)}
    launches = []
    for r in data:
        name = r[col["Kernel Name"]]
        if kernel.split("_kernel")[0] not in name:
            continue
        rec = {"kernel": name.split("(")[0].replace("void ", "").replace("<unnamed>::", "")}
        for metric, key in WANT.items():
            if metric in col:
                v = float(r[col[metric]].replace(",", ""))
                rec[key] = v * UNIT.get(units[col[metric]], 1.0)
        rec["traffic_bytes"] = rec.get("dram_read_bytes", 0.0) + rec.get("dram_write_bytes", 0.0)
        launches.append(rec)
    if not launches:
        raise SystemExit(f"no launch of {kernel} in {rep}")
    doc = {"kernel": kernel, "kernel_hash": kernel_hash(kernel), "source": os.path.basename(rep),
           "command": "ncu --set full --clock-control none --import-source on (see profiles/README_r02.md)", "launches": launches}
    with open(out, "w") as f:
        json.dump(doc, f, indent=1)
    print(json.dumps(doc, indent=1))


if __name__ == "__main__":
    main(*sys.argv[1:4])
