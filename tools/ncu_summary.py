#!/usr/bin/env python
"""Text summary of an `ncu --set full` capture (.ncu-rep): the handful of metrics the roofline discussion in DESIGN.md uses,
for the first launch of every distinct kernel.      python tools/ncu_summary.py profiles/ncu_attn_r02_final.ncu-rep [...]"""
import csv
import io
import subprocess
import sys

METRICS = ["gpu__time_duration.sum", "sm__cycles_elapsed.max.per_second",
           "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
           "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
           "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
           "sm__warps_active.avg.pct_of_peak_sustained_active", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
           "dram__bytes_read.sum", "dram__bytes_write.sum", "launch__registers_per_thread", "launch__block_size",
           "launch__grid_size", "launch__shared_mem_per_block_dynamic"]


def main(paths):
    for rep in paths:
        raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
        rows = list(csv.reader(io.StringIO(raw)))
        hdr, units, data = rows[0], rows[1], rows[2:]
        col = {h: i for i, h in enumerate(hdr)}
        print(f"source: {rep}")
        seen = set()
        for r in data:
            name = r[col["Kernel Name"]]
            key = name.split("(")[0]
            if key in seen:
                continue
            seen.add(key)
            print(f"\nkernel: {name[:120]}")
            for m in METRICS:
                if m in col:
                    print(f"    {m:75s} {r[col[m]]} {units[col[m]]}")
        print()


if __name__ == "__main__":
    main(sys.argv[1:])
