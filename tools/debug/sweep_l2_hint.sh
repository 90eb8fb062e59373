#!/bin/bash
# DRAM bytes and duration of the 2-CTA GEMM (gate/up fwd, dgrad, wgrad at M = 7864) for every L2 eviction-hint setting
for h in 0 1 2 3; do
  MB200_GEMM_L2_HINT=$h ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none \
    -k regex:gemm_sm100_2cta -c 6 --csv --log-file gpurun_out/l2hint_$h.csv python tools/ncu_targets.py gemm2cta > /dev/null 2>&1
done
for h in 0 1 2 3; do
  MB200_GEMM_L2_HINT=$h python tools/bench_kernels.py gemm > gpurun_out/kb_gemm_l2hint_$h.jsonl 2>&1
done
echo swept
