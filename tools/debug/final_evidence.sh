#!/bin/bash
# Round-end evidence run (one gpurun call): full GPU suite, smoke, ncu captures the bench JSON quotes, the bench itself,
# the kernel micro-benchmarks.  Everything lands in gpurun_out/; copy what should be judged into profiles/.
cd "$(dirname "$0")/../.."
timeout 400 python -m pytest tests -m gpu -q 2>&1 | tail -15 > gpurun_out/pytest_gpu_final.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/smoke_final.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:gemm_sm100_2cta -c 3 -f -o gpurun_out/ncu_gemm2cta_final python tools/ncu_targets.py gemm2cta > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:merge_rows_kernel -c 2 -f -o gpurun_out/ncu_merge_rows_final python tools/ncu_targets.py merge > /dev/null 2>&1
python tools/ncu_to_json.py gpurun_out/ncu_gemm2cta_final.ncu-rep gemm_sm100_2cta_kernel profiles/ncu_gemm2cta_r02_final.json > gpurun_out/ncu_to_json.log 2>&1
python tools/ncu_to_json.py gpurun_out/ncu_merge_rows_final.ncu-rep merge_rows_kernel profiles/ncu_merge_rows_r02_final.json >> gpurun_out/ncu_to_json.log 2>&1
cp profiles/ncu_gemm2cta_r02_final.json profiles/ncu_merge_rows_r02_final.json gpurun_out/
timeout 600 python bench.py > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err
timeout 250 python tools/bench_kernels.py gemm attn row decode > gpurun_out/kernel_bench_final.jsonl 2>&1
MB200_RMSNORM_BWD_PREFETCH=0 timeout 100 python tools/bench_kernels.py row > gpurun_out/kernel_bench_row_noprefetch.jsonl 2>&1
echo evidence done
