for g in 8 11 16 21 31; do
  MB200_GEMM_GROUP_M=$g ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:gemm_sm100_2cta -c 6 --csv --log-file gpurun_out/gm_$g.csv python tools/ncu_targets.py gemm2cta > /dev/null 2>&1
done
echo swept
