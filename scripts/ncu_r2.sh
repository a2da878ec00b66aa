#!/bin/bash
# ncu evidence for round 2 (1 GPU).  Each command runs plain first (exit 0), then under ncu, as
# B200_PROFILING.md requires.  usage: ncu_r2.sh posthoc | lasso   (two calls: gpurun_out <= 64 MiB)
mkdir -p gpurun_out
NULLCMD="python bench.py --workload null --block 2048 --steps 1 --warmup 1"
C5CMD="python bench.py --workload cfg5 --block 2048 --steps 1 --warmup 1"
LASSO="python bench.py --workload cfg3 --block 96 --steps 1 --warmup 0 --no-e2e --no-cpu"
if [ "$1" = "posthoc" ]; then
  $NULLCMD > gpurun_out/ncu_null_plain.log 2>&1 &&
  ncu --set full --clock-control none --import-source on -k regex:calc_post -c 1 -f -o gpurun_out/r2_calc_post $NULLCMD > gpurun_out/ncu_null.log 2>&1
  echo "null ncu rc=$?"
  $C5CMD > gpurun_out/ncu_c5_plain.log 2>&1 &&
  ncu --set full --clock-control none --import-source on -k regex:"sample_kernel|cor_prep|cor_scale|gemm_tn" -c 5 -f -o gpurun_out/r2_posthoc $C5CMD > gpurun_out/ncu_c5.log 2>&1
  echo "cfg5 ncu rc=$?"
  grep -h "^{" gpurun_out/ncu_null_plain.log gpurun_out/ncu_c5_plain.log | cut -c1-1500
else
  $LASSO > gpurun_out/ncu_lasso_plain.log 2>&1 &&
  ncu --set full --clock-control none --import-source on -k regex:"lasso_round_gram|gemm_tn" -s 62 -c 2 -f -o gpurun_out/r2_lasso $LASSO > gpurun_out/ncu_lasso.log 2>&1
  echo "lasso ncu rc=$?"
  $LASSO > gpurun_out/ncu_lasso_plain2.log 2>&1 &&
  ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2_launches_cfg3_block96.csv $LASSO > gpurun_out/ncu_launches.log 2>&1
  echo "launch list rc=$?"
fi
ls -la gpurun_out/*.ncu-rep; du -sh gpurun_out
