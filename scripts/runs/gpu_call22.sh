#!/bin/bash
mkdir -p gpurun_out
timeout 800 python scripts/probe_variants.py 256 "SAVER_LASSO_DBG=65536" "SAVER_LASSO_DBG=0" "SAVER_LASSO_DBG=128" > gpurun_out/r2_split_ab.log 2>&1
echo rc=$? >> gpurun_out/r2_split_ab.log
cat gpurun_out/r2_split_ab.log
