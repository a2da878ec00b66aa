#!/bin/bash
mkdir -p gpurun_out
timeout 800 python scripts/probe_variants.py 256 "SAVER_LASSO_DBG=32768" "SAVER_LASSO_DBG=0" "SAVER_LASSO_DBG=128" > gpurun_out/r2_screen_ab.log 2>&1
echo rc=$? >> gpurun_out/r2_screen_ab.log
cat gpurun_out/r2_screen_ab.log
