#!/bin/bash
mkdir -p gpurun_out
timeout 800 python scripts/probe_variants.py 256 "SAVER_LASSO_ORDER=0" "SAVER_LASSO_ORDER=1" "SAVER_LASSO_ORDER=2" "SAVER_LASSO_ORDER=1,SAVER_L2_PERSIST=0" "SAVER_LASSO_ORDER=1,SAVER_LASSO_DBG=128" "SAVER_LASSO_ORDER=0,SAVER_LASSO_DBG=128" > gpurun_out/r2_order_ab.log 2>&1
echo rc=$? >> gpurun_out/r2_order_ab.log
cat gpurun_out/r2_order_ab.log
