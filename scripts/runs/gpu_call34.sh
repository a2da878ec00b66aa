#!/bin/bash
mkdir -p gpurun_out
timeout 500 python scripts/probe_variants.py 256 "SAVER_LASSO_DBG=131200" "SAVER_LASSO_DBG=128" > gpurun_out/r2_defer_ab.log 2>&1
cat gpurun_out/r2_defer_ab.log
timeout 100 python scripts/probe_heavy.py "" >> gpurun_out/r2_defer_ab.log 2>&1; tail -1 gpurun_out/r2_defer_ab.log
