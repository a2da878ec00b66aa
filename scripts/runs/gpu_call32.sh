#!/bin/bash
mkdir -p gpurun_out
timeout 500 python scripts/probe_variants.py 256 "SAVER_LASSO_DBG=65536" "SAVER_LASSO_DBG=0" "SAVER_LASSO_DBG=128" > gpurun_out/r2_gramsplit_ab.log 2>&1
cat gpurun_out/r2_gramsplit_ab.log
