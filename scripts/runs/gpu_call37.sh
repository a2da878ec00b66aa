#!/bin/bash
mkdir -p gpurun_out
timeout 600 python scripts/probe_variants.py 256 "SAVER_LASSO_DBG=128" "SAVER_LASSO_DBG=262272" "SAVER_LASSO_DBG=128" "SAVER_LASSO_DBG=262272" "SAVER_LASSO_DBG=262144" "SAVER_LASSO_DBG=0" > gpurun_out/r2_lazy_ab2.log 2>&1
cat gpurun_out/r2_lazy_ab2.log
