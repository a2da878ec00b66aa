#!/bin/bash
mkdir -p gpurun_out
timeout 800 python scripts/probe_variants.py 256 "SAVER_STRONG_SLACK=1.0" "SAVER_STRONG_SLACK=0.5,SAVER_LASSO_DBG=128" "SAVER_STRONG_SLACK=0.25,SAVER_LASSO_DBG=128" "SAVER_STRONG_SLACK=0.0,SAVER_LASSO_DBG=128" > gpurun_out/r2_slack_ab.log 2>&1
echo rc=$? >> gpurun_out/r2_slack_ab.log
cat gpurun_out/r2_slack_ab.log
