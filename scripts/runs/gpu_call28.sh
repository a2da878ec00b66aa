#!/bin/bash
mkdir -p gpurun_out
timeout 300 python scripts/probe_heavy.py "" "SAVER_CHORD_TOL=1e2" "SAVER_CHORD_TOL=1e6" > gpurun_out/r2_chord_heavy.log 2>&1
cat gpurun_out/r2_chord_heavy.log | tail -5
timeout 500 python scripts/probe_variants.py 256 "SAVER_CHORD_TOL=1e30,SAVER_LASSO_DBG=128" "SAVER_LASSO_DBG=128" "SAVER_CHORD_TOL=1e2,SAVER_LASSO_DBG=128" > gpurun_out/r2_chord_cfg3.log 2>&1
cat gpurun_out/r2_chord_cfg3.log
