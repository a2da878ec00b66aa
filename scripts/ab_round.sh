#!/bin/bash
# A/B runs of the Gram round kernel on a slice of the bench workload (scripts/diag_perf.py).
# SAVER_LASSO_DBG bits: 128 phase timers, 256 plain streaming loops, 1024 first-generation warp
# sweeps, 2048 problems taken in completion order, 4096 index-order sweep repeated in every
# strong-set phase; SAVER_GRAM_HS=0 disables the shared-memory packed Gram block;
# SAVER_GRAM_THREADS / SAVER_GRAM_CTAS / SAVER_GRAM_STAGED pick the CTA shape;
# SAVER_L2_PERSIST=0 drops the persisting-L2 window; SAVER_LASSO_PROBLEMS = problems in flight.
G=${1:-1024}
run() { echo "=== $1"; shift; env "$@" python scripts/diag_perf.py $G 2>&1 | grep -v Warning | grep -v "^  print" | tail -6 | grep "^genes\|phases\|cyc/step\|activations"; }
run "default" SAVER_LASSO_DBG=128
run "index-order sweep in every strong-set phase" SAVER_LASSO_DBG=4224
