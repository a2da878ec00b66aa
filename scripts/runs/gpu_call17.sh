#!/bin/bash
mkdir -p gpurun_out
for sl in 0.25 0.0; do
  echo "== strong_slack $sl" >> gpurun_out/r2_fast_slack.log
  SAVER_STRONG_SLACK=$sl python bench.py --workload fast --steps 1 --warmup 0 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); print('value', d['value'], 's', d['ms_per_step']/1e3, d['config']['workload'][-120:], 'launches', d['gpu_launches'])
    else: print(l.rstrip())
" >> gpurun_out/r2_fast_slack.log 2>&1
done
cat gpurun_out/r2_fast_slack.log
