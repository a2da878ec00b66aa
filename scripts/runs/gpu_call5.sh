#!/bin/bash
mkdir -p gpurun_out
run() {
  echo "== $*" >> gpurun_out/r2_occ.log
  env "$@" python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); print('value', d['value'], 'ms/step', d['ms_per_step'], 'share', d['roofline']['share_of_gpu_time'])
    else: print(l.rstrip())
" >> gpurun_out/r2_occ.log 2>&1
}
run SAVER_GRAM_THREADS=256 SAVER_GRAM_STAGED=0
run SAVER_GRAM_THREADS=128 SAVER_GRAM_STAGED=0
run SAVER_GRAM_THREADS=512 SAVER_GRAM_STAGED=0
cat gpurun_out/r2_occ.log
