#!/bin/bash
mkdir -p gpurun_out
run() {
  echo "== $*" >> gpurun_out/r2_regs.log
  env "$@" python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); print('value', d['value'], 'ms/step', d['ms_per_step'], 'share', d['roofline']['share_of_gpu_time']); print(' phases', d.get('phases'))
    else: print(l.rstrip())
" >> gpurun_out/r2_regs.log 2>&1
}
run SAVER_X=1
run SAVER_LASSO_DBG=128
run SAVER_GRAM_THREADS=512
cat gpurun_out/r2_regs.log
python bench.py --workload null --steps 3 --warmup 1 2>&1 | tail -1 | cut -c1-1200
