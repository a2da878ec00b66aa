#!/bin/bash
mkdir -p gpurun_out
run() {
  echo "== $*" >> gpurun_out/r2_ab.log
  env "$@" python bench.py --steps 2 --warmup 1 --no-e2e --parity 4 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); print('value', d['value'], 'ms/step', d['ms_per_step'], 'share', d['roofline']['share_of_gpu_time']); print(' parity', d['parity']['median_rel_err_estimate'], d['parity']['per_gene_median_rel_err'], d['parity']['lambda_min_equal']); print(' phases', d.get('phases'))
    else: print(l.rstrip())
" >> gpurun_out/r2_ab.log 2>&1
}
run SAVER_X=1
run SAVER_LASSO_DBG=16384
run SAVER_LASSO_DBG=128
cat gpurun_out/r2_ab.log
