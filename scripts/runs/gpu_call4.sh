#!/bin/bash
mkdir -p gpurun_out
python scripts/probe_parity.py 24 > gpurun_out/r2_parity_probe.log 2>&1
cat gpurun_out/r2_parity_probe.log
for T in 512 256; do
  echo "== SAVER_GRAM_THREADS=$T" >> gpurun_out/r2_threads.log
  SAVER_GRAM_THREADS=$T python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); print('value', d['value'], 'ms/step', d['ms_per_step'], 'share', d['roofline']['share_of_gpu_time'])
    else: print(l.rstrip())
" >> gpurun_out/r2_threads.log 2>&1
done
cat gpurun_out/r2_threads.log
