#!/bin/bash
mkdir -p gpurun_out
( time python -m pytest tests -m gpu -q ) > gpurun_out/r2_tests_b.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2_tests_b.log
tail -25 gpurun_out/r2_tests_b.log
( time python bench.py --steps 2 --warmup 1 --parity 4 ) > gpurun_out/r2_bench_b.json 2> gpurun_out/r2_bench_b.err
echo "bench rc=$?"
python - <<'PY'
import json
for l in open('gpurun_out/r2_bench_b.json'):
    if l.startswith('{'):
        d=json.loads(l)
        print('value',d['value'],'e2e',d['e2e']['value'],'ms/step',d['ms_per_step'])
        print('roofline frac',d['roofline']['frac'],'share',d['roofline']['share_of_gpu_time'])
        print('parity',d.get('parity')); print('cpu',d.get('cpu_baseline'))
PY
tail -5 gpurun_out/r2_bench_b.err
