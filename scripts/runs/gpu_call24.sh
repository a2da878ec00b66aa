#!/bin/bash
mkdir -p gpurun_out
( time python -m pytest tests -m gpu -q -x ) > gpurun_out/r2_tests_d.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2_tests_d.log
tail -15 gpurun_out/r2_tests_d.log
( time python bench.py --steps 2 --warmup 1 --parity 4 ) > gpurun_out/r2_bench_e.json 2> gpurun_out/r2_bench_e.err
echo "bench rc=$?"
python - <<'PY'
import json
for l in open('gpurun_out/r2_bench_e.json'):
    if l.startswith('{'):
        d=json.loads(l)
        print('value',d['value'],'e2e',d['e2e']['value'],'ms/step',d['ms_per_step'], 'e2e s', d['e2e']['s_per_step'])
        print('roofline frac',d['roofline']['frac'],'share',d['roofline']['share_of_gpu_time'])
        print('parity',d.get('parity')); print('cpu',d.get('cpu_baseline'))
PY
tail -3 gpurun_out/r2_bench_e.err
