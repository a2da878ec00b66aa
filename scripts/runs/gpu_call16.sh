#!/bin/bash
mkdir -p gpurun_out
( time python -m pytest tests -m gpu -q ) > gpurun_out/r2_tests_c.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2_tests_c.log
tail -12 gpurun_out/r2_tests_c.log
( time python bench.py --steps 3 --warmup 1 --parity 4 ) > gpurun_out/r2_bench_d.json 2> gpurun_out/r2_bench_d.err
echo "bench rc=$?"
python - <<'PY'
import json
for l in open('gpurun_out/r2_bench_d.json'):
    if l.startswith('{'):
        d=json.loads(l)
        print('value',d['value'],'e2e',d['e2e']['value'],'ms/step',d['ms_per_step'], 'e2e s', d['e2e']['s_per_step'])
        print('roofline frac',d['roofline']['frac'],'share',d['roofline']['share_of_gpu_time'])
        print('parity',d.get('parity')); print('cpu',d.get('cpu_baseline'))
PY
tail -3 gpurun_out/r2_bench_d.err
python bench.py --workload fast --steps 1 --warmup 1 > gpurun_out/r2_bench_fast.json 2> gpurun_out/r2_bench_fast.err; echo "fast rc=$?"; cut -c1-900 gpurun_out/r2_bench_fast.json; tail -3 gpurun_out/r2_bench_fast.err
python bench.py --workload cfg5 --steps 3 --warmup 1 > gpurun_out/r2_bench_cfg5.json 2> gpurun_out/r2_bench_cfg5.err; echo "cfg5 rc=$?"; cut -c1-1600 gpurun_out/r2_bench_cfg5.json; tail -3 gpurun_out/r2_bench_cfg5.err
