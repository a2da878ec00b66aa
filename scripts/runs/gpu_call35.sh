#!/bin/bash
mkdir -p gpurun_out
( time python -m pytest tests/test_gpu_lasso.py tests/test_gpu_shapes.py tests/test_gpu_saver.py tests/test_zz_gpu_golden_cv.py tests/test_gpu_sparse.py -m gpu -q ) > gpurun_out/r2_tests_g.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2_tests_g.log
tail -12 gpurun_out/r2_tests_g.log
( time python bench.py --steps 2 --warmup 1 --parity 6 ) > gpurun_out/r2_bench_g.json 2> gpurun_out/r2_bench_g.err
echo "bench rc=$?"
python - <<'PY'
import json
for l in open('gpurun_out/r2_bench_g.json'):
    if l.startswith('{'):
        d=json.loads(l)
        print('value',d['value'],'e2e',d['e2e']['value'],'ms/step',d['ms_per_step'], 'e2e s', d['e2e']['s_per_step'])
        print('roofline frac',d['roofline']['frac'],'share',d['roofline']['share_of_gpu_time'])
        print('parity',d.get('parity')); print('cpu',d.get('cpu_baseline'))
PY
tail -3 gpurun_out/r2_bench_g.err
