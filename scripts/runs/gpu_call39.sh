#!/bin/bash
mkdir -p gpurun_out
( time python -m pytest tests -m gpu -q ) > gpurun_out/r2_tests_final.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2_tests_final.log
tail -8 gpurun_out/r2_tests_final.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_smoke_final.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/r2_smoke_final.log
python bench.py --workload null --steps 3 --warmup 1 > gpurun_out/r2_bench_null.json 2> gpurun_out/r2_bench_null.err; echo "null rc=$?"; cut -c1-700 gpurun_out/r2_bench_null.json
python bench.py --workload cfg2 --steps 2 --warmup 1 --parity 8 > gpurun_out/r2_bench_cfg2.json 2> gpurun_out/r2_bench_cfg2.err; echo "cfg2 rc=$?"
python - <<'PY'
import json
for l in open('gpurun_out/r2_bench_cfg2.json'):
    if l.startswith('{'):
        d=json.loads(l)
        print('cfg2 value',d['value'],'e2e',d['e2e']['value'],'ms/step',d['ms_per_step'])
        print('roofline frac',d['roofline']['frac'],'share',d['roofline']['share_of_gpu_time'])
        print('parity',d.get('parity'))
PY
python bench.py --workload fast --steps 1 --warmup 1 > gpurun_out/r2_bench_fast.json 2> gpurun_out/r2_bench_fast.err; echo "fast rc=$?"; cut -c1-500 gpurun_out/r2_bench_fast.json
