#!/bin/bash
mkdir -p gpurun_out
python scripts/probe_parity.py 24 2>&1 | head -8 > gpurun_out/r2_parity_probe_c.log
cat gpurun_out/r2_parity_probe_c.log
( SAVER_LASSO_DBG=128 python bench.py --steps 2 --warmup 1 --parity 4 --no-e2e ) > gpurun_out/r2_bench_c.json 2> gpurun_out/r2_bench_c.err
echo "bench rc=$?"
python - <<'PY'
import json
for l in open('gpurun_out/r2_bench_c.json'):
    if l.startswith('{'):
        d=json.loads(l)
        print('value',d['value'],'ms/step',d['ms_per_step'])
        print('roofline frac',d['roofline']['frac'],'share',d['roofline']['share_of_gpu_time'])
        print('parity',d.get('parity'))
        print('phases', d.get('phases'))
PY
tail -5 gpurun_out/r2_bench_c.err
