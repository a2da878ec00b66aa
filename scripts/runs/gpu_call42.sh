#!/bin/bash
mkdir -p gpurun_out
timeout 150 python bench.py --steps 2 --warmup 1 --no-cpu > gpurun_out/r2_bench_h.json 2> gpurun_out/r2_bench_h.err
echo "bench rc=$?"
python - <<'PY'
import json
for l in open('gpurun_out/r2_bench_h.json'):
    if l.startswith('{'):
        d=json.loads(l)
        print('value',d['value'],'e2e',d['e2e']['value'],'ms/step',d['ms_per_step'], 'e2e s', d['e2e']['s_per_step'])
        print('roofline frac',d['roofline']['frac'],'share',d['roofline']['share_of_gpu_time'], 'gemm', d['roofline']['kernels']['gemm_tn_f64_kernel'])
PY
tail -3 gpurun_out/r2_bench_h.err
