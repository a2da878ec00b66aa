#!/bin/bash
mkdir -p gpurun_out
N=${1:-4}
( time python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29534 bench.py --gpus $N --steps 2 --warmup 1 ) > gpurun_out/r2_bench_cfg3_n$N.json 2> gpurun_out/r2_bench_cfg3_n$N.err; echo "cfg3 N=$N rc=$?"
python - <<PY
import json
for l in open('gpurun_out/r2_bench_cfg3_n$N.json'):
    if l.startswith('{'):
        d=json.loads(l)
        print('value',d['value'],'e2e',d['e2e'],'ms/step',d['ms_per_step'])
        print('roofline frac',d['roofline']['frac'],'share',d['roofline']['share_of_gpu_time'])
        print('bcast',d['design_broadcast'],'nranks',d['comm_nranks_seen'],'status',d['status_counts'])
PY
tail -4 gpurun_out/r2_bench_cfg3_n$N.err
