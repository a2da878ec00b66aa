#!/bin/bash
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader > gpurun_out/r2_n4.log 2>&1
( time python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 4 --steps 2 --warmup 1 ) > gpurun_out/r2_bench_n4.json 2> gpurun_out/r2_bench_n4.err
echo "bench n4 rc=$?" | tee -a gpurun_out/r2_n4.log
tail -c 3000 gpurun_out/r2_bench_n4.json; grep -v "^$" gpurun_out/r2_bench_n4.err | tail -40
ls gpurun_out/bench_rank*.err 2>/dev/null && cat gpurun_out/bench_rank*.err | tail -60
