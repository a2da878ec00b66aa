#!/bin/bash
mkdir -p gpurun_out
( time timeout 1400 python bench.py --workload cfg4 --steps 1 ) > gpurun_out/r2_bench_cfg4.json 2> gpurun_out/r2_bench_cfg4.err; echo "cfg4 rc=$?"; cut -c1-1800 gpurun_out/r2_bench_cfg4.json; tail -12 gpurun_out/r2_bench_cfg4.err; free -g | head -2; nvidia-smi --query-gpu=memory.used --format=csv
