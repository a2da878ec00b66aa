#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_gpu_posthoc.py -m gpu -q -x 2>&1 | tail -5
python bench.py --workload cfg5 --steps 3 --warmup 1 > gpurun_out/r2_bench_cfg5_b.json 2> gpurun_out/r2_bench_cfg5_b.err; echo "cfg5 rc=$?"; cut -c1-2500 gpurun_out/r2_bench_cfg5_b.json; tail -5 gpurun_out/r2_bench_cfg5_b.err
