#!/bin/bash
mkdir -p gpurun_out
timeout 400 python scripts/probe_fastpath.py 1200 2000 2>&1 | tail -22 > gpurun_out/r2_fastpath_after.log; cat gpurun_out/r2_fastpath_after.log
python bench.py --workload fast --steps 1 --warmup 1 > gpurun_out/r2_bench_fast_b.json 2> gpurun_out/r2_bench_fast_b.err; echo "fast rc=$?"; cut -c1-1000 gpurun_out/r2_bench_fast_b.json; tail -3 gpurun_out/r2_bench_fast_b.err
