#!/bin/bash
mkdir -p gpurun_out
( time python -m pytest tests -m gpu -x -q ) > gpurun_out/r2_tests_a.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2_tests_a.log
tail -15 gpurun_out/r2_tests_a.log
( time python bench.py --steps 2 --warmup 1 ) > gpurun_out/r2_bench_a.json 2> gpurun_out/r2_bench_a.err
echo "bench rc=$?"
tail -c 6000 gpurun_out/r2_bench_a.json; tail -20 gpurun_out/r2_bench_a.err
