#!/bin/bash
mkdir -p gpurun_out
( time CFG4_G=3000 CFG4_N=4000 python bench.py --workload cfg4 --block 256 --steps 1 ) > gpurun_out/r2_cfg4_smoke.json 2> gpurun_out/r2_cfg4_smoke.err; echo "smoke rc=$?"; cut -c1-1500 gpurun_out/r2_cfg4_smoke.json; tail -8 gpurun_out/r2_cfg4_smoke.err
