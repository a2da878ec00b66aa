#!/bin/bash
mkdir -p gpurun_out
timeout 120 python scripts/gemm_probe.py > gpurun_out/r2_gemm_probe.log 2>&1; cat gpurun_out/r2_gemm_probe.log
timeout 200 python scripts/probe_variants.py 256 "" 2>&1 | tail -2 | tee -a gpurun_out/r2_gemm_probe.log
( time python -m pytest tests -m gpu -q -x ) > gpurun_out/r2_tests_gemm.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2_tests_gemm.log
tail -6 gpurun_out/r2_tests_gemm.log
