#!/bin/bash
mkdir -p gpurun_out
( time python -m pytest tests -m gpu -q --deselect tests/test_gpu_lasso.py --deselect tests/test_gpu_maxcor.py --deselect tests/test_gpu_posthoc.py --deselect tests/test_gpu_properties.py --deselect tests/test_gpu_saver.py --deselect tests/test_gpu_shapes.py ) > gpurun_out/r2_tests_e.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2_tests_e.log
tail -15 gpurun_out/r2_tests_e.log
