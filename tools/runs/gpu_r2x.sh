#!/bin/bash
mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_gpu_engine.py tests/test_gpu_fullsize.py -q -x -p no:cacheprovider ) > gpurun_out/r2x_tests.log 2>&1
echo "rc=$?" >> gpurun_out/r2x_tests.log
tail -3 gpurun_out/r2x_tests.log
