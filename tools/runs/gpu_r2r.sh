#!/bin/bash
mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_gpu_checkpoint.py tests/test_gpu_engine.py -q -x -p no:cacheprovider -k "checkpoint or golden or shapes or llama31" -s ) > gpurun_out/r2r_tests.log 2>&1
echo "rc=$?" >> gpurun_out/r2r_tests.log
tail -12 gpurun_out/r2r_tests.log
