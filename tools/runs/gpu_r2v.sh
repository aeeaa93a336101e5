#!/bin/bash
mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_gpu_prefill.py tests/test_gpu_engine.py -q -x -p no:cacheprovider -k "multi_prompt or group_admission" -s ) > gpurun_out/r2v_tests.log 2>&1
echo "rc=$?" >> gpurun_out/r2v_tests.log
tail -25 gpurun_out/r2v_tests.log
