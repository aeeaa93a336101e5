#!/bin/bash
mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_gpu_longctx.py -q -x -p no:cacheprovider -k "wide_batch" -s ) > gpurun_out/r2s_tests.log 2>&1
echo "rc=$?" >> gpurun_out/r2s_tests.log
tail -6 gpurun_out/r2s_tests.log
timeout 300 python tools/step_vs_b.py 256 16,32,48,64,96,128 2>&1 | tee gpurun_out/r2s_step_vs_b.log
timeout 300 python tools/step_vs_b.py 1024 32,64,128 2>&1 | tee -a gpurun_out/r2s_step_vs_b.log
