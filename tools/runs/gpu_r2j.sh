#!/bin/bash
mkdir -p gpurun_out
( timeout 600 python -m pytest tests/test_gpu_longctx.py -q -x -p no:cacheprovider -k "layers_batched" ) > gpurun_out/r2j_tests.log 2>&1
echo "rc=$?" >> gpurun_out/r2j_tests.log
timeout 300 python tools/timeline_batch.py 8 1024 > gpurun_out/r2j_tl_b8.log 2>&1
CL_BMEGA_MAX_FLIGHT=0 CL_BMEGA_PAUSE=0 timeout 300 python tools/timeline_batch.py 8 1024 > gpurun_out/r2j_tl_b8_nocap.log 2>&1
CL_BMEGA_MAX_FLIGHT=6 timeout 300 python tools/timeline_batch.py 8 1024 > gpurun_out/r2j_tl_b8_f6.log 2>&1
timeout 300 python tools/timeline_batch.py 32 1024 > gpurun_out/r2j_tl_b32.log 2>&1
tail -3 gpurun_out/r2j_tests.log
cat gpurun_out/r2j_tl_b8.log; head -2 gpurun_out/r2j_tl_b8_nocap.log; tail -1 gpurun_out/r2j_tl_b8_nocap.log; head -2 gpurun_out/r2j_tl_b8_f6.log; tail -1 gpurun_out/r2j_tl_b8_f6.log; head -2 gpurun_out/r2j_tl_b32.log; tail -1 gpurun_out/r2j_tl_b32.log
