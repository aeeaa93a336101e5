#!/bin/bash
mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_gpu_prefill.py tests/test_gpu_longctx.py tests/test_gpu_engine.py -q -x -p no:cacheprovider -k "chunked or paths_agree or prefill_4096 or hf_golden or tiny_engine" ) > gpurun_out/r2e_tests.log 2>&1
echo "rc=$?" >> gpurun_out/r2e_tests.log
CL_GEMM_MT=1 CL_PREFILL_FUSED=0 timeout 300 python tools/prefill_profile.py 4096 > gpurun_out/r2e_prof_unfused.log 2>&1
CL_GEMM_MT=1 timeout 300 python tools/prefill_profile.py 4096 > gpurun_out/r2e_prof_fused.log 2>&1
CL_GEMM_MT=1 timeout 300 python tools/prefill_profile.py 128 > gpurun_out/r2e_prof_128.log 2>&1
timeout 300 python tools/batch_step_profile.py 8 1024 > gpurun_out/r2e_step_b8.log 2>&1
timeout 300 python tools/batch_step_profile.py 32 1024 > gpurun_out/r2e_step_b32.log 2>&1
tail -4 gpurun_out/r2e_tests.log
tail -11 gpurun_out/r2e_prof_unfused.log; tail -9 gpurun_out/r2e_prof_fused.log; tail -9 gpurun_out/r2e_prof_128.log
grep graph gpurun_out/r2e_step_b8.log; tail -13 gpurun_out/r2e_step_b8.log; grep graph gpurun_out/r2e_step_b32.log; tail -13 gpurun_out/r2e_step_b32.log
