#!/bin/bash
mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_gpu_prefill.py tests/test_gpu_longctx.py -q -x -p no:cacheprovider -k "gemm_tcgen05 or chunked or paths_agree or unaligned or prefill_4096" ) > gpurun_out/r2g_tests.log 2>&1
echo "rc=$?" >> gpurun_out/r2g_tests.log
CL_PREFILL_FUSED=0 timeout 300 python tools/prefill_profile.py 4096 > gpurun_out/r2g_prof_unfused.log 2>&1
CL_PREFILL_FUSED=1 timeout 300 python tools/prefill_profile.py 4096 > gpurun_out/r2g_prof_fused.log 2>&1
timeout 300 python tools/batch_step_profile.py 8 1024 > gpurun_out/r2g_step_b8.log 2>&1
timeout 300 python tools/batch_step_profile.py 32 1024 > gpurun_out/r2g_step_b32.log 2>&1
timeout 300 python tools/timeline_batch.py 8 1024 > gpurun_out/r2g_tl_b8.log 2>&1
timeout 300 python tools/timeline_batch.py 32 1024 > gpurun_out/r2g_tl_b32.log 2>&1
tail -4 gpurun_out/r2g_tests.log
grep "prefill profile" gpurun_out/r2g_prof_unfused.log | tail -12; grep "prefill profile" gpurun_out/r2g_prof_fused.log | tail -11
tail -16 gpurun_out/r2g_step_b8.log; tail -16 gpurun_out/r2g_step_b32.log
cat gpurun_out/r2g_tl_b8.log; cat gpurun_out/r2g_tl_b32.log
