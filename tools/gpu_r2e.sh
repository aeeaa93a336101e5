#!/bin/bash
mkdir -p gpurun_out
CL_GEMM_MT=1 timeout 300 python tools/prefill_profile.py 4096 > gpurun_out/r2e_prof_mt1.log 2>&1
CL_GEMM_MT=2 timeout 300 python tools/prefill_profile.py 4096 > gpurun_out/r2e_prof_mt2.log 2>&1
CL_GEMM_MT=1 timeout 300 python tools/prefill_profile.py 128 > gpurun_out/r2e_prof_128.log 2>&1
tail -12 gpurun_out/r2e_prof_mt1.log; tail -12 gpurun_out/r2e_prof_mt2.log; tail -12 gpurun_out/r2e_prof_128.log
