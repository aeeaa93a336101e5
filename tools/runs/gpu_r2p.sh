#!/bin/bash
mkdir -p gpurun_out
L=gpurun_out/r2p_pre.log; : > $L
for pre in 0 1 2 4; do
  echo "== CL_GEMM_PRE=$pre ctx 1024" >> $L
  CL_GEMM_PRE=$pre timeout 200 python tools/step_vs_b.py 1024 4,8,16,32 >> $L 2>&1
done
for pre in 0 2; do
  echo "== CL_GEMM_PRE=$pre ctx 256" >> $L
  CL_GEMM_PRE=$pre timeout 200 python tools/step_vs_b.py 256 8,32 >> $L 2>&1
done
echo "== CL_BATCH_GEMM_MIN=2 ctx 256 / 4096" >> $L
CL_BATCH_GEMM_MIN=2 timeout 200 python tools/step_vs_b.py 256 2,3 >> $L 2>&1
CL_BATCH_GEMM_MIN=2 timeout 200 python tools/step_vs_b.py 4096 2 >> $L 2>&1
timeout 200 python tools/step_vs_b.py 4096 2 >> $L 2>&1
cat $L
timeout 200 python tools/prefill_profile.py 146 2>&1 | tail -14 > gpurun_out/r2p_prefill146.log
cat gpurun_out/r2p_prefill146.log
