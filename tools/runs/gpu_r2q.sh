#!/bin/bash
mkdir -p gpurun_out
L=gpurun_out/r2q_pre.log; : > $L
for pre in 4 6 8; do
  echo "== CL_GEMM_PRE=$pre ctx 1024" >> $L
  CL_GEMM_PRE=$pre timeout 200 python tools/step_vs_b.py 1024 2,8,32 >> $L 2>&1
done
echo "== CL_GEMM_PRE=8 ctx 256" >> $L
CL_GEMM_PRE=8 timeout 200 python tools/step_vs_b.py 256 8,32 >> $L 2>&1
cat $L
P=gpurun_out/r2q_prefill_small.log; : > $P
timeout 200 python tools/prefill_profile.py 146 2>&1 | tail -13 >> $P
echo "== wall, CL_SMALL_PDL=0" >> $P
timeout 200 python tools/prefill_wall.py 64,128,146,192,256 >> $P 2>&1
echo "== wall, CL_SMALL_PDL=1" >> $P
CL_SMALL_PDL=1 timeout 200 python tools/prefill_wall.py 64,128,146,192,256 >> $P 2>&1
echo "== wall, CL_SMALL_PDL=1 CL_GEMM_PRE=4" >> $P
CL_SMALL_PDL=1 CL_GEMM_PRE=4 timeout 200 python tools/prefill_wall.py 64,128,146,192,256 >> $P 2>&1
cat $P
