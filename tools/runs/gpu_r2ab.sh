#!/bin/bash
mkdir -p gpurun_out
timeout 120 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -1
( CL_GEMM_MT_B128=1 timeout 600 python -m pytest tests/test_gpu_longctx.py -q -x -p no:cacheprovider -k "wide_batch" -s ) > gpurun_out/r2ab_tests.log 2>&1
echo "rc=$?" >> gpurun_out/r2ab_tests.log
grep -v "^\[oracle\]" gpurun_out/r2ab_tests.log | cut -c1-200 | tail -5
L=gpurun_out/r2ab_step_vs_b.log; : > $L
for m in 0 1; do
  echo "== CL_GEMM_MT_B128=$m" >> $L
  CL_GEMM_MT_B128=$m timeout 300 python tools/step_vs_b.py 256 80,100,128 >> $L 2>&1
done
cat $L
