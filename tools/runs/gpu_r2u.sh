#!/bin/bash
mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_gpu_longctx.py tests/test_gpu_engine.py -q -x -p no:cacheprovider -k "wide_batch or layers_batched or batched_decode or mistral" ) > gpurun_out/r2u_tests.log 2>&1
echo "rc=$?" >> gpurun_out/r2u_tests.log
tail -4 gpurun_out/r2u_tests.log
L=gpurun_out/r2u_step_vs_b.log; : > $L
for sm in 0 -1; do
  echo "== CL_BATCH_ATTN_SMALL=$sm" >> $L
  CL_BATCH_ATTN_SMALL=$sm timeout 300 python tools/step_vs_b.py 256 8,16,32,64,128 >> $L 2>&1
  CL_BATCH_ATTN_SMALL=$sm timeout 300 python tools/step_vs_b.py 1024 16,32,64,128 >> $L 2>&1
done
cat $L
