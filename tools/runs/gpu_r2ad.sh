#!/bin/bash
mkdir -p gpurun_out
( timeout 600 python -m pytest tests/test_gpu_longctx.py tests/test_gpu_engine.py -q -x -p no:cacheprovider -k "wide_batch or layers_batched or batched_decode" ) > gpurun_out/r2ad_tests.log 2>&1
echo "rc=$?" >> gpurun_out/r2ad_tests.log
tail -3 gpurun_out/r2ad_tests.log
( CL_BATCH_ATTN_PERSIST=1 timeout 600 python -m pytest tests/test_gpu_longctx.py -q -x -p no:cacheprovider -k "wide_batch" ) > gpurun_out/r2ad_tests_persist.log 2>&1
echo "rc=$?" >> gpurun_out/r2ad_tests_persist.log
tail -3 gpurun_out/r2ad_tests_persist.log
L=gpurun_out/r2ad_step_vs_b.log; : > $L
for m in 0 1; do
  echo "== CL_BATCH_ATTN_PERSIST=$m" >> $L
  CL_BATCH_ATTN_PERSIST=$m timeout 200 python tools/step_vs_b.py 256 48,64,128 >> $L 2>&1
  CL_BATCH_ATTN_PERSIST=$m timeout 200 python tools/step_vs_b.py 1024 64,128 >> $L 2>&1
done
cat $L
