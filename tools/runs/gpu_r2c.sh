#!/bin/bash
mkdir -p gpurun_out
( timeout 600 python -m pytest tests/test_gpu_prefill.py -q -x -p no:cacheprovider -k "gemm_tcgen05 or chunked or paths_agree" ) > gpurun_out/r2c_tests.log 2>&1
echo "rc=$?" >> gpurun_out/r2c_tests.log
CL_GEMM_MT=1 timeout 300 python tools/gemm_bench.py 4096 > gpurun_out/r2c_gemm_mt1.jsonl 2>&1
CL_GEMM_MT=2 timeout 300 python tools/gemm_bench.py 4096 > gpurun_out/r2c_gemm_mt2.jsonl 2>&1
CL_GEMM_MT=2 timeout 300 python tools/gemm_bench.py 1024 > gpurun_out/r2c_gemm_mt2_1024.jsonl 2>&1
CL_GEMM_MT=1 timeout 300 python tools/gemm_bench.py 1024 > gpurun_out/r2c_gemm_mt1_1024.jsonl 2>&1
timeout 600 python bench.py --steps 32 --warmup 5 --no-box --no-cpu-baseline --no-extra-configs > gpurun_out/r2c_bench.json 2> gpurun_out/r2c_bench.err
tail -4 gpurun_out/r2c_tests.log; cat gpurun_out/r2c_gemm_mt1.jsonl gpurun_out/r2c_gemm_mt2.jsonl; tail -1 gpurun_out/r2c_gemm_mt1_1024.jsonl gpurun_out/r2c_gemm_mt2_1024.jsonl; python -c "
import json;d=json.loads(open('gpurun_out/r2c_bench.json').read().strip().splitlines()[-1]);print(d['value'], d['roofline']['prefill'])"
