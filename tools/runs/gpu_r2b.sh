#!/bin/bash
# round-2 GPU pass B: tcgen05 prefill attention — parity, microbench, engine prefill
mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_gpu_prefill.py -q -x -p no:cacheprovider -k "attn_prefill" ) > gpurun_out/r2b_tests_attn.log 2>&1
echo "rc=$?" >> gpurun_out/r2b_tests_attn.log
timeout 300 python tools/attn_bench.py 512 1024 4096 8192 > gpurun_out/r2b_attn_bench.jsonl 2> gpurun_out/r2b_attn_bench.err
( timeout 900 python -m pytest tests/test_gpu_prefill.py tests/test_gpu_longctx.py -q -x -s -p no:cacheprovider -k "chunked or paths_agree or prefill_4096" ) > gpurun_out/r2b_tests_engine.log 2>&1
echo "rc=$?" >> gpurun_out/r2b_tests_engine.log
timeout 600 python bench.py --steps 32 --warmup 5 --no-box --no-cpu-baseline --no-extra-configs > gpurun_out/r2b_bench.json 2> gpurun_out/r2b_bench.err
tail -15 gpurun_out/r2b_tests_attn.log; cat gpurun_out/r2b_attn_bench.jsonl; tail -3 gpurun_out/r2b_attn_bench.err; tail -8 gpurun_out/r2b_tests_engine.log; python -c "
import json;d=json.loads(open('gpurun_out/r2b_bench.json').read().strip().splitlines()[-1]);print(d['value'], d['roofline']['prefill'])"
