#!/bin/bash
mkdir -p gpurun_out
( timeout 900 python -m pytest tests -q -x -m gpu -p no:cacheprovider ) > gpurun_out/r2l_tests.log 2>&1
echo "rc=$?" >> gpurun_out/r2l_tests.log
tail -4 gpurun_out/r2l_tests.log
timeout 900 python bench.py > gpurun_out/r2l_bench.json 2> gpurun_out/r2l_bench.err
echo "bench rc=$?"; tail -3 gpurun_out/r2l_bench.err
python - <<PY
import json
d=json.loads(open('gpurun_out/r2l_bench.json').read().strip().splitlines()[-1])
print("value", d['value'], 'ms/step', d['ms_per_step'], 'roofline', d['roofline']['frac'], 'e2e', d['e2e']['value'])
b=d['box']
for k in ('config4','saturated'):
    print(k, {x: b[k][x] for x in ('concurrency','requests','ok','req_per_s','tok_per_s','p50_latency_s','wall_s')})
PY
