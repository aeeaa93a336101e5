#!/bin/bash
mkdir -p gpurun_out
timeout 600 python bench.py > gpurun_out/last_bench.json 2> gpurun_out/last_bench.err
echo "rc=$?"
python - <<PY
import json
d=json.loads(open('gpurun_out/last_bench.json').read().strip().splitlines()[-1])
print("value", d['value'], 'ms/step', d['ms_per_step'], 'roofline', d['roofline']['frac'], 'e2e', d['e2e']['value'], 'cpu', d['cpu_baseline']['value'])
b=d['box']
print(b.get('error'))
for k in ('config4','saturated'):
    print(k, {x: b[k][x] for x in ('concurrency','requests','ok','req_per_s','tok_per_s','p50_latency_s')})
PY
