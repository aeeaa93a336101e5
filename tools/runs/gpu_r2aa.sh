#!/bin/bash
mkdir -p gpurun_out
CL_BOX_MAX_BATCH=128 timeout 900 python bench.py --no-cpu-baseline > gpurun_out/r2aa_bench_mb128_full.json 2> gpurun_out/r2aa_bench.err
python - <<PY
import json
d=json.loads(open('gpurun_out/r2aa_bench_mb128_full.json').read().strip().splitlines()[-1])
print("value", d['value'], 'ms/step', d['ms_per_step'], 'roofline', d['roofline']['frac'], 'e2e', d['e2e'], 'extra', d['extra'])
b=d['box']
print(b.get('error'))
for k in ('config4','saturated'):
    print(k, {x: b[k][x] for x in ('concurrency','requests','ok','req_per_s','tok_per_s','p50_latency_s')})
PY
