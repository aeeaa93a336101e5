#!/bin/bash
mkdir -p gpurun_out
CL_BOX_SCENARIOS="saturated,config4,c33,c48,saturated" timeout 900 python bench.py --no-cpu-baseline --no-extra-configs --steps 32 > gpurun_out/r2o_bench.json 2> gpurun_out/r2o_bench.err
echo "bench rc=$?"; tail -3 gpurun_out/r2o_bench.err
python - <<PY
import json
d=json.loads(open('gpurun_out/r2o_bench.json').read().strip().splitlines()[-1])
b=d['box']
for k in b:
    if isinstance(b[k], dict) and 'req_per_s' in b[k]:
        print(k, {x: b[k][x] for x in ('concurrency','requests','ok','req_per_s','p50_latency_s','wall_s','scheduler')})
PY
