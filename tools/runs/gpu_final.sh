#!/bin/bash
mkdir -p gpurun_out
( timeout 1200 python -m pytest tests -x -q -m gpu -p no:cacheprovider ) > gpurun_out/final_tests.log 2>&1
echo "rc=$?" >> gpurun_out/final_tests.log
tail -4 gpurun_out/final_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 900 python bench.py > gpurun_out/final_bench.json 2> gpurun_out/final_bench.err
echo "bench rc=$?"; tail -2 gpurun_out/final_bench.err
timeout 900 python bench.py --impl reference --steps 40 --warmup 2 > gpurun_out/final_bench_reference.json 2> gpurun_out/final_bench_reference.err
echo "reference rc=$?"; tail -c 600 gpurun_out/final_bench_reference.json
CL_BOX_MAX_BATCH=128 timeout 900 python bench.py --no-cpu-baseline --no-extra-configs --steps 32 > gpurun_out/final_bench_mb128.json 2> gpurun_out/final_bench_mb128.err
python - <<PY
import json
for f in ('final_bench','final_bench_mb128'):
    d=json.loads(open('gpurun_out/%s.json'%f).read().strip().splitlines()[-1])
    print(f, "value", d['value'], 'ms/step', d['ms_per_step'], 'roofline', d['roofline']['frac'], 'e2e', d['e2e']['value'], 'clocks', d['clocks'])
    b=d['box']
    print(b.get('error'))
    for k in ('config4','saturated'):
        print(k, {x: b[k][x] for x in ('concurrency','requests','ok','req_per_s','tok_per_s','p50_latency_s','scheduler')})
PY
