#!/bin/bash
mkdir -p gpurun_out
echo "nproc $(nproc) cpu.max $(cat /sys/fs/cgroup/cpu.max 2>/dev/null)"
timeout 900 python bench.py --no-cpu-baseline > gpurun_out/r2t_bench.json 2> gpurun_out/r2t_bench.err
echo "bench rc=$?"; tail -3 gpurun_out/r2t_bench.err
python - <<PY
import json
d=json.loads(open('gpurun_out/r2t_bench.json').read().strip().splitlines()[-1])
print("value", d['value'], 'ms/step', d['ms_per_step'], 'roofline', d['roofline']['frac'], 'e2e', d['e2e']['value'])
b=d['box']
print(b.get('error'))
for k in ('config4','saturated'):
    print(k, {x: b[k][x] for x in ('concurrency','requests','ok','req_per_s','tok_per_s','p50_latency_s','wall_s','scheduler','gateway_processes')})
PY
timeout 300 python tools/sched_probe.py --max-batch 64 --conc 64,128 2>&1 | tee gpurun_out/r2t_probe64.log
CL_SCHED_BLOCKING_SYNC=0 timeout 300 python tools/sched_probe.py --max-batch 64 --conc 128 2>&1 | tee -a gpurun_out/r2t_probe64.log
timeout 300 python tools/sched_probe.py --max-batch 128 --conc 128,256 --waves 4 2>&1 | tee gpurun_out/r2t_probe128.log
