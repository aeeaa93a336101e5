#!/bin/bash
# usage: bash tools/gpu_scale.sh N   (run under gpurun --gpus N): bench.py with N worker peers incl. the box leg
N=$1
mkdir -p gpurun_out
echo "nproc $(nproc) cpu.max $(cat /sys/fs/cgroup/cpu.max 2>/dev/null)"
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 295$((20+N)) bench.py --gpus $N --steps 20 --warmup 5 \
   > gpurun_out/r2k_bench_n$N.json 2> gpurun_out/r2k_bench_n$N.err
echo "rc=$?"
tail -3 gpurun_out/r2k_bench_n$N.err
python - <<PY
import json
d=json.loads(open('gpurun_out/r2k_bench_n$N.json').read().strip().splitlines()[-1])
print("N=$N value", d['value'], 'tok/s;', 'ms/step', d['ms_per_step'])
b=d['box']
for k in ('config4','saturated'):
    print(k, {x: b[k][x] for x in ('concurrency','requests','ok','req_per_s','tok_per_s','p50_latency_s','per_worker_requests','errors','scheduler')})
PY
