#!/bin/bash
mkdir -p gpurun_out
nvidia-smi --query-gpu=clocks.sm,power.draw,clocks_event_reasons.sw_power_cap,clocks_event_reasons.active --format=csv,noheader -lms 100 > gpurun_out/r2d_clocks.csv &
SMI=$!
echo "== mt1 sustained" >> gpurun_out/r2d_clocks.csv
CL_GEMM_MT=1 timeout 300 python tools/gemm_bench.py 4096 300 > gpurun_out/r2d_gemm_mt1_sus.jsonl 2>&1
echo "== mt2 sustained" >> gpurun_out/r2d_clocks.csv
CL_GEMM_MT=2 timeout 300 python tools/gemm_bench.py 4096 300 > gpurun_out/r2d_gemm_mt2_sus.jsonl 2>&1
echo "== torch matmul sustained" >> gpurun_out/r2d_clocks.csv
timeout 300 python - > gpurun_out/r2d_torch.jsonl 2>&1 <<'PY'
import torch, json, time
for (T,n,k) in ((4096,6144,4096),(4096,28672,4096),(4096,4096,14336)):
    x=torch.randn(T,k,device='cuda',dtype=torch.bfloat16); w=torch.randn(n,k,device='cuda',dtype=torch.bfloat16)
    for it in (10,300):
        for _ in range(3): y=x@w.t()
        torch.cuda.synchronize(); e0=torch.cuda.Event(enable_timing=True); e1=torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(it): y=x@w.t()
        e1.record(); torch.cuda.synchronize()
        ms=e0.elapsed_time(e1)/it
        print(json.dumps({"lib":"cublas","T":T,"n":n,"k":k,"iters":it,"ms":round(ms,4),"tflops":round(2*T*n*k/ms/1e9,1)}),flush=True)
PY
kill $SMI
cat gpurun_out/r2d_gemm_mt1_sus.jsonl gpurun_out/r2d_gemm_mt2_sus.jsonl gpurun_out/r2d_torch.jsonl
python - <<'PY'
import re
sec=None; acc={}
for ln in open('gpurun_out/r2d_clocks.csv'):
    if ln.startswith('=='): sec=ln.strip(); acc[sec]=[]; continue
    if sec:
        f=ln.split(',')
        try: acc[sec].append((float(f[0].split()[0]), float(f[1].split()[0]), f[2].strip()))
        except Exception: pass
for k,v in acc.items():
    v=[x for x in v if x[1]>300]
    if v: print(k, "n",len(v),"sm_mhz median", sorted(x[0] for x in v)[len(v)//2], "power max", max(x[1] for x in v), "power_cap active", sum(1 for x in v if x[2].lower().startswith('active')))
PY
