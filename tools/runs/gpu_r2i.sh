#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/timeline_batch.py 8 1024 > gpurun_out/r2i_tl_b8.log 2>&1
CL_BMEGA_MAX_FLIGHT=0 CL_BMEGA_PAUSE=0 timeout 300 python tools/timeline_batch.py 8 1024 > gpurun_out/r2i_tl_b8_nocap.log 2>&1
timeout 300 python tools/timeline_batch.py 32 1024 > gpurun_out/r2i_tl_b32.log 2>&1
timeout 300 python tools/prefill_profile.py 4096 > gpurun_out/r2i_prof.log 2>&1
timeout 600 python bench.py --steps 20 --warmup 5 --no-box --no-cpu-baseline --no-extra-configs > gpurun_out/r2i_bench.json 2> gpurun_out/r2i_bench.err
cat gpurun_out/r2i_tl_b8.log; head -2 gpurun_out/r2i_tl_b8_nocap.log; tail -1 gpurun_out/r2i_tl_b8_nocap.log; head -2 gpurun_out/r2i_tl_b32.log; tail -1 gpurun_out/r2i_tl_b32.log
grep "prefill profile" gpurun_out/r2i_prof.log | tail -11 | head -1
python -c "
import json;d=json.loads(open('gpurun_out/r2i_bench.json').read().strip().splitlines()[-1]);print(d['value'], d['ms_per_step'], d['roofline']['frac'], d['roofline']['prefill'], d['clocks'])"
