#!/bin/bash
mkdir -p gpurun_out
rm -f gpurun_out/parity_longctx.jsonl
( time timeout 1700 python -m pytest tests -m gpu -q -s -x -p no:cacheprovider ) > gpurun_out/r2h_tests.log 2>&1
echo "tests rc=$?" >> gpurun_out/r2h_tests.log
CL_PREFILL_FUSED=0 timeout 300 python tools/prefill_profile.py 4096 > gpurun_out/r2h_prof_unfused.log 2>&1
CL_PREFILL_FUSED=1 timeout 300 python tools/prefill_profile.py 4096 > gpurun_out/r2h_prof_silu.log 2>&1
timeout 300 python tools/timeline_batch.py 8 1024 > gpurun_out/r2h_tl_b8.log 2>&1
timeout 300 python tools/timeline_batch.py 32 1024 > gpurun_out/r2h_tl_b32.log 2>&1
CL_BMEGA_MAX_FLIGHT=0 CL_BMEGA_PAUSE=0 timeout 300 python tools/timeline_batch.py 8 1024 > gpurun_out/r2h_tl_b8_nocap.log 2>&1
timeout 900 python bench.py --steps 64 --warmup 5 > gpurun_out/r2h_bench.json 2> gpurun_out/r2h_bench.err
grep -c "passed\|PASSED" gpurun_out/r2h_tests.log; tail -6 gpurun_out/r2h_tests.log
grep "prefill profile" gpurun_out/r2h_prof_unfused.log | tail -12 | head -1; grep "prefill profile" gpurun_out/r2h_prof_silu.log | tail -11
cat gpurun_out/r2h_tl_b8.log; head -1 gpurun_out/r2h_tl_b32.log; tail -1 gpurun_out/r2h_tl_b32.log; head -1 gpurun_out/r2h_tl_b8_nocap.log; tail -1 gpurun_out/r2h_tl_b8_nocap.log
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2h_bench.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['roofline']['frac'], d['roofline']['prefill'], d['e2e'])
print(json.dumps(d['box'])[:900]); print(d['configs'])
PY
