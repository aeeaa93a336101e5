#!/bin/bash
mkdir -p gpurun_out
( timeout 1200 python -m pytest tests/test_gpu_prefill.py tests/test_gpu_longctx.py tests/test_gpu_engine.py tests/test_gpu_fullsize.py tests/test_gpu_checkpoint.py -q -x -s -p no:cacheprovider -k "not greedy_256 and not b8_long and not b1_long and not mistral and not layers_batched" ) > gpurun_out/r2e_tests.log 2>&1
echo "rc=$?" >> gpurun_out/r2e_tests.log
CL_PREFILL_FUSED=0 timeout 300 python tools/prefill_profile.py 4096 > gpurun_out/r2e_prof_unfused.log 2>&1
timeout 300 python tools/prefill_profile.py 4096 > gpurun_out/r2e_prof_fused.log 2>&1
CL_PREFILL_SMALL_MAX=0 timeout 300 python tools/prefill_profile.py 128 > gpurun_out/r2e_prof_128_tiles.log 2>&1
timeout 300 python - > gpurun_out/r2e_small_prefill.log 2>&1 <<'PY'
import sys, time, numpy as np
sys.path.insert(0, '.')
from crowdllama_b200 import engine as eng
with eng.Engine(preset="llama3-8b", seed=1234, max_batch=1) as e:
    for T in (32, 64, 128, 256, 257, 512, 1024):
        ids = np.array([(i * 7919 + 13) % e.cfg["vocab_size"] for i in range(T)], np.int32)
        ts = []
        for rep in range(4):
            s = e.seq_create(); t0 = time.time(); e.prefill(s, ids); ts.append((time.time() - t0) * 1e3); e.seq_free(s)
        print(f"prefill T={T}: {min(ts[1:]):.3f} ms (host wall incl. logits read-back)", flush=True)
PY
timeout 300 python tools/batch_step_profile.py 8 1024 > gpurun_out/r2e_step_b8.log 2>&1
timeout 300 python tools/batch_step_profile.py 32 1024 > gpurun_out/r2e_step_b32.log 2>&1
tail -6 gpurun_out/r2e_tests.log
tail -11 gpurun_out/r2e_prof_unfused.log; tail -9 gpurun_out/r2e_prof_fused.log; tail -9 gpurun_out/r2e_prof_128_tiles.log; cat gpurun_out/r2e_small_prefill.log
grep graph gpurun_out/r2e_step_b8.log; tail -13 gpurun_out/r2e_step_b8.log; grep graph gpurun_out/r2e_step_b32.log; tail -13 gpurun_out/r2e_step_b32.log
