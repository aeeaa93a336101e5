#!/bin/bash
# round-2 GPU pass E: validate fused epilogues, short-prompt prefill, budgeted admission, persistent batched kernel
mkdir -p gpurun_out
rm -f gpurun_out/parity_longctx.jsonl
( timeout 1500 python -m pytest tests/test_gpu_prefill.py tests/test_gpu_engine.py tests/test_gpu_fullsize.py tests/test_gpu_checkpoint.py tests/test_gpu_kernels.py -q -x -s -p no:cacheprovider ) > gpurun_out/r2e_tests1.log 2>&1
echo "rc=$?" >> gpurun_out/r2e_tests1.log
( timeout 1500 python -m pytest tests/test_gpu_longctx.py -q -s -p no:cacheprovider -k "prefill_4096 or layers_batched" ) > gpurun_out/r2e_tests2.log 2>&1
echo "rc=$?" >> gpurun_out/r2e_tests2.log
CL_PREFILL_FUSED=0 timeout 300 python tools/prefill_profile.py 4096 > gpurun_out/r2e_prof_unfused.log 2>&1
CL_PREFILL_FUSED=1 timeout 300 python tools/prefill_profile.py 4096 > gpurun_out/r2e_prof_fused.log 2>&1
CL_PREFILL_SMALL_MAX=0 timeout 300 python tools/prefill_profile.py 128 > gpurun_out/r2e_prof_128_tiles.log 2>&1
CL_PREFILL_SMALL_MAX=256 CL_PREFILL_FUSED=1 timeout 300 python - > gpurun_out/r2e_small_prefill.log 2>&1 <<'PY'
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
CL_BATCH_MEGA=1 timeout 300 python - > gpurun_out/r2e_batch_mega.log 2>&1 <<'PY'
import sys, numpy as np
sys.path.insert(0, '.')
from crowdllama_b200 import engine as eng
for B in (2, 4, 8, 16, 32):
    with eng.Engine(preset="llama3-8b", seed=1234, max_batch=B) as e:
        seqs = [e.seq_create() for _ in range(B)]
        for s in seqs:
            e.seq_fake_fill(s, 1024)
        e.decode_greedy_batch(seqs, [17 + b for b in range(B)], 4)
        _, ms = e.decode_greedy_batch(seqs, [17 + b for b in range(B)], 16)
        print(f"persistent batched kernel: B={B} ctx=1024: {ms / 16:.3f} ms/step", flush=True)
PY
tail -5 gpurun_out/r2e_tests1.log; tail -12 gpurun_out/r2e_tests2.log
tail -11 gpurun_out/r2e_prof_unfused.log; tail -9 gpurun_out/r2e_prof_fused.log; tail -9 gpurun_out/r2e_prof_128_tiles.log; cat gpurun_out/r2e_small_prefill.log
grep graph gpurun_out/r2e_step_b8.log; tail -13 gpurun_out/r2e_step_b8.log; grep graph gpurun_out/r2e_step_b32.log; tail -13 gpurun_out/r2e_step_b32.log
tail -8 gpurun_out/r2e_batch_mega.log
