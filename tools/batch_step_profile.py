"""In-pipeline breakdown of one batched decode step (default path): B sequences at ctx C, eager launches with an event
after each (CL_STEP_PROFILE=1), plus the CUDA-graph step time for reference.
    python tools/batch_step_profile.py B [ctx]"""
import os
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
ctx = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
from crowdllama_b200 import engine as eng  # noqa: E402

with eng.Engine(preset="llama3-8b", seed=1234, max_batch=B) as e:          # graph path: the number that counts
    seqs = [e.seq_create() for _ in range(B)]
    for s in seqs:
        e.seq_fake_fill(s, ctx)
    e.decode_greedy_batch(seqs, [17 + b for b in range(B)], 4)
    _, ms = e.decode_greedy_batch(seqs, [17 + b for b in range(B)], 16)
    print(f"graph: B={B} ctx={ctx}: {ms / 16:.3f} ms/step", flush=True)
os.environ["CL_GRAPH"] = "0"
os.environ["CL_STEP_PROFILE"] = "1"
with eng.Engine(preset="llama3-8b", seed=1234, max_batch=B) as e:
    seqs = [e.seq_create() for _ in range(B)]
    for s in seqs:
        e.seq_fake_fill(s, ctx)
    e.decode_greedy_batch(seqs, [17 + b for b in range(B)], 3)
