"""ncu workload: Llama-3-8B, B sequences with 1024 cached tokens each (synthetic cache: no prefill kernels in the
capture), a few eager batched decode steps."""
import os
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
os.environ.setdefault("CL_GRAPH", "0")
from crowdllama_b200 import engine as eng  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
ctx = int(sys.argv[3]) if len(sys.argv) > 3 else 1024
with eng.Engine(preset="llama3-8b", seed=1234, max_batch=B) as e:
    seqs = [e.seq_create() for _ in range(B)]
    for s in seqs:
        e.seq_fake_fill(s, ctx)
    ids, ms = e.decode_greedy_batch(seqs, [17 + b for b in range(B)], steps)
    print("ms/step", ms / steps)
