"""ncu workload: Llama-3-8B, B sequences (ctx 1024), a few eager batched decode steps."""
import os
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
os.environ.setdefault("CL_GRAPH", "0")
from crowdllama_b200 import engine as eng  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
with eng.Engine(preset="llama3-8b", seed=1234, max_batch=B) as e:
    V = e.cfg["vocab_size"]
    seqs, firsts = [], []
    for b in range(B):
        s = e.seq_create()
        lg = e.prefill(s, np.array([(i * 7919 + 13 + b) % V for i in range(1024)], np.int32))
        seqs.append(s); firsts.append(int(lg.argmax()))
    ids, ms = e.decode_greedy_batch(seqs, firsts, steps)
    print("ms/step", ms / steps)
