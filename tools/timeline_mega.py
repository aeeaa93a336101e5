"""Phase timeline of the persistent decode kernel (CL_TIMELINE=1, CL_MEGA=1): per layer, us between stamps."""
import os
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
os.environ["CL_TIMELINE"] = "1"
os.environ["CL_MEGA"] = "1"
from crowdllama_b200 import engine as eng  # noqa: E402

ctx = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
names = ["top->x(norm)", "qkv gemv", "epi+B0", "attn tiles", "B1", "combine+B2", "x+o gemv", "epi+B3", "norm", "gateup gemv",
         "epi+B4", "act load", "down gemv", "epi+B5(next top)"]
with eng.Engine(preset="llama3-8b", seed=1234, max_batch=1) as e:
    V = e.cfg["vocab_size"]
    ids = np.array([(i * 7919 + 13) % V for i in range(ctx)], np.int32)
    s = e.seq_create()
    lg = e.prefill(s, ids)
    out, ms = e.decode_greedy(s, int(lg.argmax()), 32)
    L = e.cfg["n_layers"]
    full = e.debug_timeline().ravel()
    dbg = full[L * 16: L * 16 + 5]
    print(f"lookahead (this CTA, last step): prefetched {dbg[0]} tiles, budget-limited calls {dbg[1]}, jumps {dbg[2]}, calls {dbg[3]}, demand tiles {dbg[4]}")
    raw = full.astype(np.float64)[: L * 16].reshape(L, 16)[:, :14]
    print(f"ms/step {ms/32:.4f}  (CTA {os.environ.get('CL_TIMELINE_CTA', '0')})")
    d = np.zeros((L - 2, 14))
    for l in range(1, L - 1):
        t = raw[l]
        nxt = raw[l + 1][0]
        d[l - 1, :13] = np.diff(t) / 1e3
        d[l - 1, 13] = (nxt - t[13]) / 1e3
    mean = d.mean(0)
    for n, v in zip(names, mean):
        print(f"   {n:18s} {v:7.2f} us")
    print(f"   layer total        {mean.sum():7.2f} us")
