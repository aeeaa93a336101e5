"""Per-phase timeline of the persistent batched decode kernel (decode_mega_batch.cu): globaltimer stamps of CTA 0.
    CL_BATCH_MEGA=1 CL_TIMELINE=1 python tools/timeline_batch.py B [ctx]"""
import os
import sys
from pathlib import Path

import numpy as np

os.environ["CL_BATCH_MEGA"] = "1"
os.environ["CL_TIMELINE"] = "1"
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from crowdllama_b200 import engine as eng  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
ctx = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
with eng.Engine(preset="llama3-8b", seed=1234, max_batch=B) as e:
    seqs = [e.seq_create() for _ in range(B)]
    for s in seqs:
        e.seq_fake_fill(s, ctx)
    e.decode_greedy_batch(seqs, [17 + b for b in range(B)], 4)
    _, ms = e.decode_greedy_batch(seqs, [17 + b for b in range(B)], 8)
    L = e.cfg["n_layers"]
    raw = e.debug_timeline().reshape(-1)
    tl = raw[: L * 16].reshape(L, 16).astype(np.float64) / 1e3       # us
    dbg = raw[L * 16: L * 16 + 16].astype(np.float64) / 1965.0       # cycles -> us at 1965 MHz
    print(f"CTA 0 producer (whole step, us): total {dbg[0]:.0f}, blocked on: free slot {dbg[1]:.0f}, flight cap / pause {dbg[2]:.0f}, "
          f"X slot {dbg[3]:.0f}, phase publication {dbg[4]:.0f}, kv_ready {dbg[5]:.0f};  MMA warp blocked on: W tile {dbg[8]:.0f}, X tile {dbg[9]:.0f}, "
          f"accumulator {dbg[10]:.0f}")
    names = ["R0 resid+norm", "barrier A", "G0 q|k|v epilogue done", "barrier B", "AT attention", "barrier C", "G1 o", "barrier D", "R1 resid+norm",
             "barrier E", "G2 gate|up", "barrier F", "R2 silu", "barrier G", "G3 down"]
    order = [15, 0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14]
    print(f"B={B} ctx={ctx}: {ms / 8:.3f} ms/step; per-layer phase durations of CTA 0 (us), mean over layers 2..{L - 2}:")
    tot = 0.0
    for i in range(1, 16):
        d = tl[2:L - 1, order[i]] - tl[2:L - 1, order[i - 1]]
        tot += d.mean()
        print(f"  {names[i - 1]:28s} {d.mean():7.2f}  (min {d.min():6.2f} max {d.max():6.2f})")
    lay = tl[3:L - 1, 15] - tl[2:L - 2, 15]
    print(f"  layer period {lay.mean():.2f} us (sum of phases {tot:.2f})")
