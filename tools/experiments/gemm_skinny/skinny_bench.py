"""Stand-alone timing + CTA-0 timeline of the batched-decode projection kernel (gemm_skinny.cu) on Llama-3-8B shapes."""
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from crowdllama_b200 import engine as eng  # noqa: E402

T = int(sys.argv[1]) if len(sys.argv) > 1 else 8
kbs = [int(x) for x in sys.argv[2].split(",")] if len(sys.argv) > 2 else [4, 8, 16]
shapes = {"qkv": (6144, 4096), "o": (4096, 4096), "gate_up": (28672, 4096), "down": (4096, 14336)}
names = ["start", "claim", "W issued", "pdl_wait", "mma 1st full", "mma unit commit", "epi 1st full", "epi loop end", "epi resolved", "end", "prod end"]
rng = np.random.default_rng(0)
for name, (n, k) in shapes.items():
    w = (rng.integers(0, 2**16, size=(n, k), dtype=np.uint16) & 0x3FFF) | 0x3C00   # bf16 in [1, 2) / small exponents
    x = (rng.integers(0, 2**16, size=(T, k), dtype=np.uint16) & 0x3FFF) | 0x3C00
    for kb in kbs:
        y, ms, dbg = eng.op_gemm_skinny(x, w, target_kb=kb, iters=20, want_dbg=True)
        gb = n * k * 2 / 1e9
        t0 = dbg[0]
        tl = " ".join(f"{names[i]}={(dbg[i]-t0)/1e3:.1f}" for i in range(1, 11) if dbg[i])
        print(f"{name:8s} T={T} kb={kb:2d} {ms*1e3:7.1f} us  {gb/ (ms*1e-3) :7.0f} GB/s | {tl}")
        ent = dbg[16:16 + 2 * 148:2].astype(np.float64); setup = dbg[17:17 + 2 * 148:2].astype(np.float64); ex = dbg[336:336 + 148].astype(np.float64)
        ok = ent > 0
        if ok.any():
            e0 = ent[ok].min()
            print(f"          CTAs {ok.sum()}: entry spread {(ent[ok].max()-e0)/1e3:.1f} us, setup {np.median(setup[ok]-ent[ok])/1e3:.1f} us (max {(setup[ok]-ent[ok]).max()/1e3:.1f}), "
                  f"exit min/med/max {(ex[ok].min()-e0)/1e3:.1f}/{(np.median(ex[ok])-e0)/1e3:.1f}/{(ex[ok].max()-e0)/1e3:.1f} us")
