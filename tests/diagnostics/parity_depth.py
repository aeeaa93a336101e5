"""Diagnostic: engine vs oracle error as a function of depth (Llama-3-8B layer shapes, L layers)."""
import os
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
from crowdllama_b200 import engine as eng  # noqa: E402
from oracle import oracle as oc  # noqa: E402

for mega in ("1", "0"):
    os.environ["CL_MEGA"] = mega
    for L in (1, 2, 4, 8, 16):
        cfg = dict(oc.PRESETS["llama3-8b"])
        cfg["n_layers"] = L
        cfg["max_seq_len"] = 128
        m = oc.Model(cfg, seed=1234)
        prompt = np.array([(i * 7919 + 13) % cfg["vocab_size"] for i in range(12)], np.int32)
        so = m.new_seq()
        lo = so.forward(prompt)
        with eng.Engine(model=cfg, seed=1234, max_batch=1) as e:
            s = e.seq_create()
            lg = None
            for t in prompt:                    # token-wise: decode kernels only
                lg, _ = e.decode_step(s, int(t))
            hid = e.debug_hidden()
            ho = m.hidden(L)
            rel = float(np.linalg.norm(hid - ho) / np.linalg.norm(ho))
            print(f"mega={mega} L={L:2d}: hidden rel err {rel:.3e}  |h|rms {np.sqrt((ho**2).mean()):.3f}  logits max err {np.abs(lg - lo).max():.4f} "
                  f"(logit std {lo.std():.3f})", flush=True)
