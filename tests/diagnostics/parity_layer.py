"""Diagnostic: one layer (Llama-3-8B shapes), per-op path: compare q, attention output, activation, hidden with the oracle."""
import os
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
os.environ["CL_MEGA"] = sys.argv[1] if len(sys.argv) > 1 else "0"
from crowdllama_b200 import engine as eng  # noqa: E402
from oracle import oracle as oc  # noqa: E402
import ctypes as C  # noqa: E402

cfg = dict(oc.PRESETS["llama3-8b"]); cfg["n_layers"] = 1; cfg["max_seq_len"] = 128
m = oc.Model(cfg, seed=1234)
prompt = [(i * 7919 + 13) % cfg["vocab_size"] for i in range(6)]
so = m.new_seq()


def rel(a, b):
    return float(np.linalg.norm(a - b) / (np.linalg.norm(b) + 1e-30)), float((a != b).mean())


with eng.Engine(model=cfg, seed=1234, max_batch=1) as e:
    s = e.seq_create()
    for t in prompt:
        so.forward([t])
        e.decode_step(s, int(t))
        qd = cfg["n_heads"] * cfg["head_dim"]
        bufs = {}
        for name, n, cnt in (("q", -1, qd), ("attn", -2, qd), ("act", -3, cfg["d_ff"]), ("h", cfg["d_model"], cfg["d_model"])):
            out = np.empty(cnt, np.float32)
            rc = eng.lib().cl_debug_hidden(e._h, out.ctypes.data_as(C.c_void_p), n)
            assert rc == 0, rc
            bufs[name] = out
        print(f"pos {len(so)-1}: q {rel(bufs['q'], m.debug_vec(0))}  attn {rel(bufs['attn'], m.debug_vec(1))}  "
              f"act {rel(bufs['act'], m.debug_vec(2))}  h {rel(bufs['h'], m.hidden(1))}", flush=True)
