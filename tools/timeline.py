"""Print the in-step timeline (CL_TIMELINE=1): per kernel node, when CTA 0 started, when its dependency was
satisfied, when it began consuming and when it finished — relative to the first node of the layer."""
import os
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
os.environ["CL_TIMELINE"] = "1"
from crowdllama_b200 import engine as eng  # noqa: E402

ctx = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
with eng.Engine(preset="llama3-8b", seed=1234, max_batch=1) as e:
    V = e.cfg["vocab_size"]
    ids = np.array([(i * 7919 + 13) % V for i in range(ctx)], np.int32)
    s = e.seq_create()
    lg = e.prefill(s, ids)
    out, ms = e.decode_greedy(s, int(lg.argmax()), 32)
    tl = e.debug_timeline().astype(np.float64)
    print(f"ms/step {ms/32:.4f}")
    names = ["qkv", "attn", "o", "gateup", "down"]
    L = e.cfg["n_layers"]
    for l in (1, 15, 30):
        base = tl[l * 5][0]
        print(f"layer {l}: (us since the layer's qkv CTA0 start)  start / dep-ok / consuming / done")
        for k in range(5):
            r = (tl[l * 5 + k] - base) / 1e3
            print(f"   {names[k]:7s} {r[0]:8.2f} {r[1]:8.2f} {r[2]:8.2f} {r[3]:8.2f}")
        nxt = (tl[(l + 1) * 5][0] - base) / 1e3
        print(f"   next layer qkv start {nxt:8.2f}   layer time (done-to-done) {(tl[l*5+4][3]-tl[(l-1)*5+4][3])/1e3:8.2f}")
    per_layer = np.diff(tl[4::5][:L, 3]) / 1e3
    print("layer done-to-done us: mean %.2f min %.2f max %.2f" % (per_layer.mean(), per_layer.min(), per_layer.max()))
    # average over layers 1..L-1 of each stamp relative to previous layer's down done
    rel = np.zeros((5, 4))
    for l in range(1, L):
        base = tl[(l - 1) * 5 + 4][3]
        for k in range(5):
            rel[k] += (tl[l * 5 + k] - base) / 1e3
    rel /= (L - 1)
    print("mean over layers, us since previous layer's down-proj done:  start / dep-ok / consuming / done")
    for k in range(5):
        print(f"   {names[k]:7s} " + " ".join(f"{v:8.2f}" for v in rel[k]))
