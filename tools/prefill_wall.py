"""Host wall time of one prefill call (incl. the logits read-back) against the prompt length."""
import sys
import time
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from crowdllama_b200 import engine as eng  # noqa: E402

Ts = [int(x) for x in sys.argv[1].split(",")] if len(sys.argv) > 1 else [32, 64, 128, 146, 192, 256, 257, 512, 1024]
with eng.Engine(preset="llama3-8b", seed=1234, max_batch=1) as e:
    for T in Ts:
        ids = np.array([(i * 7919 + 13) % e.cfg["vocab_size"] for i in range(T)], np.int32)
        ts = []
        for rep in range(5):
            s = e.seq_create(); t0 = time.time(); e.prefill(s, ids); ts.append((time.time() - t0) * 1e3); e.seq_free(s)
        print(f"prefill T={T}: {min(ts[1:]):.3f} ms (host wall incl. logits read-back)", flush=True)
