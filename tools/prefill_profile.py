"""In-pipeline per-kernel-class device times of a prefill (CUDA events after every launch): CL_PREFILL_PROFILE=1."""
import os
import sys
from pathlib import Path

import numpy as np

os.environ["CL_PREFILL_PROFILE"] = "1"
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from crowdllama_b200 import engine as eng  # noqa: E402

T = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
with eng.Engine(preset="llama3-8b", seed=1234, max_batch=1) as e:
    ids = np.array([(i * 7919 + 13) % e.cfg["vocab_size"] for i in range(T)], np.int32)
    for rep in range(3):
        s = e.seq_create()
        e.prefill(s, ids)
        e.seq_free(s)
