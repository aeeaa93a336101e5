"""Batched decode step time against the batch size B (graph path), synthetic KV of `ctx` tokens per sequence."""
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from crowdllama_b200 import engine as eng  # noqa: E402

ctx = int(sys.argv[1]) if len(sys.argv) > 1 else 256
Bs = [int(x) for x in sys.argv[2].split(",")] if len(sys.argv) > 2 else [1, 2, 4, 8, 12, 16, 17, 24, 28, 29, 30, 31, 32]
with eng.Engine(preset="llama3-8b", seed=1234, max_batch=max(32, max(Bs)), max_seqs=max(32, max(Bs)) + 2) as e:
    for B in Bs:
        seqs = [e.seq_create() for _ in range(B)]
        for s in seqs:
            e.seq_fake_fill(s, ctx)
        e.decode_greedy_batch(seqs, [17 + b for b in range(B)], 4)
        ids, ms = e.decode_greedy_batch(seqs, [17 + b for b in range(B)], 32)
        print(f"B={B:3d} ctx={ctx}: {ms / 32:.3f} ms/step", flush=True)
        for s in seqs:
            e.seq_free(s)
