"""world_size-2 gloo test (CPU) of the N-replica benchmark plumbing: barrier, max-over-ranks timing and the
whole-job aggregate (units of all ranks / slowest rank's time) that bench.py reports for --gpus N."""
import os
import socket
import subprocess
import sys
import textwrap
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]

CHILD = textwrap.dedent("""
    import json, os, sys, time
    sys.path.insert(0, %r)
    from crowdllama_b200.distutil import Group, aggregate_throughput
    g = Group(backend="gloo")
    assert g.world == 2 and g.device == "cpu"
    g.barrier()
    ms = 3.0 + g.rank          # rank 1 is the slow replica
    out = {"rank": g.rank, "max": g.max(ms), "sum": g.sum(256.0), "agg": aggregate_throughput(g, 256.0, ms * 1e-3)}
    g.barrier()
    g.close()
    print("RESULT " + json.dumps(out), flush=True)
""") % str(ROOT)


def test_two_rank_gloo_aggregate():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    procs = []
    for r in range(2):
        env = dict(os.environ, RANK=str(r), LOCAL_RANK=str(r), WORLD_SIZE="2", MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
        procs.append(subprocess.Popen([sys.executable, "-c", CHILD], env=env, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True))
    outs = []
    for p in procs:
        o, e = p.communicate(timeout=120)
        assert p.returncode == 0, e[-2000:]
        outs.append([l for l in o.splitlines() if l.startswith("RESULT ")][0][7:])
    import json
    res = [json.loads(o) for o in outs]
    for r in res:
        assert r["max"] == 4.0 and r["sum"] == 512.0
        assert abs(r["agg"] - 512.0 / 4.0e-3) < 1e-6        # 2 x 256 tokens / slowest replica's 4 ms


def test_single_process_group_is_a_noop():
    from crowdllama_b200.distutil import Group, aggregate_throughput
    env = {k: os.environ.pop(k) for k in ("RANK", "WORLD_SIZE", "LOCAL_RANK") if k in os.environ}
    try:
        g = Group()
        assert g.world == 1 and g.max(2.5) == 2.5 and g.sum(7) == 7.0
        g.barrier()
        assert aggregate_throughput(g, 100, 0.5) == 200.0
    finally:
        os.environ.update(env)
