"""BASELINE.json configs[4] ("Config 5") on N worker peers: Mistral-7B shapes, ~8K-token prompts, 256 new
tokens, a closed loop of 64 clients for >= 60 s through the gateway stand-in (/api/chat -> FindBestWorker -> worker peers
-> cl_handle_message -> scheduler), every worker's paged-KV pool capped so that it has to preempt (evict + recompute).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 \\
        tools/box_config5.py [--seconds 60] [--clients 64] [--pool-tokens 98304]
    python tools/box_config5.py --seconds 30            # one worker

One JSON object on rank 0's stdout: req/s, generated and prompt tok/s, per-worker request counts, preemptions per worker,
errors.  (The zero-corruption check — teacher-forced replay of evicted requests — needs token ids and lives in the
single-GPU tool, tools/evict_bench.py.)"""
import argparse
import json
import os
import socket
import subprocess
import sys
import threading
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
PRESET, MODEL = "mistral-7b", "mistral:7b"
GEN = 256


def client(a):
    import urllib.request
    from crowdllama_b200 import gateway
    from crowdllama_b200.worker import STATS_PROTOCOL, STOP_PROTOCOL
    import bench
    addrs = [("127.0.0.1", a.base_port + i) for i in range(a.workers)]
    out = {"bench": "box_config5", "workers": a.workers, "clients": a.clients, "gen_tokens": GEN, "ctx": [a.ctx_min, a.ctx_max],
           "pool_tokens_per_worker": a.pool_tokens, "max_batch_per_worker": a.max_batch}
    try:
        for ad in addrs:
            if not bench._wait_port(ad, 900):
                raise RuntimeError(f"worker {ad} did not come up")
        gw = gateway.make_server(addrs, port=a.base_port - 1)
        threading.Thread(target=gw.serve_forever, daemon=True).start()
        url = f"http://127.0.0.1:{a.base_port - 1}/api/chat"
        words = "the quick brown fox jumps over the lazy dog while continuous batching keeps every streaming multiprocessor busy ".split()
        for ad in addrs:                                                # warm: graphs, prefill workspaces
            gateway.request_inference(ad, MODEL, "warm " * 400, False)
        done, lock, stop = [], threading.Lock(), threading.Event()

        def one_client(cid):
            r = np.random.default_rng(1000 + cid)
            while not stop.is_set():
                n = int(r.integers(a.ctx_min, a.ctx_max + 1))
                text = " ".join(words[int(i)] for i in r.integers(0, len(words), size=n // 5))[: n - 40]
                body = json.dumps({"model": MODEL, "messages": [{"role": "user", "content": f"{cid:03d} {text}"}], "stream": False}).encode()
                t0 = time.time()
                try:
                    with urllib.request.urlopen(urllib.request.Request(url, body, {"Content-Type": "application/json"}), timeout=1800) as resp:
                        o = json.loads(resp.read())
                    ok = bool(o.get("done")) and not o["message"]["content"].startswith("Error:")
                    with lock:
                        done.append((time.time(), time.time() - t0, len(text), ok, "" if ok else o["message"]["content"][:120]))
                except Exception as ex:  # noqa: BLE001
                    with lock:
                        done.append((time.time(), time.time() - t0, len(text), False, repr(ex)[:120]))
        with gw.lock:
            gw.counts.clear()
        t_start = time.time()
        th = [threading.Thread(target=one_client, args=(i,), daemon=True) for i in range(a.clients)]
        [t.start() for t in th]
        time.sleep(a.seconds)
        stop.set()
        t_win = time.time()
        [t.join(timeout=600) for t in th]
        with lock:
            inw = [d for d in done if d[0] <= t_win and d[3]]
            errs = [d[4] for d in done if not d[3]]
        wall = t_win - t_start
        stats = []
        for ad in addrs:
            with socket.create_connection(ad, timeout=10) as s:
                s.sendall((STATS_PROTOCOL + "\n").encode())
                data = b""
                while chunk := s.recv(65536):
                    data += chunk
            stats.append(json.loads(data))
        out.update(seconds=round(wall, 1), requests_completed=len(inw), requests_per_s=round(len(inw) / wall, 3),
                   gen_tokens_per_s=round(len(inw) * GEN / wall, 1), prompt_tokens_per_s=round(sum(d[2] for d in inw) / wall, 1),
                   p50_latency_s=round(float(np.median([d[1] for d in inw])), 2) if inw else None, errors=len(errs), error_samples=errs[:3],
                   per_worker_requests=dict(sorted(gw.counts.items())), preemptions_per_worker=[int(s["preemptions"]) for s in stats],
                   preemptions=int(sum(s["preemptions"] for s in stats)), kv_pages_total_per_worker=stats[0]["kv_pages_total"])
        gw.shutdown()
    except Exception as ex:  # noqa: BLE001
        out["error"] = repr(ex)
    for ad in addrs:
        try:
            with socket.create_connection(ad, timeout=5) as s:
                s.sendall((STOP_PROTOCOL + "\n").encode())
        except OSError:
            pass
    print(json.dumps(out), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--seconds", type=float, default=60.0)
    ap.add_argument("--clients", type=int, default=64)
    ap.add_argument("--pool-tokens", type=int, default=96 * 1024)
    ap.add_argument("--max-batch", type=int, default=16)
    ap.add_argument("--ctx-min", type=int, default=7000)
    ap.add_argument("--ctx-max", type=int, default=7900)
    ap.add_argument("--client", action="store_true")
    ap.add_argument("--workers", type=int, default=1)
    ap.add_argument("--base-port", type=int, default=23001)
    a = ap.parse_args()
    if a.client:
        return client(a)
    from crowdllama_b200 import engine as eng
    from crowdllama_b200.distutil import Group
    from crowdllama_b200.worker import WorkerServer
    grp = Group()
    c = eng.model_preset(PRESET)
    kv_tok = 2 * c["n_layers"] * c["n_kv_heads"] * c["head_dim"] * 2
    e = eng.Engine(preset=PRESET, model_name=MODEL, device=grp.local_rank, seed=1234, max_batch=a.max_batch, kv_pool_bytes=a.pool_tokens * kv_tok,
                   start_scheduler=True)
    base = 23000 + (int(os.environ.get("MASTER_PORT", "29500")) % 500) * 16 + 1
    srv = WorkerServer(("127.0.0.1", base + grp.rank), e, peer_id=f"b200-worker-{grp.rank}", sampling=eng.greedy(GEN, ignore_eos=True))
    threading.Thread(target=srv.serve_forever, kwargs={"poll_interval": 0.1}, daemon=True).start()
    grp.barrier()
    proc = None
    if grp.rank == 0:
        proc = subprocess.Popen([sys.executable, __file__, "--client", "--workers", str(max(grp.world, 1)), "--base-port", str(base),
                                 "--seconds", str(a.seconds), "--clients", str(a.clients), "--pool-tokens", str(a.pool_tokens),
                                 "--max-batch", str(a.max_batch), "--ctx-min", str(a.ctx_min), "--ctx-max", str(a.ctx_max)],
                                stdout=subprocess.PIPE, text=True, env={**os.environ, "CUDA_VISIBLE_DEVICES": "", "RANK": "0", "WORLD_SIZE": "1"})
    srv.stop_event.wait(a.seconds + 1500)
    if proc is not None:
        so, _ = proc.communicate(timeout=120)
        print(so.strip().splitlines()[-1], flush=True)
    srv.shutdown()
    e.close()
    grp.barrier()
    grp.close()


if __name__ == "__main__":
    main()
