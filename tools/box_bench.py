"""Request-level benchmark (BASELINE.json configs[3]): N local worker peers (one process per GPU) behind the
gateway stand-in, C concurrent /api/chat requests, each a ~128-token prompt and G greedy tokens.
Prints one JSON line: req/s, aggregate tok/s, per-worker request counts (the reference's routing is
random among ties, so the imbalance is part of the result).

    python tools/box_bench.py --workers 1 --concurrency 8 --gen 128
"""
import argparse
import json
import multiprocessing as mp
import sys
import threading
import time
import urllib.request
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))


def _worker(device, port, preset, model_name, gen, max_batch, ev):
    from crowdllama_b200 import worker
    worker.serve(device, port, preset, model_name, max_batch=max_batch, greedy_tokens=gen, ready_event=ev)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workers", type=int, default=1)
    ap.add_argument("--concurrency", type=int, default=8)
    ap.add_argument("--requests", type=int, default=0, help="total requests (default = concurrency)")
    ap.add_argument("--gen", type=int, default=128)
    ap.add_argument("--preset", default="llama3-8b")
    ap.add_argument("--model-name", default="llama3:8b")
    ap.add_argument("--max-batch", type=int, default=8)
    ap.add_argument("--base-port", type=int, default=9101)
    a = ap.parse_args()
    mp.set_start_method("spawn")
    procs, addrs = [], []
    for i in range(a.workers):
        ev = mp.Event()
        p = mp.Process(target=_worker, args=(i, a.base_port + i, a.preset, a.model_name, a.gen, a.max_batch, ev), daemon=True)
        p.start()
        procs.append((p, ev))
        addrs.append(("127.0.0.1", a.base_port + i))
    for p, ev in procs:
        if not ev.wait(300):
            raise SystemExit("worker did not come up")
    from crowdllama_b200 import gateway
    gw = gateway.make_server(addrs, port=a.base_port - 100)
    threading.Thread(target=gw.serve_forever, daemon=True).start()
    url = f"http://127.0.0.1:{a.base_port - 100}/api/chat"
    prompt = ("Explain, step by step, why the sky appears blue during the day and red at sunset, and what changes on Mars. " * 2)[:118]
    n_req = a.requests or a.concurrency
    lat, errs = [], []

    def one(i):
        body = json.dumps({"model": a.model_name, "messages": [{"role": "user", "content": f"{i:03d} {prompt}"}], "stream": False}).encode()
        t0 = time.time()
        try:
            with urllib.request.urlopen(urllib.request.Request(url, body, {"Content-Type": "application/json"}), timeout=900) as r:
                out = json.loads(r.read())
            assert out["done"] and out["model"] == a.model_name and out["message"]["role"] == "assistant"
            lat.append(time.time() - t0)
        except Exception as ex:  # noqa: BLE001
            errs.append(str(ex))

    # warm EVERY worker directly (graph capture, prefill workspace), then one request through the gateway
    for addr in addrs:
        for _ in range(2):
            gateway.request_inference(addr, a.model_name, "warm-up " + prompt, False)
    one(-1)
    errs.clear()
    lat.clear()
    gw.counts.clear()
    sem = threading.Semaphore(a.concurrency)
    threads = []
    t0 = time.time()
    for i in range(n_req):
        sem.acquire()
        th = threading.Thread(target=lambda i=i: (one(i), sem.release()))
        th.start()
        threads.append(th)
    for th in threads:
        th.join()
    dt = time.time() - t0
    ok = len(lat)
    print(json.dumps({"bench": "box", "workers": a.workers, "concurrency": a.concurrency, "requests": n_req, "ok": ok, "errors": errs[:3],
                      "gen_tokens": a.gen, "wall_s": round(dt, 3), "req_per_s": round(ok / dt, 3),
                      "tok_per_s": round(ok * a.gen / dt, 1), "mean_latency_s": round(sum(lat) / max(ok, 1), 3),
                      "per_worker_requests": dict(gw.counts), "preset": a.preset}), flush=True)
    gw.shutdown()
    for p, _ in procs:
        p.terminate()


if __name__ == "__main__":
    main()
