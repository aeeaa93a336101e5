"""Worker peer stand-in: one process per GPU serving the reference's two stream protocols over TCP.

The reference's worker is a libp2p host (pkg/peer/peer.go); Go and libp2p are absent here, so this harness
keeps the *application* protocols and replaces the transport by plain TCP with a multistream-style first
line naming the protocol:
    "/crowdllama/inference/1.0.0\\n"  then one length-prefixed BaseMessage each way   (peer.go:190-256)
    "/crowdllama/metadata/1.0.0\\n"   then the worker writes its Resource JSON and closes (peer.go:284-316)
The handler behind the inference protocol is `handler.worker_api_handler(engine)` — the same closure shape
the Go cgo shim installs into Peer.APIHandler.

    python -m crowdllama_b200.worker --device 0 --port 9101 --preset llama3-8b --model-name llama3:8b
"""
from __future__ import annotations

import argparse
import socketserver
import threading

from . import engine as eng
from . import handler as H
from .router import resource_from_engine

INFERENCE_PROTOCOL = "/crowdllama/inference/1.0.0"     # pkg/crowdllama/types.go:20
METADATA_PROTOCOL = "/crowdllama/metadata/1.0.0"       # pkg/crowdllama/types.go:16
STOP_PROTOCOL = "/crowdllama-b200/bench-stop/1.0.0"    # harness only: the load generator tells the worker peers it is done
STATS_PROTOCOL = "/crowdllama-b200/bench-stats/1.0.0"  # harness only: raw cl_engine_stats as JSON (preemptions, KV pages ...)


class _Stream:
    def __init__(self, sock):
        self.r = sock.makefile("rb")
        self.w = sock.makefile("wb")

    def read(self, n):
        return self.r.read(n)

    def write(self, b):
        n = self.w.write(b)
        self.w.flush()
        return n


class WorkerServer(socketserver.ThreadingTCPServer):
    allow_reuse_address = True
    daemon_threads = True
    request_queue_size = 256

    def __init__(self, addr, engine: "eng.Engine", peer_id: str, sampling=None):
        self.engine, self.peer_id = engine, peer_id
        self.api_handler = H.worker_api_handler(engine, sampling)
        self.served = 0
        self.stop_event = threading.Event()
        self._lock = threading.Lock()
        super().__init__(addr, _Conn)


class _Conn(socketserver.BaseRequestHandler):
    def handle(self):
        srv: WorkerServer = self.server
        s = _Stream(self.request)
        proto = s.r.readline().decode(errors="replace").strip()
        if proto == METADATA_PROTOCOL:
            s.write(resource_from_engine(srv.peer_id, srv.engine).to_json())
        elif proto == STOP_PROTOCOL:
            srv.stop_event.set()
        elif proto == STATS_PROTOCOL:
            import json
            s.write(json.dumps(srv.engine.stats()).encode())
        elif proto == INFERENCE_PROTOCOL:
            if H.handle_inference_stream(srv.api_handler, s, worker_mode=True):
                with srv._lock:
                    srv.served += 1


def serve(device: int, port: int, preset: str, model_name: str, max_batch: int = 8, greedy_tokens: int = 0, host="127.0.0.1",
          ready_event=None, seed: int = 1234, tokenizer: str | None = None, chat_family: str | None = None):
    e = eng.Engine(preset=preset, model_name=model_name, device=device, seed=seed, max_batch=max_batch, start_scheduler=True)
    if tokenizer:
        e.load_tokenizer(tokenizer, chat_family)      # HF tokenizer.json instead of the byte-level fallback
    sampling = eng.greedy(greedy_tokens, ignore_eos=True) if greedy_tokens > 0 else None   # None = Ollama defaults
    srv = WorkerServer((host, port), e, peer_id=f"b200-worker-{device}", sampling=sampling)
    if ready_event is not None:
        ready_event.set()
    try:
        srv.serve_forever(poll_interval=0.2)
    finally:
        srv.server_close()
        e.close()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--device", type=int, default=0)
    ap.add_argument("--port", type=int, default=9101)
    ap.add_argument("--preset", default="llama3-8b")
    ap.add_argument("--model-name", default="llama3:8b")
    ap.add_argument("--max-batch", type=int, default=8)
    ap.add_argument("--greedy-tokens", type=int, default=0, help="> 0: force greedy with this num_predict (benchmarks)")
    ap.add_argument("--tokenizer", default=None, help="path of an HF tokenizer.json (default: byte-level fallback)")
    ap.add_argument("--chat-family", default=None, help="llama3 | mistral | zephyr | chatml (default: auto-detect)")
    a = ap.parse_args()
    print(f"worker on cuda:{a.device} port {a.port} serving {a.model_name}", flush=True)
    serve(a.device, a.port, a.preset, a.model_name, a.max_batch, a.greedy_tokens, tokenizer=a.tokenizer, chat_family=a.chat_family)


if __name__ == "__main__":
    main()
