"""Routing mirror: crowdllama.Resource (/root/reference/pkg/crowdllama/types.go:30-40) and
Manager.FindBestWorker (/root/reference/pkg/peermanager/manager.go:338-387), plus truthful worker
metadata fed from cl_engine_stats (replaces the constants at peer.go:319-358)."""
from __future__ import annotations

import json
import random
from dataclasses import asdict, dataclass, field
from datetime import datetime, timezone


@dataclass
class Resource:
    peer_id: str = ""
    supported_models: list = field(default_factory=list)
    tokens_throughput: float = 0.0
    vram_gb: int = 0
    load: float = 0.0
    gpu_model: str = ""
    last_updated: str = ""
    version: str = "unknown"
    worker_mode: bool = False

    def to_json(self) -> bytes:                                      # types.go:58-64
        return json.dumps(asdict(self)).encode()

    @classmethod
    def from_json(cls, data: bytes) -> "Resource":                   # types.go:67-74
        try:
            d = json.loads(data)
        except Exception as ex:
            raise ValueError(f"failed to unmarshal CrowdLlamaResource: {ex}") from ex
        known = {k: d[k] for k in cls.__dataclass_fields__ if k in d}
        return cls(**known)

    def get_dht_key(self) -> str:                                    # types.go:77-79
        return "/ipns/" + self.peer_id


def advertised_throughput(tokens_per_sec: float) -> float:
    """Quantise the engine's capacity figure to half-octave buckets.  FindBestWorker compares scores with a strict '>' and
    breaks exact ties at random (Go map order, manager.go:369-377); the reference's workers all advertise the constant
    150, so identical machines tie and share the load.  cl_stats.tokens_per_sec is a load-independent estimate (device
    memory bandwidth / model bytes x max_batch) for the same reason: a MEASURED rate drops as the batch fills (a step
    at B = 32 takes 1.6x a step at B = 1), so busy workers would advertise less than idle ones and one idle worker
    would win every request until the next refresh (8 peers: 96 of 192 requests on one worker).  The bucket keeps
    small differences between boards of one model (memory clock bins) from breaking the tie."""
    if tokens_per_sec <= 0:
        return 0.0
    import math
    return float(round(2.0 ** (round(math.log2(tokens_per_sec) * 2.0) / 2.0), 1))


LOAD_FLAG_AT = 2.0          # (active + queued) / max_batch at which a worker advertises Load = 1


def advertised_load(load: float) -> float:
    """Two levels only: 0 in normal operation, 1 once a whole extra batch is waiting (load >= 2: a new request would sit
    through a full generation before it gets a slot).  The gateway sees metadata that is 2-30 s old (DiscoveryInterval
    10 s, MetadataUpdateInterval 30 s, manager.go:99-101) while a chat lasts about a second, so the advertised Load
    describes the past: a fine-grained load makes the momentarily least loaded worker win EVERY request until the next
    refresh (measured: 36 / 100 / 42 / 78 requests over four identical workers), and even a flag at load >= 1 shuns
    every worker that happened to be full at refresh time (8 peers, 256 clients: 60 ... 138 requests per worker, 0.63
    of 8x one worker; tools/route_sim.py reproduces it: 0.77 of balanced at a 2 s refresh, 0.56 at 10 s, against 0.92
    with the flag at 2).  Below the flag all workers tie, FindBestWorker's random tie-break spreads the requests as it
    does with the reference's constants, and a closed loop balances itself: a worker with fewer running requests
    completes fewer per second than it receives."""
    return 1.0 if load >= LOAD_FLAG_AT else 0.0


def resource_from_engine(peer_id: str, engine, version: str = "b200") -> Resource:
    st = engine.stats()
    return Resource(peer_id=peer_id, supported_models=[engine.model_name], tokens_throughput=advertised_throughput(float(st["tokens_per_sec"])),
                    vram_gb=int(st["vram_gb"]), load=advertised_load(float(st["load"])), gpu_model=st["gpu_model"],
                    last_updated=datetime.now(timezone.utc).isoformat(), version=version, worker_mode=True)


def find_best_worker(workers, required_model: str, rng: random.Random | None = None):
    """manager.go:338-387.  Exact-string model match, score = tokens_throughput / (1 + load), strict
    '>' against a best score that starts at 0 — so workers with score 0 are never selected, and ties go
    to whichever worker is visited first.  Go iterates a map (random order): modelled by a shuffle."""
    workers = [w for w in workers if w.worker_mode]
    if not workers:
        return None
    suitable = [w for w in workers if required_model in w.supported_models]
    if not suitable:
        return None
    order = list(suitable)
    (rng or random).shuffle(order)
    selected, best = None, 0.0
    for w in order:
        score = w.tokens_throughput / (1 + w.load)
        if score > best:
            best, selected = score, w
    return selected
