#!/usr/bin/env python
"""bench.py — headline benchmark of the B200 worker decode engine (contract: see README / DESIGN.md §6).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

Workload (BASELINE.json configs[2], the configuration the metric "decode tokens/sec/GPU
(Llama-3-8B, seq 4K)" is quoted on): Llama-3-8B shapes, bf16 synthetic seeded weights (no
checkpoints exist offline), a 4096-token prompt is prefilled, then a greedy 1-token decode loop.
One "step" = one decoded token = one pass of the whole token step over all weights (15 GB) and the
KV cache of the sequence.  N GPUs = N independent replicas (one process per GPU, no collective on
the data path — SURVEY.md §8e): weak scaling, value = N*K tokens / max-over-ranks device time.

`--impl reference` times the reference arm: the reference worker's CPU path.  The reference's own
implementation (Ollama v0.9.6) cannot be built or installed offline, so the arm runs the CPU
oracle port of the same token step (oracle/, OpenMP over all host cores) on the same config.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

PRESET = "llama3-8b"
CTX = 4096
SEED = 1234
METRIC = "decode tokens/sec (Llama-3-8B bf16, seq 4K; aggregate over local worker peers)"


def load_peaks():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        try:
            d = json.loads(p.read_text())
            return float(d["hbm_gbs"]), float(d.get("bf16_tflops_sustained", 1453.9)), "measured"
        except Exception:  # noqa: BLE001
            pass
    return 6650.0, 1400.0, "fallback"


def model_bytes(cfg):
    p_read = cfg["n_layers"] * ((cfg["n_heads"] + 2 * cfg["n_kv_heads"]) * cfg["head_dim"] * cfg["d_model"]
                                + cfg["d_model"] * cfg["n_heads"] * cfg["head_dim"] + 3 * cfg["d_ff"] * cfg["d_model"]
                                + 2 * cfg["d_model"]) + cfg["d_model"] + cfg["vocab_size"] * cfg["d_model"]
    kv_tok = 2 * cfg["n_layers"] * cfg["n_kv_heads"] * cfg["head_dim"] * 2
    return p_read, kv_tok


def prompt_ids(n, vocab):
    return np.array([(i * 7919 + 13) % vocab for i in range(n)], np.int32)


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, device):
        self.device, self.proc, self.lines = device, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "50",
                                          "-i", str(self.device)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.06)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:  # noqa: BLE001
            self.proc.kill()
        sm, mx, pw, reasons = [], [], [], set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1])); pw.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(pw) if pw else None, "samples": len(sm), "reasons": sorted(reasons)}


def cpu_baseline(steps_cap_s=25.0, max_tokens=8, warm=1):
    """The CPU oracle port on this box's host cores: Llama-3-8B shapes, synthetic weights, KV cache
    pre-filled to 4096 positions (timing only), a few greedy decode steps."""
    from oracle import oracle as oc
    cfg = dict(oc.PRESETS[PRESET])
    cfg["max_seq_len"] = CTX + 64
    threads = oc.effective_cpus()
    oc.set_threads(threads)
    t0 = time.time()
    m = oc.Model(cfg, seed=SEED)
    gen_s = time.time() - t0
    s = m.new_seq(CTX + 64)
    s.fake_fill(CTX)
    tok = 17
    for _ in range(warm):
        lg = s.forward([tok]); tok = int(lg.argmax())
    t0 = time.time()
    n = 0
    while n < max_tokens and (time.time() - t0) < steps_cap_s:
        lg = s.forward([tok]); tok = int(lg.argmax()); n += 1
    dt = time.time() - t0
    return {"value": round(n / dt, 4), "unit": "tokens/s", "cores": threads, "kind": "port",
            "sample": f"{n} greedy decode steps at ctx {CTX} (KV pre-filled with a synthetic pattern), Llama-3-8B shapes, "
                      f"seeded bf16 weights generated in {gen_s:.1f}s; oracle/llama_oracle.c with OpenMP x{threads}",
            "ms_per_step": round(dt / max(n, 1) * 1e3, 2)}, m, s


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    cb, m, s = cpu_baseline(steps_cap_s=0.0, max_tokens=0, warm=0)  # builds model + pre-filled cache
    tok, t_first = 17, time.time()
    lg = s.forward([tok]); tok = int(lg.argmax())
    t_tok = time.time() - t_first
    warm = max(0, min(args.warmup, 2) - 1)
    for _ in range(warm):
        lg = s.forward([tok]); tok = int(lg.argmax())
    steps = max(3, min(args.steps, int(150.0 / max(t_tok, 1e-3))))
    t0 = time.time()
    for _ in range(steps):
        lg = s.forward([tok]); tok = int(lg.argmax())
    dt = time.time() - t0
    val = steps / dt
    cb.update(value=round(val, 4), ms_per_step=round(dt / steps * 1e3, 2),
              sample=f"{steps} greedy decode steps at ctx {CTX} (bounded from --steps {args.steps}); " + cb["sample"].split(";", 1)[-1].strip())
    line = {"impl": "reference", "metric": METRIC, "value": round(val, 4), "unit": "tokens/s", "n_gpus": args.gpus, "steps": steps,
            "warmup": warm + 1, "ms_per_step": round(dt / steps * 1e3, 2), "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": {"workload": f"configs[2]: Llama-3-8B bf16 shapes, ctx {CTX}, 1-token greedy decode loop", "preset": PRESET,
                       "ctx": CTX, "weights": f"synthetic seed {SEED}", "note": "reference arm = CPU oracle port of the worker's "
                       "model step (Ollama v0.9.6 cannot be built offline); rank 0 only"},
            "cpu_baseline": cb, "e2e": {"value": round(val, 4), "unit": "tokens/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)
    return 0


def run_ours(args):
    from crowdllama_b200 import engine as eng
    hbm, tflops, peak_src = load_peaks()
    from crowdllama_b200.distutil import Group
    grp = Group()
    rank, local, world = grp.rank, grp.local_rank, grp.world
    barrier = lambda _d, _l: grp.barrier()                      # noqa: E731
    max_over_ranks = lambda _d, x, _l: grp.max(x)               # noqa: E731
    dist = None
    if world != args.gpus and world > 1:
        print(f"warning: WORLD_SIZE={world} != --gpus {args.gpus}", file=sys.stderr)
    n_gpus = max(world, 1)
    K, W = args.steps, max(args.warmup, 3)
    t0 = time.time()
    e = eng.Engine(preset=args.preset, model_name="llama3:8b", device=local, seed=SEED, max_batch=1, max_seqs=2, start_scheduler=True)
    init_s = time.time() - t0
    cfg = e.cfg
    p_read, kv_tok = model_bytes(cfg)
    ctx = min(args.ctx, cfg["max_seq_len"] - (K + W + 8))
    ids = prompt_ids(ctx, cfg["vocab_size"])

    # ---- prefill (tcgen05 path) through the C-ABI with a host prompt
    s = e.seq_create()
    t0 = time.time()
    lg = e.prefill(s, ids)
    prefill_ms_first = (time.time() - t0) * 1e3
    first = int(lg.argmax())
    # ---- warm-up decode steps (also captures the CUDA graph)
    wids, _ = e.decode_greedy(s, first, W)
    nxt = int(wids[-1])
    launches0 = e.stats()["kernel_launches"]
    # ---- timed region: exactly K steps, device time from CUDA events on the launching stream
    barrier(dist, local)
    clk = ClockSampler(local)
    clk.start()
    out_ids, ms = e.decode_greedy(s, nxt, K)
    barrier(dist, local)
    clocks = clk.stop()
    launches = e.stats()["kernel_launches"] - launches0
    ms_max = max_over_ranks(dist, ms, local)
    value = n_gpus * K / (ms_max * 1e-3)
    mean_ctx = ctx + W + K / 2.0
    step_bytes = 2 * p_read + kv_tok * (mean_ctx + 1)
    achieved = step_bytes / (ms / K * 1e-3) / 1e9
    e.seq_free(s)

    # ---- e2e: the request path (cl_generate_ids -> continuous-batching scheduler), HOST prompt in, HOST ids out,
    #      every step's token read back to the host; prefill measured separately inside the same call
    r = e.generate_ids(ids, eng.greedy(K, ignore_eos=True))     # warm (prefill workspace etc. already hot)
    barrier(dist, local)
    r = e.generate_ids(ids, eng.greedy(K, ignore_eos=True))
    barrier(dist, local)
    dec_s = max_over_ranks(dist, r.decode_ns * 1e-9, local)
    e2e_val = n_gpus * (r.n_generated - 1) / dec_s
    req_s = n_gpus / max_over_ranks(dist, r.total_ns * 1e-9, local)
    prefill_ms = r.prefill_ns * 1e-6
    prefill_flops = 2.0 * ctx * (p_read - cfg["vocab_size"] * cfg["d_model"]) + 4.0 * cfg["n_layers"] * cfg["n_heads"] * cfg["head_dim"] * ctx * ctx / 2

    # ---- dominant kernel live: gate|up GEMV (48% of the step's bytes), weights rotated over > L2
    kern = None
    if rank == 0 and not args.no_kernel_bench:
        try:
            n_gu, k_gu = 2 * cfg["d_ff"], cfg["d_model"]
            rng = np.random.default_rng(0)
            wgu = (rng.integers(0, 1 << 16, size=(n_gu, k_gu), dtype=np.uint16) & 0xBFFF)
            xv = rng.standard_normal(k_gu).astype(np.float32)
            variant = 1 if os.environ.get("CL_GEMV_VARIANT", "1") != "0" else 0
            _, kms = eng.op_gemv(wgu, xv, variant=variant, iters=40, device=local)
            kgbs = n_gu * k_gu * 2 / (kms * 1e-3) / 1e9
            kern = {"name": "gemv gate|up [28672x4096] bf16" if n_gu == 28672 else f"gemv gate|up [{n_gu}x{k_gu}]",
                    "variant": "tma-ring" if variant == 1 else "ldg", "ms": round(kms, 5), "achieved": round(kgbs, 1), "peak": hbm,
                    "unit": "GB/s", "frac": round(kgbs / hbm, 4), "share_of_step_bytes": round(cfg["n_layers"] * n_gu * k_gu * 2 / step_bytes, 3)}
        except Exception as ex:  # noqa: BLE001
            kern = {"error": str(ex)}
    e.close()
    grp.close()

    if rank != 0:
        return 0
    cb = None
    if n_gpus == 1 and not args.no_cpu_baseline:
        try:
            cb, _m, _s = cpu_baseline()
            del _m, _s
        except Exception as ex:  # noqa: BLE001
            cb = {"value": None, "unit": "tokens/s", "cores": os.cpu_count(), "kind": "port", "sample": f"failed: {ex}"}
    line = {
        "metric": METRIC, "value": round(value, 2), "unit": "tokens/s", "n_gpus": n_gpus, "steps": K, "warmup": W,
        "ms_per_step": round(ms_max / K, 4), "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16",
        "data": "synthetic",
        "config": {"workload": f"configs[2]: Llama-3-8B bf16, {ctx}-token prefill then 1-token greedy decode loop, batch 1 per GPU",
                   "preset": args.preset, "ctx": ctx, "weights": f"synthetic counter-based seed {SEED}", "page_size": 32,
                   "replicas": n_gpus, "l2": "inputs larger than L2: every step streams 15 GB of weights (L2 = 126 MB)",
                   "decode_path": os.environ.get("CL_GEMV_VARIANT", "1"), "pdl": os.environ.get("CL_PDL", "1")},
        "per_gpu_tokens_per_s": round(value / n_gpus, 2),
        "roofline": {"bound": "hbm", "achieved": round(achieved, 1), "peak": hbm, "unit": "GB/s", "frac": round(achieved / hbm, 4),
                     "traffic": None, "peak_source": peak_src, "scope": "whole token step (one CUDA-graph launch)",
                     "algorithmic_bytes_per_step": int(step_bytes), "dominant_kernel": kern},
        "cpu_baseline": cb,
        "e2e": {"value": round(e2e_val, 2), "unit": "tokens/s", "h2d_bytes_per_step": round(ctx * 4 / K, 1),
                "d2h_bytes_per_step": 4 * 2, "path": "cl_generate_ids -> scheduler; host prompt ids in, one token id read back per step",
                "prefill_ms": round(prefill_ms, 2), "prefill_tflops": round(prefill_flops / (prefill_ms * 1e-3) / 1e12, 1),
                "requests_per_s": round(req_s, 4)},
        "gpu_launches": int(launches), "clocks": clocks,
        "extra": {"init_s": round(init_s, 1), "first_prefill_ms_incl_workspace_alloc": round(prefill_ms_first, 1),
                  "bf16_tflops_sustained_peak": tflops},
    }
    print(json.dumps(line), flush=True)
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=256)
    ap.add_argument("--warmup", type=int, default=8)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--preset", default=PRESET)
    ap.add_argument("--ctx", type=int, default=CTX)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-kernel-bench", action="store_true")
    args = ap.parse_args()
    return run_reference(args) if args.impl == "reference" else run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
