"""ctypes binding of the CPU oracle (oracle/llama_oracle.c).  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module.  Parity is UNPINNED against the reference binary (see llama_oracle.h); the
restatement is pinned against HF transformers through tests/golden/.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from pathlib import Path

import numpy as np

_HERE = Path(__file__).resolve().parent
_LIB_PATH = _HERE / "_build" / "liboracle.so"

KINDS = dict(EMBED=0, LM_HEAD=1, FINAL_NORM=2, ATTN_NORM=3, WQ=4, WK=5, WV=6, WO=7,
             FFN_NORM=8, WGATE=9, WUP=10, WDOWN=11)
LINEAR_SCALE = np.float32(1.35e-4)
NORM_SCALE = np.float32(1.0 / 4096.0)


class OcConfig(C.Structure):
    _fields_ = [(n, C.c_int32) for n in ("n_layers", "d_model", "n_heads", "n_kv_heads", "head_dim",
                                         "d_ff", "vocab_size", "max_seq_len")] + \
               [("rope_theta", C.c_float), ("rms_eps", C.c_float), ("rope_factor", C.c_float), ("rope_low_freq_factor", C.c_float),
                ("rope_high_freq_factor", C.c_float), ("rope_original_max_pos", C.c_int32)]


class OcSampling(C.Structure):
    _fields_ = [("temperature", C.c_float), ("top_k", C.c_int32), ("top_p", C.c_float),
                ("repeat_penalty", C.c_float), ("repeat_last_n", C.c_int32), ("seed", C.c_uint64)]


def build(force: bool = False) -> Path:
    src_m = max((_HERE / f).stat().st_mtime for f in ("llama_oracle.c", "llama_oracle.h", "Makefile"))
    if force or not _LIB_PATH.exists() or _LIB_PATH.stat().st_mtime < src_m:
        if (_HERE / "llama_oracle.c").exists() and _which("gcc"):
            subprocess.run(["make", "-C", str(_HERE)], check=True, capture_output=True)
    return _LIB_PATH


def _which(x):
    from shutil import which
    return which(x)


_lib = None


def effective_cpus() -> int:
    """CPUs this process may really use: affinity mask capped by the cgroup CPU quota (a container can
    show 128 cores in nproc while being throttled to a few; spinning OpenMP threads then crawl)."""
    n = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    try:
        quota, period = Path("/sys/fs/cgroup/cpu.max").read_text().split()[:2]          # cgroup v2
        if quota != "max":
            n = min(n, max(1, int(int(quota) / int(period))))
    except Exception:  # noqa: BLE001
        try:                                                                             # cgroup v1
            quota = int(Path("/sys/fs/cgroup/cpu/cpu.cfs_quota_us").read_text())
            period = int(Path("/sys/fs/cgroup/cpu/cpu.cfs_period_us").read_text())
            if quota > 0 and period > 0:
                n = min(n, max(1, quota // period))
        except Exception:  # noqa: BLE001
            pass
    return max(1, n)


def lib():
    global _lib
    if _lib is None:
        build()
        os.environ.setdefault("OMP_WAIT_POLICY", "passive")   # do not spin between parallel regions
        os.environ.setdefault("OMP_PROC_BIND", "false")
        L = C.CDLL(str(_LIB_PATH))
        vp, i32, i64, u64, f32 = C.c_void_p, C.c_int32, C.c_int64, C.c_uint64, C.c_float
        P = C.POINTER
        sig = {
            "oc_bf16_from_f32": (C.c_uint16, [f32]),
            "oc_synth_int": (i32, [u64, i32, u64]),
            "oc_synth_bf16": (None, [u64, i32, u64, u64, f32, vp]),
            "oc_model_create": (vp, [P(OcConfig)]),
            "oc_model_destroy": (None, [vp]),
            "oc_model_fill_synthetic": (None, [vp, u64]),
            "oc_model_set_tensor": (C.c_int, [vp, i32, i32, vp, i64]),
            "oc_model_set_act_rounding": (None, [vp, i32]),
            "oc_set_num_threads": (None, [i32]),
            "oc_get_num_threads": (i32, []),
            "oc_seq_create": (vp, [vp, i32]),
            "oc_seq_destroy": (None, [vp]),
            "oc_seq_len": (i32, [vp]),
            "oc_seq_truncate": (None, [vp, i32]),
            "oc_seq_fake_fill": (None, [vp, i32]),
            "oc_forward": (C.c_int, [vp, vp, vp, i32, vp, i32]),
            "oc_prefill_block": (C.c_int, [vp, vp, vp, i32, vp]),
            "oc_seq_kv": (C.c_int, [vp, i32, i32, i32, i32, vp]),
            "oc_debug_hidden": (C.c_int, [vp, i32, vp]),
            "oc_argmax": (i32, [vp, i32]),
            "oc_debug_vec": (C.c_int, [vp, i32, vp]),
            "oc_greedy": (C.c_int, [vp, vp, i32, i32, vp, vp]),
            "oc_sample": (i32, [vp, i32, P(OcSampling), vp, i32, u64]),
            "oc_op_gemv": (None, [vp, vp, vp, i32, i32]),
            "oc_op_rmsnorm": (None, [vp, vp, f32, i32, i32, vp]),
            "oc_op_rope": (None, [vp, i32, i32, i32, f32]),
            "oc_op_attention": (None, [vp, vp, vp, i32, i32, i32, i32, vp]),
            "oc_op_gemm": (None, [vp, vp, vp, i32, i32, i32]),
        }
        for name, (res, args) in sig.items():
            fn = getattr(L, name)
            fn.restype, fn.argtypes = res, args
        L.oc_set_num_threads(min(effective_cpus(), int(os.environ.get("CL_ORACLE_THREADS", "32"))))
        _lib = L
    return _lib


def _ptr(a: np.ndarray):
    return a.ctypes.data_as(C.c_void_p)


# ---- numpy mirror of the scalar pieces (used to cross-check the C code) ------------------------
def np_bf16_from_f32(x: np.ndarray) -> np.ndarray:
    u = np.ascontiguousarray(x, dtype=np.float32).view(np.uint32).astype(np.uint64)
    lsb = (u >> 16) & 1
    return (((u + 0x7FFF + lsb) >> 16) & 0xFFFF).astype(np.uint16)


def np_f32_from_bf16(h: np.ndarray) -> np.ndarray:
    return (h.astype(np.uint32) << 16).view(np.float32)


def np_bf16_round(x: np.ndarray) -> np.ndarray:
    return np_f32_from_bf16(np_bf16_from_f32(x))


def np_synth_int(seed: int, key: int, first: int, n: int) -> np.ndarray:
    M = np.uint64
    with np.errstate(over="ignore"):
        idx = np.arange(first, first + n, dtype=np.uint64)
        x = M(seed) * M(0x9E3779B97F4A7C15) + ((M(key) << M(40)) | idx)
        x ^= x >> M(30); x *= M(0xBF58476D1CE4E5B9)
        x ^= x >> M(27); x *= M(0x94D049BB133111EB)
        x ^= x >> M(31)
    r = (x & M(0xFFFFFFFF)).astype(np.uint32)
    s = (r & 0xFF).astype(np.int32) + ((r >> 8) & 0xFF).astype(np.int32) + \
        ((r >> 16) & 0xFF).astype(np.int32) + (r >> 24).astype(np.int32)
    return s - 510


def np_synth_bf16(seed: int, key: int, first: int, n: int, scale=LINEAR_SCALE) -> np.ndarray:
    return np_bf16_from_f32(np_synth_int(seed, key, first, n).astype(np.float32) * np.float32(scale))


# ---- object wrappers ----------------------------------------------------------------------------
PRESETS = {
    "llama3-8b": dict(n_layers=32, d_model=4096, n_heads=32, n_kv_heads=8, head_dim=128, d_ff=14336,
                      vocab_size=128256, max_seq_len=8192, rope_theta=5e5, rms_eps=1e-5),
    "mistral-7b": dict(n_layers=32, d_model=4096, n_heads=32, n_kv_heads=8, head_dim=128, d_ff=14336,
                       vocab_size=32000, max_seq_len=8192 + 512, rope_theta=1e6, rms_eps=1e-5),
    "tinyllama-1.1b": dict(n_layers=22, d_model=2048, n_heads=32, n_kv_heads=4, head_dim=64, d_ff=5632,
                           vocab_size=32000, max_seq_len=2048, rope_theta=1e4, rms_eps=1e-5),
    # "llama3" rotary frequency scaling (Llama-3.1 / 3.2), tied embeddings in the real checkpoints
    "llama3.1-8b": dict(n_layers=32, d_model=4096, n_heads=32, n_kv_heads=8, head_dim=128, d_ff=14336, vocab_size=128256,
                        max_seq_len=32768, rope_theta=5e5, rms_eps=1e-5, rope_factor=8.0, rope_low_freq_factor=1.0,
                        rope_high_freq_factor=4.0, rope_original_max_pos=8192),
    "llama3.2-1b": dict(n_layers=16, d_model=2048, n_heads=32, n_kv_heads=8, head_dim=64, d_ff=8192, vocab_size=128256,
                        max_seq_len=8192, rope_theta=5e5, rms_eps=1e-5, rope_factor=32.0, rope_low_freq_factor=1.0,
                        rope_high_freq_factor=4.0, rope_original_max_pos=8192),
    "tiny-test": dict(n_layers=2, d_model=256, n_heads=4, n_kv_heads=2, head_dim=64, d_ff=512,
                      vocab_size=512, max_seq_len=512, rope_theta=1e4, rms_eps=1e-5),
}


class Model:
    def __init__(self, cfg: dict | str, seed: int | None = None):
        if isinstance(cfg, str):
            cfg = PRESETS[cfg]
        self.cfg = dict(cfg)
        self._c = OcConfig(**cfg)
        self._h = lib().oc_model_create(C.byref(self._c))
        if seed is not None:
            lib().oc_model_fill_synthetic(self._h, seed)

    def close(self):
        if self._h:
            lib().oc_model_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_tensor(self, layer: int, kind: str, bf16: np.ndarray):
        a = np.ascontiguousarray(bf16, dtype=np.uint16)
        rc = lib().oc_model_set_tensor(self._h, layer, KINDS[kind], _ptr(a), a.size)
        if rc:
            raise ValueError(f"set_tensor({layer},{kind}) rc={rc} n={a.size}")

    def set_act_rounding(self, mode: int):
        lib().oc_model_set_act_rounding(self._h, mode)

    def new_seq(self, max_len: int | None = None) -> "Seq":
        return Seq(self, max_len or self.cfg["max_seq_len"])

    def debug_vec(self, which: int) -> np.ndarray:
        n = {0: self.cfg["n_heads"] * self.cfg["head_dim"], 1: self.cfg["n_heads"] * self.cfg["head_dim"], 2: self.cfg["d_ff"],
             3: self.cfg["d_model"]}[which]
        out = np.empty(n, np.float32)
        lib().oc_debug_vec(self._h, which, _ptr(out))
        return out

    def hidden(self, layer: int) -> np.ndarray:
        out = np.empty(self.cfg["d_model"], np.float32)
        lib().oc_debug_hidden(self._h, layer, _ptr(out))
        return out


class Seq:
    def __init__(self, model: Model, max_len: int):
        self.m = model
        self._h = lib().oc_seq_create(model._h, max_len)

    def close(self):
        if self._h:
            lib().oc_seq_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __len__(self):
        return lib().oc_seq_len(self._h)

    def truncate(self, n: int):
        lib().oc_seq_truncate(self._h, n)

    def fake_fill(self, n: int):
        lib().oc_seq_fake_fill(self._h, n)

    def forward(self, ids, all_logits: bool = False) -> np.ndarray:
        ids = np.ascontiguousarray(ids, dtype=np.int32)
        V = self.m.cfg["vocab_size"]
        out = np.empty((len(ids), V) if all_logits else (V,), np.float32)
        rc = lib().oc_forward(self.m._h, self._h, _ptr(ids), len(ids), _ptr(out), 1 if all_logits else 0)
        if rc:
            raise RuntimeError(f"oc_forward rc={rc}")
        return out

    def prefill_block(self, ids) -> np.ndarray:
        """Layer-major prefill (bit-identical to forward(ids), fast enough for thousands of tokens); last-position logits."""
        ids = np.ascontiguousarray(ids, dtype=np.int32)
        out = np.empty(self.m.cfg["vocab_size"], np.float32)
        rc = lib().oc_prefill_block(self.m._h, self._h, _ptr(ids), len(ids), _ptr(out))
        if rc:
            raise RuntimeError(f"oc_prefill_block rc={rc}")
        return out

    def kv(self, layer: int, which: int, t0: int, n: int) -> np.ndarray:
        """Cached K (which=0) / V (which=1) of `layer`, tokens t0..t0+n-1 as [n][n_kv*head_dim] fp32."""
        out = np.empty((n, self.m.cfg["n_kv_heads"] * self.m.cfg["head_dim"]), np.float32)
        rc = lib().oc_seq_kv(self._h, layer, which, t0, n, _ptr(out))
        if rc:
            raise RuntimeError(f"oc_seq_kv rc={rc}")
        return out

    def greedy(self, first_id: int, n_steps: int):
        ids = np.empty(n_steps, np.int32)
        margins = np.empty(n_steps, np.float32)
        rc = lib().oc_greedy(self.m._h, self._h, first_id, n_steps, _ptr(ids), _ptr(margins))
        if rc:
            raise RuntimeError(f"oc_greedy rc={rc}")
        return ids, margins


def sample(logits: np.ndarray, temperature=0.8, top_k=40, top_p=0.9, repeat_penalty=1.1, repeat_last_n=64,
           seed=0, history=None, step=0) -> int:
    lg = np.ascontiguousarray(logits, dtype=np.float32)
    sp = OcSampling(temperature, top_k, top_p, repeat_penalty, repeat_last_n, seed)
    hist = np.ascontiguousarray(history if history is not None else [], dtype=np.int32)
    return lib().oc_sample(_ptr(lg), lg.size, C.byref(sp), _ptr(hist) if hist.size else None, hist.size, step)


def gemv(w_bf16: np.ndarray, x: np.ndarray) -> np.ndarray:
    w = np.ascontiguousarray(w_bf16, dtype=np.uint16)
    x = np.ascontiguousarray(x, dtype=np.float32)
    y = np.empty(w.shape[0], np.float32)
    lib().oc_op_gemv(_ptr(w), _ptr(x), _ptr(y), w.shape[0], w.shape[1])
    return y


def gemm(x_bf16: np.ndarray, w_bf16: np.ndarray) -> np.ndarray:
    x = np.ascontiguousarray(x_bf16, dtype=np.uint16)
    w = np.ascontiguousarray(w_bf16, dtype=np.uint16)
    y = np.empty((x.shape[0], w.shape[0]), np.float32)
    lib().oc_op_gemm(_ptr(x), _ptr(w), _ptr(y), x.shape[0], w.shape[0], w.shape[1])
    return y


def rmsnorm(h: np.ndarray, gain: np.ndarray, eps: float, round_bf16: bool = True) -> np.ndarray:
    h = np.ascontiguousarray(h, dtype=np.float32)
    g = np.ascontiguousarray(gain, dtype=np.float32)
    out = np.empty_like(h)
    lib().oc_op_rmsnorm(_ptr(h), _ptr(g), eps, h.size, 1 if round_bf16 else 0, _ptr(out))
    return out


def rope(v: np.ndarray, n_heads: int, head_dim: int, pos: int, theta: float) -> np.ndarray:
    out = np.ascontiguousarray(v, dtype=np.float32).copy()
    lib().oc_op_rope(_ptr(out), n_heads, head_dim, pos, theta)
    return out


def attention(q: np.ndarray, kc: np.ndarray, vc: np.ndarray, n_heads: int, n_kv: int, head_dim: int) -> np.ndarray:
    q = np.ascontiguousarray(q, dtype=np.float32)
    kc = np.ascontiguousarray(kc, dtype=np.float32)
    vc = np.ascontiguousarray(vc, dtype=np.float32)
    ctx = kc.shape[0]
    out = np.empty(n_heads * head_dim, np.float32)
    lib().oc_op_attention(_ptr(q), _ptr(kc), _ptr(vc), ctx, n_heads, n_kv, head_dim, _ptr(out))
    return out


def set_threads(n: int):
    lib().oc_set_num_threads(n)


def num_threads() -> int:
    return lib().oc_get_num_threads()
