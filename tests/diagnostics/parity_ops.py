"""Diagnostic: actual error of every decode op vs the oracle (relative to the rms of the reference)."""
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
from crowdllama_b200 import engine as eng  # noqa: E402
from oracle import oracle as oc  # noqa: E402

rng = np.random.default_rng(0)


def rb(shape, scale=0.02):
    return oc.np_bf16_from_f32((rng.standard_normal(shape) * scale).astype(np.float32))


def rel(a, b):
    return float(np.abs(a - b).max() / (np.sqrt((b.astype(np.float64) ** 2).mean()) + 1e-12))


d, F, H, KV, HD = 4096, 14336, 32, 8, 128
w = rb((6144, d)); x = oc.np_bf16_round(rng.standard_normal(d).astype(np.float32))
for v in (0, 1):
    print(f"gemv v{v} [6144x4096]: max err / rms = {rel(eng.op_gemv(w, x, variant=v), oc.gemv(w, x)):.3e}")
h = (rng.standard_normal(d) * 3).astype(np.float32); g = (1 + 0.1 * rng.standard_normal(d)).astype(np.float32)
xn = oc.rmsnorm(h, g, 1e-5, True)
for v in (0, 1):
    got = eng.op_rmsnorm_gemv(w, h, g, 1e-5, variant=v)
    print(f"rmsnorm+gemv v{v}: {rel(got, oc.gemv(w, xn)):.3e}")
wgu = rb((2 * 2048, d))
ref = oc.gemv(wgu, xn); gate, up = ref[0::2], ref[1::2]
act_ref = oc.np_bf16_round((gate / (1 + np.exp(-gate)) * up).astype(np.float32))
for v in (0, 1):
    print(f"rmsnorm+gateup v{v}: {rel(eng.op_rmsnorm_gateup(wgu, h, g, 1e-5, variant=v), act_ref):.3e}")
wd = rb((d, F)); xa = oc.np_bf16_round(rng.standard_normal(F).astype(np.float32)); res = rng.standard_normal(d).astype(np.float32)
for v in (0, 1):
    print(f"down+resid v{v}: {rel(eng.op_gemv_residual(wd, xa, res, variant=v), res + oc.gemv(wd, xa)):.3e}")
pos = 777
q, k, vv = eng.op_qkv_rope_append(w, h, g, 1e-5, H, KV, HD, pos, 5e5, variant=1)
y = oc.gemv(w, xn)
q_ref = oc.np_bf16_round(oc.rope(y[:4096], H, HD, pos, 5e5)); k_ref = oc.np_bf16_round(oc.rope(y[4096:5120], KV, HD, pos, 5e5))
print(f"qkv+rope: q {rel(q, q_ref):.3e}  k {rel(oc.np_f32_from_bf16(k), k_ref):.3e}  v {rel(oc.np_f32_from_bf16(vv), oc.np_bf16_round(y[5120:])):.3e}")
print(f"   (fraction of q elements that differ: {(q != q_ref).mean():.4f})")
for ctx in (12, 300, 4100):
    qq = oc.np_bf16_round(rng.standard_normal(H * HD).astype(np.float32) * 1.28)
    kc = rb((ctx, KV, HD), 1.28); vc = rb((ctx, KV, HD), 1.28)
    out = eng.op_attn_decode(qq, kc, vc, H, KV, HD)
    ref_f = oc.attention(qq, oc.np_f32_from_bf16(kc), oc.np_f32_from_bf16(vc), H, KV, HD)
    print(f"attention ctx {ctx}: vs rounded ref {rel(out, oc.np_bf16_round(ref_f)):.3e}  frac differing {(out != oc.np_bf16_round(ref_f)).mean():.4f}")
