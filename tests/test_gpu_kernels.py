"""Kernel-level parity (through the C-ABI cl_op_* entry points, host buffers) against the CPU oracle."""
import numpy as np
import pytest

from crowdllama_b200 import engine as eng
from oracle import oracle as oc

pytestmark = pytest.mark.gpu


def _rand_bf16(rng, shape, scale=0.02):
    return oc.np_bf16_from_f32((rng.standard_normal(shape) * scale).astype(np.float32))


def _tol(ref):
    return 2e-3 * float(np.sqrt((ref.astype(np.float64) ** 2).mean()) + 1e-9)


def test_synth_weights_bit_exact():
    for seed, key, n in [(1234, 0, 5000), (7, 16 * 31 + 11, 70001), (2**63 + 5, 1, 4096)]:
        got = eng.op_synth_weights(seed, key, n, float(oc.LINEAR_SCALE))
        np.testing.assert_array_equal(got, oc.np_synth_bf16(seed, key, 0, n))


SHAPES = [(64, 64), (512, 1024), (1536, 2048), (6144, 4096), (4096, 14336), (2048, 8192), (1000, 272)]


@pytest.mark.parametrize("variant", [0, 1])
@pytest.mark.parametrize("n,k", SHAPES)
def test_gemv_matches_oracle(variant, n, k):
    if variant == 1 and k not in (1024, 2048, 4096, 8192, 14336):
        pytest.skip("ring kernel is specialised for K in {1024,2048,4096,8192,14336}")
    rng = np.random.default_rng(n * 31 + k)
    w = _rand_bf16(rng, (n, k))
    x = oc.np_bf16_round(rng.standard_normal(k).astype(np.float32))
    ref = oc.gemv(w, x)
    got = eng.op_gemv(w, x, variant=variant)
    assert np.abs(got - ref).max() <= _tol(ref)
    resid = rng.standard_normal(n).astype(np.float32)
    got_r = eng.op_gemv_residual(w, x, resid, variant=variant)
    assert np.abs(got_r - (resid + ref)).max() <= _tol(ref) + 1e-6


@pytest.mark.parametrize("variant", [0, 1])
@pytest.mark.parametrize("n,k", [(512, 1024), (6144, 4096), (2560, 2048)])
def test_rmsnorm_gemv_and_gateup(variant, n, k):
    rng = np.random.default_rng(n + k + variant)
    w = _rand_bf16(rng, (n, k))
    h = (rng.standard_normal(k) * 3).astype(np.float32)
    g = (1 + 0.1 * rng.standard_normal(k)).astype(np.float32)
    xn = oc.rmsnorm(h, g, 1e-5, round_bf16=True)
    ref = oc.gemv(w, xn)
    got = eng.op_rmsnorm_gemv(w, h, g, 1e-5, variant=variant)
    # a handful of xn elements may round to the neighbouring bf16 value (different fp32 summation
    # order of the mean square): allow a bf16-ulp-sized band
    assert np.abs(got - ref).max() <= 10 * _tol(ref)
    gate, up = ref[0::2], ref[1::2]
    act_ref = oc.np_bf16_round((gate / (1 + np.exp(-gate)) * up).astype(np.float32))
    act = eng.op_rmsnorm_gateup(w, h, g, 1e-5, variant=variant)
    assert np.abs(act - act_ref).max() <= 10 * _tol(act_ref) + 2 ** -8 * np.abs(act_ref).max()


@pytest.mark.parametrize("n_heads,n_kv,hd", [(32, 8, 128), (32, 4, 64), (4, 2, 64), (2, 1, 64), (8, 8, 128)])
@pytest.mark.parametrize("ctx", [1, 2, 31, 32, 33, 257, 1500, 4100])
def test_attn_decode_matches_oracle(n_heads, n_kv, hd, ctx):
    """Split-KV paged attention over tokens 0..ctx-1 (TMA-staged pages, scrambled block table)."""
    rng = np.random.default_rng(ctx * 7 + n_heads + hd)
    q = oc.np_bf16_round(rng.standard_normal(n_heads * hd).astype(np.float32))
    kc = _rand_bf16(rng, (ctx, n_kv, hd), 1.0)
    vc = _rand_bf16(rng, (ctx, n_kv, hd), 1.0)
    out = eng.op_attn_decode(q, kc, vc, n_heads, n_kv, hd, page_size=32)
    ref = oc.np_bf16_round(oc.attention(q, oc.np_f32_from_bf16(kc), oc.np_f32_from_bf16(vc), n_heads, n_kv, hd))
    assert np.isfinite(out).all()      # the op NaN-fills cache slots past ctx: they must never leak into the result
    assert np.abs(out - ref).max() <= 2 ** -7 * max(1.0, np.abs(ref).max())


def test_attn_decode_page_sizes():
    rng = np.random.default_rng(5)
    n_heads, n_kv, hd, ctx = 32, 8, 128, 700
    q = oc.np_bf16_round(rng.standard_normal(n_heads * hd).astype(np.float32))
    kc = _rand_bf16(rng, (ctx, n_kv, hd), 1.0)
    vc = _rand_bf16(rng, (ctx, n_kv, hd), 1.0)
    ref = oc.np_bf16_round(oc.attention(q, oc.np_f32_from_bf16(kc), oc.np_f32_from_bf16(vc), n_heads, n_kv, hd))
    for p in (16, 32, 64):
        out = eng.op_attn_decode(q, kc, vc, n_heads, n_kv, hd, page_size=p)
        assert np.abs(out - ref).max() <= 2 ** -7 * max(1.0, np.abs(ref).max())


@pytest.mark.parametrize("variant", [0, 1])
@pytest.mark.parametrize("d,n_heads,n_kv,hd,pos", [(4096, 32, 8, 128, 777), (2048, 32, 4, 64, 0), (1024, 4, 2, 64, 33)])
def test_qkv_rope_append_matches_oracle(variant, d, n_heads, n_kv, hd, pos):
    """Fused RMSNorm + q|k|v GEMV + RoPE + bf16 round + KV append (the EPI_QKV epilogue)."""
    rng = np.random.default_rng(d + pos)
    theta = 5e5
    qd, kvd = n_heads * hd, n_kv * hd
    w = _rand_bf16(rng, (qd + 2 * kvd, d))
    h = (rng.standard_normal(d) * 2).astype(np.float32)
    g = (1 + 0.1 * rng.standard_normal(d)).astype(np.float32)
    q, k, v = eng.op_qkv_rope_append(w, h, g, 1e-5, n_heads, n_kv, hd, pos, theta, variant=variant)
    xn = oc.rmsnorm(h, g, 1e-5, round_bf16=True)
    y = oc.gemv(w, xn)
    q_ref = oc.np_bf16_round(oc.rope(y[:qd], n_heads, hd, pos, theta))
    k_ref = oc.np_bf16_round(oc.rope(y[qd:qd + kvd], n_kv, hd, pos, theta))
    v_ref = oc.np_bf16_round(y[qd + kvd:])
    tol = 2 ** -6 * max(1.0, float(np.abs(y).max()))     # bf16 rounding of slightly different fp32 sums
    assert np.abs(q - q_ref).max() <= tol
    assert np.abs(oc.np_f32_from_bf16(k) - k_ref).max() <= tol
    assert np.abs(oc.np_f32_from_bf16(v) - v_ref).max() <= tol
