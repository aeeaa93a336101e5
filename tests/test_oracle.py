"""Pin the CPU oracle: golden vectors from HF transformers (tests/golden/make_golden.py), known
answers of the synthetic weight generator, and internal consistency of the restatement."""
from pathlib import Path

import numpy as np
import pytest

from st_util import fixture_cfg

from oracle import oracle as oc

G = Path(__file__).resolve().parent / "golden"


def _load_hf(name):
    z = np.load(G / name)
    cfg = fixture_cfg(z)
    m = oc.Model(cfg)
    m.set_tensor(0, "EMBED", z["embed"])
    m.set_tensor(0, "LM_HEAD", z["lm_head"])
    m.set_tensor(0, "FINAL_NORM", z["final_norm"])
    for l in range(cfg["n_layers"]):
        for k in ("ATTN_NORM", "FFN_NORM", "WQ", "WK", "WV", "WO", "WGATE", "WUP", "WDOWN"):
            m.set_tensor(l, k, z[f"L{l}.{k}"])
    return z, cfg, m


@pytest.mark.parametrize("fixture", ["hf_tiny_llama.npz", "hf_tiny_mistral.npz", "hf_tiny_llama3geom.npz", "hf_tiny_llama31rope.npz"])
def test_oracle_matches_hf_fp32(fixture):
    """act_rounding=0 (pure fp32 activations) must reproduce HF float32 logits: pins RoPE pairing,
    GQA head mapping, norm placement, SwiGLU and the untied LM head."""
    z, cfg, m = _load_hf(fixture)
    m.set_act_rounding(0)
    s = m.new_seq()
    logits = s.forward(z["ids"], all_logits=True)
    ref = z["logits"]
    assert logits.shape == ref.shape
    err = np.abs(logits - ref).max()
    assert err < 2e-4 * max(1.0, np.abs(ref).max()), err
    assert (logits.argmax(-1) == ref.argmax(-1)).all()


@pytest.mark.parametrize("fixture", ["hf_tiny_llama.npz", "hf_tiny_mistral.npz", "hf_tiny_llama3geom.npz", "hf_tiny_llama31rope.npz"])
def test_oracle_v1_rounding_close_to_hf(fixture):
    """cl-llama v1 numerics (bf16 rounding points) stay within a bf16-sized band of HF fp32."""
    z, cfg, m = _load_hf(fixture)
    s = m.new_seq()
    logits = s.forward(z["ids"], all_logits=True)
    ref = z["logits"]
    rms = float(np.sqrt((ref ** 2).mean()))
    assert np.abs(logits - ref).max() < 5e-2 * rms


def test_incremental_equals_batch():
    m = oc.Model("tiny-test", seed=7)
    ids = [(i * 31 + 5) % 512 for i in range(20)]
    a = m.new_seq().forward(ids, all_logits=True)
    s = m.new_seq()
    s.forward(ids[:11])
    b = s.forward(ids[11:], all_logits=True)
    np.testing.assert_array_equal(a[11:], b)
    s.truncate(5)
    c = s.forward(ids[5:], all_logits=True)
    np.testing.assert_array_equal(a[5:], c)


def test_synth_known_answers():
    z = np.load(G / "synth_kat.npz")
    for name in z.files:
        kind, seed, key, first, n = name.split("_")
        seed, key, first, n = int(seed), int(key), int(first), int(n)
        if kind == "int":
            got = np.array([oc.lib().oc_synth_int(seed, key, first + i) for i in range(n)], np.int32)
        else:
            got = np.empty(n, np.uint16)
            oc.lib().oc_synth_bf16(seed, key, first, n, float(oc.LINEAR_SCALE), got.ctypes.data)
        np.testing.assert_array_equal(got, z[name])


def test_synth_statistics():
    v = oc.np_f32_from_bf16(oc.np_synth_bf16(1234, 4, 0, 1 << 18))
    assert abs(v.mean()) < 2e-4
    assert abs(v.std() - 0.02) < 5e-4


def test_bf16_rounding_matches_numpy_and_torch():
    import torch
    rng = np.random.default_rng(0)
    x = np.concatenate([rng.standard_normal(4096).astype(np.float32) * 3,
                        np.array([0.0, -0.0, 1.0, 1.00390625, 1.0078125, 3.4e38, -3.4e38, 1e-40], np.float32)])
    a = np.array([oc.lib().oc_bf16_from_f32(float(v)) for v in x], np.uint16)
    np.testing.assert_array_equal(a, oc.np_bf16_from_f32(x))
    t = torch.from_numpy(x).to(torch.bfloat16).view(torch.int16).numpy().view(np.uint16)
    np.testing.assert_array_equal(a, t)


def test_greedy_deterministic_and_thread_independent():
    m = oc.Model("tiny-test", seed=3)
    oc.set_threads(1)
    a, ma = m.new_seq().greedy(5, 24)
    oc.set_threads(4)
    b, mb = m.new_seq().greedy(5, 24)
    np.testing.assert_array_equal(a, b)
    np.testing.assert_array_equal(ma, mb)
    assert (ma >= 0).all()


def test_ops_against_numpy():
    rng = np.random.default_rng(1)
    w = oc.np_bf16_from_f32(rng.standard_normal((48, 272)).astype(np.float32))
    x = oc.np_bf16_round(rng.standard_normal(272).astype(np.float32))
    y = oc.gemv(w, x)
    ref = oc.np_f32_from_bf16(w).astype(np.float64) @ x.astype(np.float64)
    np.testing.assert_allclose(y, ref, rtol=1e-5, atol=1e-5)
    h = rng.standard_normal(300).astype(np.float32)
    g = 1 + 0.1 * rng.standard_normal(300).astype(np.float32)
    n = oc.rmsnorm(h, g, 1e-5, round_bf16=False)
    refn = h / np.sqrt((h.astype(np.float64) ** 2).mean() + 1e-5) * g
    np.testing.assert_allclose(n, refn, rtol=2e-6, atol=1e-6)
    v = rng.standard_normal(2 * 64).astype(np.float32)
    r = oc.rope(v, 2, 64, 17, 1e4)
    inv = 1e4 ** (-2 * np.arange(32) / 64)
    c, s_ = np.cos(17 * inv), np.sin(17 * inv)
    vv = v.reshape(2, 64)
    ref_r = np.concatenate([vv[:, :32] * c - vv[:, 32:] * s_, vv[:, 32:] * c + vv[:, :32] * s_], axis=1).ravel()
    np.testing.assert_allclose(r, ref_r, rtol=1e-5, atol=1e-6)
    q = rng.standard_normal(4 * 32).astype(np.float32)
    kc = rng.standard_normal((9, 2, 32)).astype(np.float32)
    vc = rng.standard_normal((9, 2, 32)).astype(np.float32)
    o = oc.attention(q, kc, vc, 4, 2, 32)
    ref_o = []
    for hh in range(4):
        sc = kc[:, hh // 2, :] @ q[hh * 32:(hh + 1) * 32] / np.sqrt(32)
        p = np.exp(sc - sc.max()); p /= p.sum()
        ref_o.append(p @ vc[:, hh // 2, :])
    np.testing.assert_allclose(o, np.concatenate(ref_o), rtol=1e-5, atol=1e-6)


def test_sampler_greedy_and_distribution():
    rng = np.random.default_rng(2)
    lg = rng.standard_normal(1000).astype(np.float32)
    assert oc.sample(lg, temperature=0.0) == int(lg.argmax())
    picks = [oc.sample(lg, temperature=0.8, top_k=40, top_p=0.9, repeat_penalty=1.0, seed=11, step=i)
             for i in range(300)]
    top40 = set(np.argsort(-lg)[:40].tolist())
    assert set(picks) <= top40
    assert len(set(picks)) > 3
    # determinism
    assert picks[:10] == [oc.sample(lg, temperature=0.8, top_k=40, top_p=0.9, repeat_penalty=1.0, seed=11, step=i)
                          for i in range(10)]
    # repeat penalty pushes a dominant token down
    lg2 = np.zeros(50, np.float32); lg2[7] = 0.2
    base = sum(oc.sample(lg2, 1.0, 0, 1.0, 1.0, 64, seed=5, step=i) == 7 for i in range(400))
    pen = sum(oc.sample(lg2, 1.0, 0, 1.0, 1.5, 64, seed=5, history=[7], step=i) == 7 for i in range(400))
    assert pen < base
