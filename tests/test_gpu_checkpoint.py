"""Checkpoint loading behind cl_engine_config.weights_path (csrc/weights_io.cpp): an HF model directory written by
transformers' own save_pretrained (tests/golden/hf_tiny_llama_ckpt, generator tests/golden/make_golden.py) is loaded by
path — architecture from config.json — and must reproduce the HF golden logits and, bit for bit, the logits of the
cl_engine_set_tensor path.  F32 / F16 / sharded variants are written by a hand-rolled safetensors writer."""
import json
from pathlib import Path

import numpy as np
import pytest

from crowdllama_b200 import engine as eng
from st_util import hf_tensors_from_fixture, write_safetensors, fixture_cfg

pytestmark = pytest.mark.gpu
G = Path(__file__).resolve().parent / "golden"


def _fixture_cfg(z):
    return fixture_cfg(z)


def _logits(e, ids):
    s = e.seq_create()
    got = [e.prefill(s, ids[:1])]
    for t in ids[1:]:
        got.append(e.decode_step(s, int(t))[0])
    e.seq_free(s)
    return np.stack(got)


def test_hf_checkpoint_directory_loads_by_path_and_matches_golden(tmp_path):
    z = np.load(G / "hf_tiny_llama.npz")
    cfg = _fixture_cfg(z)
    ids, ref = z["ids"], z["logits"]
    with eng.Engine(model=cfg, decode_path=1) as e:              # reference: the tensor-by-tensor path
        e.set_tensor(0, "EMBED", z["embed"]); e.set_tensor(0, "LM_HEAD", z["lm_head"]); e.set_tensor(0, "FINAL_NORM", z["final_norm"])
        for l in range(cfg["n_layers"]):
            for k in ("ATTN_NORM", "FFN_NORM", "WQ", "WK", "WV", "WO", "WGATE", "WUP", "WDOWN"):
                e.set_tensor(l, k, z[f"L{l}.{k}"])
        by_tensor = _logits(e, ids)
    # 1. the directory transformers wrote: config.json decides the architecture (no preset, no model config)
    with eng.Engine(weights_path=G / "hf_tiny_llama_ckpt", decode_path=1) as e:
        assert {k: e.cfg[k] for k in ("n_layers", "d_model", "n_heads", "n_kv_heads", "head_dim", "d_ff", "vocab_size")} == \
               {k: cfg[k] for k in ("n_layers", "d_model", "n_heads", "n_kv_heads", "head_dim", "d_ff", "vocab_size")}
        got = _logits(e, ids)
    np.testing.assert_array_equal(got, by_tensor)                 # same bf16 weights -> bit-identical logits
    rms = float(np.sqrt((ref ** 2).mean()))
    assert np.abs(got - ref).max() < 5e-2 * rms                   # vs HF transformers fp32 (bf16 rounding points)
    assert (got.argmax(-1) == ref.argmax(-1)).mean() >= 0.9
    # 2. a single file + explicit config; F32 and F16 sources are rounded to the same bf16 values
    tensors = hf_tensors_from_fixture(z, cfg)
    for dt in ("F32", "F16"):
        f = tmp_path / f"m_{dt}.safetensors"
        write_safetensors(f, tensors, dt)
        with eng.Engine(model=cfg, weights_path=f, decode_path=1) as e:
            got2 = _logits(e, ids)
        if dt == "F32":
            np.testing.assert_array_equal(got2, by_tensor)
        else:
            assert np.abs(got2 - by_tensor).max() < 2e-2 * rms    # fp16 cannot hold every bf16 value exactly (range)
    # 3. sharded directory (two files), tied embeddings (no lm_head.weight in the checkpoint)
    d = tmp_path / "sharded"
    d.mkdir()
    names = [n for n in tensors if n != "lm_head.weight"]
    write_safetensors(d / "model-00001-of-00002.safetensors", {n: tensors[n] for n in names[: len(names) // 2]})
    write_safetensors(d / "model-00002-of-00002.safetensors", {n: tensors[n] for n in names[len(names) // 2:]})
    (d / "config.json").write_text((G / "hf_tiny_llama_ckpt" / "config.json").read_text())
    with eng.Engine(weights_path=d, decode_path=1) as e:
        tied = _logits(e, ids)
    with eng.Engine(model=cfg, decode_path=1) as e:
        for n_, k_ in (("EMBED", "embed"), ("LM_HEAD", "embed"), ("FINAL_NORM", "final_norm")):
            e.set_tensor(0, n_, z[k_])
        for l in range(cfg["n_layers"]):
            for k in ("ATTN_NORM", "FFN_NORM", "WQ", "WK", "WV", "WO", "WGATE", "WUP", "WDOWN"):
                e.set_tensor(l, k, z[f"L{l}.{k}"])
        np.testing.assert_array_equal(tied, _logits(e, ids))
    # 4. errors are loud
    with pytest.raises(eng.EngineError):
        eng.Engine(model=cfg, weights_path=tmp_path / "nope.safetensors")
    bad = dict(cfg); bad["d_ff"] = cfg["d_ff"] * 2
    with pytest.raises(eng.EngineError):
        eng.Engine(model=bad, weights_path=G / "hf_tiny_llama_ckpt" / "model.safetensors")


def test_llama31_style_checkpoint_rope_scaling_and_tied_embeddings():
    """A directory written by transformers for a Llama-3.1 / 3.2 style config — rope_type "llama3" and
    tie_word_embeddings — loads by path (scaling parameters from config.json, lm_head from embed_tokens) and reproduces
    the HF fp32 logits; without the scaling the same weights do not."""
    z = np.load(G / "hf_tiny_llama31rope.npz")
    cfg = fixture_cfg(z)
    ids, ref = z["ids"], z["logits"]
    rms = float(np.sqrt((ref ** 2).mean()))
    with eng.Engine(weights_path=G / "hf_tiny_llama31rope_ckpt", decode_path=1) as e:
        assert (e.cfg["rope_factor"], e.cfg["rope_original_max_pos"]) == (8.0, 16)
        got = _logits(e, ids)
    assert np.abs(got - ref).max() < 5e-2 * rms
    assert (got.argmax(-1) == ref.argmax(-1)).mean() >= 0.9
    plain = dict(cfg); plain["rope_factor"] = 0.0
    with eng.Engine(model=plain, weights_path=G / "hf_tiny_llama31rope_ckpt" / "model.safetensors", decode_path=1) as e:
        off = _logits(e, ids)
    assert np.abs(off - ref).max() > 0.2 * rms                    # the fixture really exercises the scaling
