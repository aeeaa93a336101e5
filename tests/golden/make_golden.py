"""Generate the golden fixtures that pin the CPU oracle (run in the build container, CPU only).

The reference (crowdllama) holds no numerical code and no golden vectors for the model step
(SURVEY.md §8c), and its arithmetic dependency (ollama v0.9.6 / llama.cpp) is not available
offline.  The fixtures below therefore come from the public implementation of the same
architecture that IS importable here: HF transformers' LlamaForCausalLM / MistralForCausalLM,
run on CPU in float32 over bf16-representable seeded weights.

  python tests/golden/make_golden.py      -> tests/golden/hf_tiny_llama.npz, hf_tiny_mistral.npz,
                                              hf_tiny_llama3geom.npz, hf_tiny_llama31rope.npz, synth_kat.npz,
                                              hf_tiny_llama_ckpt/, hf_tiny_llama31rope_ckpt/ (config.json + model.safetensors)

Fixtures are small (< 1 MB each) and committed; tests never import transformers.
"""
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
from oracle import oracle as oc  # noqa: E402

OUT = Path(__file__).resolve().parent


def hf_fixture(kind: str, path: Path, seed: int, ckpt_dir: Path | None = None):
    from transformers import LlamaConfig, LlamaForCausalLM, MistralConfig, MistralForCausalLM
    # head_dim 64 and heads/kv in {2, 4} so the same fixtures also drive the CUDA engine
    # (its attention kernels are built for head_dim 64|128).
    if kind == "llama":
        cfg = dict(n_layers=2, d_model=128, n_heads=2, n_kv_heads=1, head_dim=64, d_ff=256, vocab_size=256,
                   max_seq_len=64, rope_theta=10000.0, rms_eps=1e-5)
    elif kind == "llama3geom":
        # the head geometry of Llama-3-8B / Mistral-7B (head_dim 128, 4 query heads per kv head, rope theta 5e5)
        # on a small residual width: pins the oracle's RoPE pairing and GQA grouping for exactly that geometry
        cfg = dict(n_layers=2, d_model=128, n_heads=4, n_kv_heads=1, head_dim=128, d_ff=256, vocab_size=256,
                   max_seq_len=64, rope_theta=500000.0, rms_eps=1e-5)
    elif kind == "llama31rope":
        # Llama-3.1 / 3.2 style: "llama3" rotary scaling (scaled to a 16-position original context so that all three
        # frequency bands — kept, blended, stretched — occur within 48 positions) and tied embeddings
        cfg = dict(n_layers=2, d_model=128, n_heads=4, n_kv_heads=2, head_dim=64, d_ff=256, vocab_size=256,
                   max_seq_len=64, rope_theta=500000.0, rms_eps=1e-5, rope_factor=8.0, rope_low_freq_factor=1.0,
                   rope_high_freq_factor=4.0, rope_original_max_pos=16)
    else:
        cfg = dict(n_layers=2, d_model=128, n_heads=4, n_kv_heads=1, head_dim=64, d_ff=192, vocab_size=320,
                   max_seq_len=64, rope_theta=1e6, rms_eps=1e-5)
    common = dict(hidden_size=cfg["d_model"], intermediate_size=cfg["d_ff"], num_hidden_layers=cfg["n_layers"],
                  num_attention_heads=cfg["n_heads"], num_key_value_heads=cfg["n_kv_heads"],
                  vocab_size=cfg["vocab_size"], max_position_embeddings=cfg["max_seq_len"],
                  rms_norm_eps=cfg["rms_eps"], rope_theta=cfg["rope_theta"], tie_word_embeddings=False,
                  head_dim=cfg["head_dim"],
                  attention_bias=False, hidden_act="silu")
    torch.manual_seed(seed)
    n_ids = 24
    if kind == "llama31rope":
        common.pop("rope_theta")
        common["tie_word_embeddings"] = True
        common["rope_parameters"] = dict(rope_type="llama3", rope_theta=cfg["rope_theta"], factor=cfg["rope_factor"],
                                         low_freq_factor=cfg["rope_low_freq_factor"], high_freq_factor=cfg["rope_high_freq_factor"],
                                         original_max_position_embeddings=cfg["rope_original_max_pos"])
        n_ids = 48
        model = LlamaForCausalLM(LlamaConfig(mlp_bias=False, **common))
    elif kind in ("llama", "llama3geom"):
        model = LlamaForCausalLM(LlamaConfig(mlp_bias=False, **common))
    else:
        model = MistralForCausalLM(MistralConfig(sliding_window=None, **common))
    model.eval()
    tensors = {}
    with torch.no_grad():
        for name, p in model.named_parameters():
            if "norm" in name:
                p.copy_(1.0 + 0.1 * torch.randn_like(p))
            else:
                p.copy_(0.08 * torch.randn_like(p))
            p.copy_(p.to(torch.bfloat16).to(torch.float32))       # bf16-representable weights
            tensors[name] = p.detach().clone()
    ids = torch.tensor([[(i * 7919 + 13) % cfg["vocab_size"] for i in range(n_ids)]])
    with torch.no_grad():
        out = model(ids, output_hidden_states=True)
    logits = out.logits[0].float().numpy()
    hidden_last = out.hidden_states[-1][0].float().numpy()  # after final norm in HF (norm applied)
    save = {"cfg_keys": np.array(list(cfg.keys())), "cfg_vals": np.array([float(v) for v in cfg.values()]),
            "ids": ids[0].numpy().astype(np.int32), "logits": logits.astype(np.float32),
            "hidden_last_normed": hidden_last.astype(np.float32)}

    def b16(t):
        return oc.np_bf16_from_f32(t.numpy().astype(np.float32).ravel())
    save["embed"] = b16(tensors["model.embed_tokens.weight"])
    save["lm_head"] = b16(tensors.get("lm_head.weight", tensors["model.embed_tokens.weight"]))   # tied: one matrix
    save["final_norm"] = b16(tensors["model.norm.weight"])
    for l in range(cfg["n_layers"]):
        pre = f"model.layers.{l}."
        m = {"ATTN_NORM": "input_layernorm.weight", "FFN_NORM": "post_attention_layernorm.weight",
             "WQ": "self_attn.q_proj.weight", "WK": "self_attn.k_proj.weight", "WV": "self_attn.v_proj.weight",
             "WO": "self_attn.o_proj.weight", "WGATE": "mlp.gate_proj.weight", "WUP": "mlp.up_proj.weight",
             "WDOWN": "mlp.down_proj.weight"}
        for k, v in m.items():
            save[f"L{l}.{k}"] = b16(tensors[pre + v])
    np.savez_compressed(path, **save)
    print("wrote", path, path.stat().st_size, "bytes")
    if ckpt_dir is not None:
        # the same model as an HF checkpoint directory (config.json + model.safetensors, bf16 — lossless, the weights
        # are bf16-representable): pins the engine's safetensors / config.json loader (csrc/weights_io.cpp) against
        # the real HF writer; tests/test_gpu_checkpoint.py loads it by path and compares with the logits above
        import shutil
        shutil.rmtree(ckpt_dir, ignore_errors=True)
        model.to(torch.bfloat16).save_pretrained(ckpt_dir, safe_serialization=True)
        for f in ckpt_dir.iterdir():
            if f.name not in ("config.json", "model.safetensors"):
                f.unlink()
        print("wrote", ckpt_dir, sorted(f.name for f in ckpt_dir.iterdir()))


def synth_kat(path: Path):
    """Known-answer vectors of the counter-based synthetic weight generator (numpy mirror)."""
    save = {}
    for seed, key, first, n in [(1234, 0, 0, 64), (1234, 4 + 16 * 31, 4096 * 4096 - 32, 32),
                                (0, 1, 525336576 - 16, 16), (2**63 + 12345, 11, 7, 33)]:
        save[f"int_{seed}_{key}_{first}_{n}"] = oc.np_synth_int(seed, key, first, n).astype(np.int32)
        save[f"bf16_{seed}_{key}_{first}_{n}"] = oc.np_synth_bf16(seed, key, first, n)
    np.savez_compressed(path, **save)
    print("wrote", path, path.stat().st_size, "bytes")


if __name__ == "__main__":
    only = sys.argv[1] if len(sys.argv) > 1 else None     # e.g. `make_golden.py llama3geom` regenerates one fixture
    if only in (None, "llama"):
        hf_fixture("llama", OUT / "hf_tiny_llama.npz", 0, ckpt_dir=OUT / "hf_tiny_llama_ckpt")
    if only in (None, "mistral"):
        hf_fixture("mistral", OUT / "hf_tiny_mistral.npz", 1)
    if only in (None, "llama3geom"):
        hf_fixture("llama3geom", OUT / "hf_tiny_llama3geom.npz", 2)
    if only in (None, "llama31rope"):
        hf_fixture("llama31rope", OUT / "hf_tiny_llama31rope.npz", 3, ckpt_dir=OUT / "hf_tiny_llama31rope_ckpt")
    if only in (None, "synth"):
        synth_kat(OUT / "synth_kat.npz")
