"""Hand-written safetensors writer for the checkpoint-loader tests (format: u64 LE header length | JSON | raw bytes)."""
import json
import struct

import numpy as np


def f32_from_bf16(bits: np.ndarray) -> np.ndarray:
    return (bits.astype(np.uint32) << 16).view(np.float32)


def write_safetensors(path, tensors: dict, dtype: str = "BF16"):
    """tensors: name -> (uint16 bf16 bits, shape).  dtype BF16 stores the bits; F32 / F16 store the (exact / rounded) value."""
    header, blobs, off = {}, [], 0
    for name, (bits, shape) in tensors.items():
        bits = np.ascontiguousarray(bits, dtype=np.uint16).reshape(-1)
        if dtype == "BF16":
            raw = bits.tobytes()
        elif dtype == "F32":
            raw = f32_from_bf16(bits).tobytes()
        else:
            raw = f32_from_bf16(bits).astype(np.float16).tobytes()
        header[name] = {"dtype": dtype, "shape": list(shape), "data_offsets": [off, off + len(raw)]}
        blobs.append(raw)
        off += len(raw)
    header["__metadata__"] = {"format": "pt"}
    hj = json.dumps(header, separators=(",", ":")).encode()
    hj += b" " * (-len(hj) % 8)
    with open(path, "wb") as f:
        f.write(struct.pack("<Q", len(hj)))
        f.write(hj)
        for b in blobs:
            f.write(b)


HF_NAMES = {"ATTN_NORM": "input_layernorm.weight", "FFN_NORM": "post_attention_layernorm.weight", "WQ": "self_attn.q_proj.weight",
            "WK": "self_attn.k_proj.weight", "WV": "self_attn.v_proj.weight", "WO": "self_attn.o_proj.weight",
            "WGATE": "mlp.gate_proj.weight", "WUP": "mlp.up_proj.weight", "WDOWN": "mlp.down_proj.weight"}


def hf_tensors_from_fixture(z, cfg) -> dict:
    """The npz golden fixture (tests/golden/make_golden.py) as HF tensor names -> (bf16 bits, shape)."""
    d, F, V = cfg["d_model"], cfg["d_ff"], cfg["vocab_size"]
    qd, kvd = cfg["n_heads"] * cfg["head_dim"], cfg["n_kv_heads"] * cfg["head_dim"]
    shp = {"ATTN_NORM": (d,), "FFN_NORM": (d,), "WQ": (qd, d), "WK": (kvd, d), "WV": (kvd, d), "WO": (d, qd), "WGATE": (F, d),
           "WUP": (F, d), "WDOWN": (d, F)}
    t = {"model.embed_tokens.weight": (z["embed"], (V, d)), "lm_head.weight": (z["lm_head"], (V, d)),
         "model.norm.weight": (z["final_norm"], (d,))}
    for l in range(cfg["n_layers"]):
        for k, nm in HF_NAMES.items():
            t[f"model.layers.{l}.{nm}"] = (z[f"L{l}.{k}"], shp[k])
    return t


FLOAT_CFG_KEYS = ("rope_theta", "rms_eps", "rope_factor", "rope_low_freq_factor", "rope_high_freq_factor")


def fixture_cfg(z) -> dict:
    """Model config stored in a golden npz (tests/golden/make_golden.py) as name / value arrays."""
    return {str(k): (float(v) if str(k) in FLOAT_CFG_KEYS else int(v)) for k, v in zip(z["cfg_keys"], z["cfg_vals"])}
