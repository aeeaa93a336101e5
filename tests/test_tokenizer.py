"""Native tokenizer (csrc/tokenizer.cpp) vs the HF `tokenizers` library — golden vectors from
tests/golden/make_tokenizer_golden.py (four small tokenizers with the pipelines of Llama-2/Mistral/TinyLlama, of newer
Metaspace conversions, of Llama-3 and of GPT-2/Qwen-style byte-level BPE).  CPU only: the tokenizer is host logic behind the C-ABI."""
import json
from pathlib import Path

import pytest

from crowdllama_b200 import engine as eng

GOLD = Path(__file__).resolve().parent / "golden" / "tokenizers"
FAMILIES = ["spm_legacy", "metaspace", "llama3", "gpt2"]


@pytest.mark.parametrize("name", FAMILIES)
def test_encode_and_decode_match_hf_tokenizers(name):
    tok = eng.HfTokenizer(GOLD / f"{name}.tokenizer.json")
    cases = json.loads((GOLD / f"{name}.cases.json").read_text())["cases"]
    assert len(cases) >= 30
    bad = []
    for c in cases:
        ids = tok.encode(c["text"])
        if ids != c["ids"]:
            bad.append((c["text"], ids, c["ids"]))
            continue
        assert tok.decode(ids) == c["decoded"], c["text"]
    assert not bad, bad[:3]


@pytest.mark.parametrize("name", FAMILIES)
def test_live_against_the_library_on_fresh_strings(name):
    """Not only the committed cases: if `tokenizers` is importable, compare on strings generated here."""
    tokenizers = pytest.importorskip("tokenizers")
    ref = tokenizers.Tokenizer.from_file(str(GOLD / f"{name}.tokenizer.json"))
    tok = eng.HfTokenizer(GOLD / f"{name}.tokenizer.json")
    import random
    rng = random.Random(7)
    words = ["sky", "blue", "Rayleigh", "it's", "DON'T", "42", "2024", "3.14", "naïve", "Zürich", "日本", "язык", "🙂", "—", "(x+1)*2",
             "\n", "\n\n", "\t", " ", "  ", "...", "foo_bar", "CamelCase", "e=mc²", "½", "<s>", "</s>", "[INST]", "<|eot_id|>", "'ll", "'RE"]
    for _ in range(300):
        text = "".join(rng.choice(words) + rng.choice(["", " ", " ", "  ", "\n"]) for _ in range(rng.randint(1, 12)))
        want = ref.encode(text, add_special_tokens=False).ids
        got = tok.encode(text)
        assert got == want, repr(text)
        assert tok.decode(got) == ref.decode(want, skip_special_tokens=True), repr(text)


@pytest.mark.parametrize("name", FAMILIES)
def test_unicode_fuzz_against_the_library(name):
    """Strings drawn from ~30 Unicode blocks (letters of many scripts, combining marks, number forms, symbols, emoji,
    exotic spaces): the split scanners and the generated \\p{L} / \\p{N} tables must agree with the library's regex engine."""
    tokenizers = pytest.importorskip("tokenizers")
    ref = tokenizers.Tokenizer.from_file(str(GOLD / f"{name}.tokenizer.json"))
    tok = eng.HfTokenizer(GOLD / f"{name}.tokenizer.json")
    import random
    rng = random.Random(11)
    blocks = [(0x20, 0x7E), (0xA0, 0xFF), (0x100, 0x24F), (0x250, 0x2FF), (0x300, 0x36F), (0x370, 0x3FF), (0x400, 0x4FF), (0x530, 0x58F),
              (0x590, 0x5FF), (0x600, 0x6FF), (0x900, 0x97F), (0xE00, 0xE7F), (0x1100, 0x11FF), (0x1E00, 0x1EFF), (0x1F00, 0x1FFF),
              (0x2000, 0x206F), (0x2070, 0x209F), (0x20A0, 0x20CF), (0x2100, 0x214F), (0x2150, 0x218F), (0x2190, 0x21FF), (0x2200, 0x22FF),
              (0x2460, 0x24FF), (0x2600, 0x26FF), (0x3000, 0x303F), (0x3040, 0x309F), (0x30A0, 0x30FF), (0x4E00, 0x4FFF), (0xAC00, 0xACFF),
              (0xFF00, 0xFFEF), (0x1F300, 0x1F64F), (0x1D400, 0x1D7FF)]
    for _ in range(700):
        s = ""
        for _ in range(rng.randint(1, 8)):
            if rng.random() < 0.35:
                s += rng.choice([" ", "  ", "\n", "\t", "a", "Z", "7", ".", "'s", "'LL", "-"])
            else:
                lo, hi = rng.choice(blocks)
                s += chr(rng.randint(lo, hi))
        want = ref.encode(s, add_special_tokens=False).ids
        got = tok.encode(s)
        assert got == want, [f"U+{ord(c):04X}" for c in s]
        assert tok.decode(got) == ref.decode(want, skip_special_tokens=True), [f"U+{ord(c):04X}" for c in s]


def test_special_ids_chat_templates_and_errors(tmp_path):
    l3 = eng.HfTokenizer(GOLD / "llama3.tokenizer.json")
    ids = l3.encode("hi", add_bos=True, chat=True)
    names = {t: l3.encode(t)[0] for t in ["<|begin_of_text|>", "<|start_header_id|>", "<|end_header_id|>", "<|eot_id|>", "<|end_of_text|>"]}
    assert ids[0] == names["<|begin_of_text|>"] == l3.bos and l3.eos == names["<|end_of_text|>"]
    assert ids[1] == names["<|start_header_id|>"] and ids.count(names["<|eot_id|>"]) == 1
    assert ids.count(names["<|end_header_id|>"]) == 2 and l3.decode(ids[len(ids) - ids[::-1].index(names["<|end_header_id|>"]):]) == "\n\n"
    assert l3.decode(ids) == "user\n\nhiassistant\n\n"                  # special tokens are skipped on decode
    sp = eng.HfTokenizer(GOLD / "spm_legacy.tokenizer.json")
    assert sp.bos == sp.encode("<s>")[0] and sp.eos == sp.encode("</s>")[0]
    chat = sp.encode("do it", add_bos=True, chat=True)
    assert chat[0] == sp.bos and chat[1] == sp.encode("[INST]")[0] and chat[-1] == sp.encode("[/INST]")[0]
    assert eng.HfTokenizer(GOLD / "spm_legacy.tokenizer.json", chat_family="zephyr").encode("x", chat=True) != sp.encode("x", chat=True)
    g2 = eng.HfTokenizer(GOLD / "gpt2.tokenizer.json")                  # <|im_start|> present -> chatml framing
    c2 = g2.encode("x", chat=True)
    assert c2[0] == g2.encode("<|im_start|>")[0] and c2.count(g2.encode("<|im_end|>")[0]) == 1 and g2.eos == g2.encode("<|endoftext|>")[0]
    # byte fallback: characters outside the vocabulary become <0xXX> tokens and decode back to the same text
    assert sp.decode(sp.encode("₿ ⌘ 𝔘")) == "₿ ⌘ 𝔘"
    with pytest.raises(eng.EngineError):
        eng.HfTokenizer(tmp_path / "missing.json")
    (tmp_path / "bad.json").write_text('{"model": {"type": "WordPiece", "vocab": {}}}')
    with pytest.raises(eng.EngineError):
        eng.HfTokenizer(tmp_path / "bad.json")
    (tmp_path / "trunc.json").write_text('{"model": {"type": "BPE", "vocab": {"a": 0')
    with pytest.raises(eng.EngineError):
        eng.HfTokenizer(tmp_path / "trunc.json")
