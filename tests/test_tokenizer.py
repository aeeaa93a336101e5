"""Native tokenizer (csrc/tokenizer.cpp) vs the HF `tokenizers` library — golden vectors from
tests/golden/make_tokenizer_golden.py (three small tokenizers with the pipelines of Llama-2/Mistral/TinyLlama, of newer
Metaspace conversions and of Llama-3).  CPU only: the tokenizer is host logic behind the C-ABI."""
import json
from pathlib import Path

import pytest

from crowdllama_b200 import engine as eng

GOLD = Path(__file__).resolve().parent / "golden" / "tokenizers"
FAMILIES = ["spm_legacy", "metaspace", "llama3"]


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
