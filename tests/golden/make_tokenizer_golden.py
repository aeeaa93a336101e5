"""Golden vectors for the native tokenizer (csrc/tokenizer.cpp), generated with the HF `tokenizers` library (0.22).

No real vocabulary files exist offline, so three SMALL tokenizers are trained here with exactly the pipeline
components the target model families ship in their tokenizer.json:
  spm_legacy   Llama-2 / Mistral-7B-v0.1 / TinyLlama: normalizer Prepend("▁") + Replace(" ", "▁"), no pre-tokenizer,
               BPE with byte_fallback, decoder Replace/ByteFallback/Fuse/Strip, "<s>" / "</s>" / "<unk>"
  metaspace    newer SentencePiece conversions: Metaspace pre-tokenizer (prepend_scheme "first", split False)
  llama3       Llama-3: Split(<tiktoken cl100k-style regex>) + ByteLevel, BPE with ignore_merges, ByteLevel decoder,
               "<|begin_of_text|>" ... "<|eot_id|>" added tokens
  gpt2         GPT-2 / Qwen-style: ByteLevel pre-tokenizer with its built-in regex, "<|endoftext|>", "<|im_start|>"
and every case string is encoded / decoded by the library.  The C++ implementation must reproduce ids and text.

    python tests/golden/make_tokenizer_golden.py      # rewrites tests/golden/tokenizers/*.json
"""
import json
from pathlib import Path

from tokenizers import AddedToken, Regex, Tokenizer, decoders, models, normalizers, pre_tokenizers, trainers

OUT = Path(__file__).resolve().parent / "tokenizers"
CORPUS = [
    "Why is the sky blue during the day and red at sunset? Because of Rayleigh scattering.",
    "The quick brown fox jumps over the lazy dog. THE QUICK BROWN FOX JUMPS OVER THE LAZY DOG!",
    "def fibonacci(n):\n    if n < 2:\n        return n\n    return fibonacci(n - 1) + fibonacci(n - 2)\n",
    "It's 12:45 on 2024-03-17; I'll pay $1,234.56 — that's 100% fine, isn't it? We've done it, they're here, I'd say, he's in.",
    "Größe, naïve café, señor, São Paulo, Zürich, œuvre, Ærø; Ελληνικά; Русский язык; 日本語のテキスト; 한국어; العربية",
    "tabs\tand\nnewlines\r\n\r\nand   multiple    spaces   at the end   ",
    "{\"model\": \"llama3:8b\", \"messages\": [{\"role\": \"user\", \"content\": \"hi\"}], \"stream\": false}",
    "x = [i ** 2 for i in range(10)]  # squares 0 1 4 9 16 25 36 49 64 81 1234567890",
    "Emoji: 🙂🚀 and symbols ©®™ ±×÷ √∞ ≈≠ ≤≥ ←→ and a snowman ☃.",
] * 3
CASES = [
    "", " ", "  ", "a", "Hello", "Hello world", " Hello  world ", "Hello, world! How are you?", "why is the sky blue? ☃",
    "It's they're WE'VE i'll I'D he's can't", "1 12 123 1234 12345 3.14159 1,000,000", "  leading and trailing  ",
    "line one\nline two\r\nline three\n\n\nend", "tab\tseparated\tvalues", "def f(x):\n    return x + 1\n",
    "Größe naïve café señor", "日本語のテキスト", "Русский язык", "emoji 🙂🚀 ok", "unseen chars: ₿ ⌘ 𝔘 ", "a" * 70,
    "mixed123abc 4five6", "....!!!???", "   \n   \n", "end with space ", "\n", "x" + " " * 9 + "y",
    "<s> literal and </s> tokens <unk>", "<|begin_of_text|>hi<|eot_id|>", "[INST] do it [/INST]",
    "<|start_header_id|>user<|end_header_id|>\n\nq<|eot_id|><|start_header_id|>assistant<|end_header_id|>\n\n",
]
LLAMA3_SPLIT = (r"(?i:'s|'t|'re|'ve|'m|'ll|'d)|[^\r\n\p{L}\p{N}]?\p{L}+|\p{N}{1,3}| ?[^\s\p{L}\p{N}]+[\r\n]*|\s*[\r\n]+|\s+(?!\S)|\s+")


def spm(metaspace: bool) -> Tokenizer:
    tok = Tokenizer(models.BPE(unk_token="<unk>", byte_fallback=True, fuse_unk=True))
    if metaspace:
        tok.pre_tokenizer = pre_tokenizers.Metaspace(replacement="▁", prepend_scheme="first", split=False)
        tok.decoder = decoders.Sequence([decoders.Replace("▁", " "), decoders.ByteFallback(), decoders.Fuse(), decoders.Strip(" ", 1, 0)])
    else:
        tok.normalizer = normalizers.Sequence([normalizers.Prepend("▁"), normalizers.Replace(" ", "▁")])
        tok.decoder = decoders.Sequence([decoders.Replace("▁", " "), decoders.ByteFallback(), decoders.Fuse(), decoders.Strip(" ", 1, 0)])
    specials = ["<unk>", "<s>", "</s>"] + [f"<0x{b:02X}>" for b in range(256)]
    tr = trainers.BpeTrainer(vocab_size=700, special_tokens=specials, show_progress=False)
    tok.train_from_iterator(CORPUS, tr)
    # the byte tokens are ordinary vocabulary in real files, only <unk>/<s>/</s> (+ chat markers) are "added"
    js = json.loads(tok.to_str())
    js["added_tokens"] = [t for t in js["added_tokens"] if not t["content"].startswith("<0x")]
    tok = Tokenizer.from_str(json.dumps(js))
    tok.add_special_tokens([AddedToken("[INST]", special=True, normalized=False), AddedToken("[/INST]", special=True, normalized=False)])
    return tok


def llama3() -> Tokenizer:
    tok = Tokenizer(models.BPE(ignore_merges=True))
    tok.pre_tokenizer = pre_tokenizers.Sequence([pre_tokenizers.Split(Regex(LLAMA3_SPLIT), behavior="isolated", invert=False),
                                                 pre_tokenizers.ByteLevel(add_prefix_space=False, trim_offsets=True, use_regex=False)])
    tok.decoder = decoders.ByteLevel(add_prefix_space=True, trim_offsets=True, use_regex=True)
    tr = trainers.BpeTrainer(vocab_size=900, special_tokens=[], initial_alphabet=pre_tokenizers.ByteLevel.alphabet(), show_progress=False)
    tok.train_from_iterator(CORPUS, tr)
    tok.add_special_tokens([AddedToken(t, special=True, normalized=False) for t in
                            ["<|begin_of_text|>", "<|end_of_text|>", "<|start_header_id|>", "<|end_header_id|>", "<|eot_id|>"]])
    return tok


def gpt2() -> Tokenizer:
    tok = Tokenizer(models.BPE())
    tok.pre_tokenizer = pre_tokenizers.ByteLevel(add_prefix_space=False, trim_offsets=True, use_regex=True)
    tok.decoder = decoders.ByteLevel(add_prefix_space=True, trim_offsets=True, use_regex=True)
    tr = trainers.BpeTrainer(vocab_size=900, special_tokens=[], initial_alphabet=pre_tokenizers.ByteLevel.alphabet(), show_progress=False)
    tok.train_from_iterator(CORPUS, tr)
    tok.add_special_tokens([AddedToken(t, special=True, normalized=False) for t in ["<|endoftext|>", "<|im_start|>", "<|im_end|>"]])
    return tok


def main():
    OUT.mkdir(parents=True, exist_ok=True)
    for name, tok in (("spm_legacy", spm(False)), ("metaspace", spm(True)), ("llama3", llama3()), ("gpt2", gpt2())):
        tok.save(str(OUT / f"{name}.tokenizer.json"))
        cases = []
        for text in CASES:
            enc = tok.encode(text, add_special_tokens=False)
            cases.append(dict(text=text, ids=enc.ids, decoded=tok.decode(enc.ids, skip_special_tokens=True)))
        (OUT / f"{name}.cases.json").write_text(json.dumps(dict(library=f"tokenizers {__import__('tokenizers').__version__}", cases=cases),
                                                           ensure_ascii=False, indent=0))
        print(name, tok.get_vocab_size(), "tokens;", sum(len(c["ids"]) for c in cases), "ids over", len(cases), "cases")


if __name__ == "__main__":
    main()
