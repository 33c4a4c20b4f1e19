"""Byte-level BPE tokenizer / detokenizer (reference inference/code/llama/Vocab.cpp:37-72 -> llama_tokenize / llama_token_to_piece).
Product: blama_b200/host/llama/Tokenizer.cpp through a vocabulary-only Model (no device).  Checked against
  * oracle/pytokenizer.py (llama.cpp's algorithm with the published regexes run by the `regex` engine), and
  * Hugging Face `tokenizers` built from the same vocabulary, merges and pattern (independent implementation).
The reference's own goldens (t-integration.cpp:41-42) use GPT-2's vocabulary file, which is not in the container."""
import os

import numpy as np
import pytest

from blama_b200 import gguf_synth as gs
from blama_b200 import host_api as H
from oracle import pytokenizer as PT

SHAPES = ["tiny-llama-q4km", "small-qwen2-q8", "tiny-llama-q8:gpt-2"]      # name[:tokenizer.ggml.pre override]

TEXTS = [
    "The first man to", "hello world", "France has a long history of", "President George W.",
    "I'm sure you'll agree that he'd rather stay; they're here, we've seen it. DON'T SHOUT, IT'S RUDE, WE'LL SEE",
    "numbers 1 12 123 1234 12345 1234567890 3.14159 2,718 x86_64 2024-10-18",
    "  leading spaces", "trailing spaces   ", "a  b   c    d", "tabs\tand\ttabs\t\t", "line\nbreaks\n\nand\r\nmore \n \n x",
    " \n", "\n\n\n", "   ", " ", "", "x", "!!!", "wait... what?!\n", "a.b,c;d:e", " ,. ;", "(parenthesised) [bracketed] {braced}",
    "été naïve café über señor", "東京 北京 日本語のテキスト", "Москва и Ελλάδα", "emoji \U0001F600\U0001F680 mix", "mixed123abc456",
    "'s 't 're 've 'm 'll 'd 'S 'T 'RE 'x ''", "it's'll", "under_score __dunder__", "tab\t\n\tnewline", "a b c　d",
    "def main(argv):\n    for i in range(10):\n        print(i)\n", "<|eot_id|>", "text<|begin_of_text|>more<|eot_id|>", "<|im_end|> after",
    "<|not_a_token|>", "<|eot_id", "१२३ ٤٥٦ Ⅻ ½", "ａｂｃ１２３", "a\x00b", "\x7f\x01",
]


@pytest.fixture(scope="module", params=SHAPES)
def setup(request, tmp_path_factory):
    name, _, pre_override = request.param.partition(":")
    path = str(tmp_path_factory.mktemp("vocab") / (name + ".gguf"))
    gs.write_gguf(path, name, pre=pre_override or None)
    shape = gs.SHAPES[name]
    tokens, types, merges = gs.synth_vocab(shape)
    sp = gs.special_tokens(shape)
    pre = pre_override or ("llama-bpe" if shape.arch == "llama" else "qwen2")
    ora = PT.Tokenizer(tokens, types, merges, pre, sp["bos"], sp["eos"], add_bos=shape.arch == "llama")
    model = H.Model(path, vocab_only=True)
    yield name, model, ora, (tokens, types, merges, pre, sp)
    model.close()


def random_texts(rng, n):
    alphabet = list("abcdefghijklmnopqrstuvwxyzABCDEFGHIJ   \n\t\r'.,;:!?-_()0123456789") + ["é", "ü", "東", "京", "я", "λ", " ", " ", "’", "—", "½", "\U0001F600", "'s", "'LL", "the ", " of", "  "]
    out = []
    for _ in range(n):
        k = int(rng.integers(1, 60))
        out.append("".join(alphabet[int(i)] for i in rng.integers(0, len(alphabet), k)))
    return out


def test_vocab_only_model_needs_no_device(setup):
    _, model, _, _ = setup
    assert model.train_ctx() == 0                  # reference "vocab only" test: no weights, no training context
    with pytest.raises(H.HostError):
        H.Instance(model, 64)                      # and no context can be created on it


def test_tokenize_matches_the_oracle(setup):
    name, model, ora, _ = setup
    rng = np.random.default_rng(11)
    for text in TEXTS + random_texts(rng, 400):
        for add_special, parse_special in ((True, True), (False, True), (False, False)):
            want = ora.tokenize(text, add_special, parse_special)
            got = model.tokenize(text, add_special, parse_special).tolist()
            assert got == want, (name, text, add_special, parse_special)


def test_round_trip_and_pieces(setup):
    name, model, ora, (tokens, types, _, _, sp) = setup
    rng = np.random.default_rng(3)
    for text in TEXTS + random_texts(rng, 200):
        ids = model.tokenize(text, add_special=False, parse_special=False)
        raw = b"".join(model.token_to_bytes(int(t)) for t in ids)
        assert raw == text.encode("utf-8"), (name, text)             # byte-level BPE is lossless
    for t in list(range(0, 300)) + [len(tokens) - 1, sp["bos"], sp["eos"], sp["eot"]] + rng.integers(0, len(tokens), 200).tolist():
        assert model.token_to_bytes(int(t), True) == ora.token_to_piece(int(t), True)
        assert model.token_to_bytes(int(t), False) == ora.token_to_piece(int(t), False)
    assert model.token_to_bytes(sp["eot"], special=False) == b""      # control tokens print only when asked to
    assert model.is_eog(sp["eot"]) and model.is_eog(sp["eos"]) and not model.is_eog(5)


def test_special_tokens_and_bos(setup):
    name, model, ora, (tokens, types, _, _, sp) = setup
    eot = tokens[sp["eot"]]
    ids = model.tokenize("a" + eot + "b", add_special=False, parse_special=True).tolist()
    assert sp["eot"] in ids and len(ids) == 3
    spelled = model.tokenize("a" + eot + "b", add_special=False, parse_special=False).tolist()
    assert sp["eot"] not in spelled and len(spelled) > 3
    with_bos = model.tokenize("hello", add_special=True).tolist()
    if name.startswith("tiny-llama"):
        assert with_bos[0] == sp["bos"] and with_bos[1:] == model.tokenize("hello", add_special=False).tolist()
    else:
        assert with_bos == model.tokenize("hello", add_special=False).tolist()      # qwen2 files do not add BOS


def test_against_huggingface_tokenizers(setup):
    tk = pytest.importorskip("tokenizers")
    name, model, _, (tokens, types, merges, pre, sp) = setup
    base = min(sp.values())
    vocab = {t: i for i, t in enumerate(tokens[:base])}
    pairs = [tuple(m.split(" ")) for m in merges]
    kind = PT.pre_type(pre)
    bpe = tk.models.BPE(vocab=vocab, merges=pairs, ignore_merges=(kind == "llama3"))
    hf = tk.Tokenizer(bpe)
    hf.pre_tokenizer = tk.pre_tokenizers.Sequence([
        tk.pre_tokenizers.Split(tk.Regex(PT.PATTERNS[kind]), behavior="isolated"),
        tk.pre_tokenizers.ByteLevel(add_prefix_space=False, use_regex=False)])
    rng = np.random.default_rng(23)
    checked = 0
    for text in TEXTS + random_texts(rng, 300):
        if "<|" in text or "\x00" in text:
            continue
        want = hf.encode(text, add_special_tokens=False).ids
        got = model.tokenize(text, add_special=False, parse_special=False).tolist()
        assert got == want, (name, text)
        checked += 1
    assert checked > 300
