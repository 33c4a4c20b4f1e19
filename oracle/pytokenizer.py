"""TEST INFRASTRUCTURE ONLY -- CPU restatement of llama.cpp's byte-level BPE tokenizer (ggml-org/llama.cpp b5187,
src/llama-vocab.cpp: llm_tokenizer_bpe + tokenizer_st_partition + llama_vocab::token_to_piece; un-vendored, restated from its
published algorithm) behind the reference's Vocab::tokenize / tokenToString (inference/code/llama/Vocab.cpp:37-72).

Unlike the product (blama_b200/host/llama/Tokenizer.cpp, a hand-written matcher) this oracle runs the pre-tokenizer patterns through the
`regex` engine exactly as written in llama.cpp's table, so the two are independent statements of the same behaviour.
Pin status: cross-checked against Hugging Face `tokenizers` (the implementation llama.cpp's tokenizer tests are generated from) in
tests/test_tokenizer.py; the reference's own goldens (t-integration.cpp:41-42, GPT-2 vocabulary) need a file that is not in the container.
Nothing under blama_b200/ may import this module."""
from __future__ import annotations

from typing import Dict, List, Sequence, Tuple

import regex

PATTERNS = {
    "llama3": r"(?i:'s|'t|'re|'ve|'m|'ll|'d)|[^\r\n\p{L}\p{N}]?\p{L}+|\p{N}{1,3}| ?[^\s\p{L}\p{N}]+[\r\n]*|\s*[\r\n]+|\s+(?!\S)|\s+",
    "qwen2": r"(?:'[sS]|'[tT]|'[rR][eE]|'[vV][eE]|'[mM]|'[lL][lL]|'[dD])|[^\r\n\p{L}\p{N}]?\p{L}+|\p{N}| ?[^\s\p{L}\p{N}]+[\r\n]*|\s*[\r\n]+|\s+(?!\S)|\s+",
    "gpt2": r"'s|'t|'re|'ve|'m|'ll|'d| ?\p{L}+| ?\p{N}+| ?[^\s\p{L}\p{N}]+|\s+(?!\S)|\s+",
}


def pre_type(name: str) -> str:
    if name in ("llama-bpe", "llama3", "llama-v3"):
        return "llama3"
    return "qwen2" if name == "qwen2" else "gpt2"


def bytes_to_unicode() -> List[str]:
    keep = list(range(33, 127)) + list(range(161, 173)) + list(range(174, 256))
    out, extra = [""] * 256, 0
    for b in range(256):
        if b in keep:
            out[b] = chr(b)
        else:
            out[b] = chr(256 + extra); extra += 1
    return out


class Tokenizer:
    def __init__(self, tokens: Sequence[str], types: Sequence[int], merges: Sequence[str], pre: str, bos: int, eos: int,
                 add_bos: bool, add_eos: bool = False):
        self.tokens, self.types = list(tokens), list(types)
        self.by_text: Dict[str, int] = {}
        for i, t in enumerate(self.tokens):
            self.by_text.setdefault(t, i)
        self.rank: Dict[Tuple[str, str], int] = {}
        for r, m in enumerate(merges):
            a, b = m.split(" ", 1) if " " not in m[1:] else (m[: m.index(" ", 1)], m[m.index(" ", 1) + 1:])
            self.rank.setdefault((a, b), r)
        self.kind = pre_type(pre)
        self.rx = regex.compile(PATTERNS[self.kind])
        self.ignore_merges = self.kind == "llama3"
        self.bos, self.eos, self.add_bos, self.add_eos = bos, eos, add_bos, add_eos
        self.b2u = bytes_to_unicode()
        self.u2b = {c: b for b, c in enumerate(self.b2u)}
        # cache_special_tokens: CONTROL | USER_DEFINED | UNKNOWN, longest text first
        self.specials = sorted((i for i, t in enumerate(self.types) if t in (2, 3, 4) and self.tokens[i]), key=lambda i: -len(self.tokens[i]))

    def _bpe(self, word: str, out: List[int]) -> None:
        if self.ignore_merges and word in self.by_text:
            out.append(self.by_text[word]); return
        sym = list(word)
        while len(sym) > 1:
            best, where = None, -1
            for i in range(len(sym) - 1):
                r = self.rank.get((sym[i], sym[i + 1]))
                if r is not None and (best is None or r < best):
                    best, where = r, i
            if best is None:
                break
            sym[where: where + 2] = [sym[where] + sym[where + 1]]
        for s in sym:
            if s in self.by_text:
                out.append(self.by_text[s])
            else:
                for ch in s.encode("utf-8"):
                    t = self.by_text.get(bytes([ch]).decode("latin-1"))
                    if t is not None and ch < 0x80:
                        out.append(t)

    def tokenize(self, text: str, add_special: bool, parse_special: bool) -> List[int]:
        out: List[int] = []
        if add_special and self.add_bos and self.bos >= 0:
            out.append(self.bos)
        frags: List[Tuple[bool, object]] = [(False, text)]
        for sp in self.specials:
            if not parse_special and self.types[sp] in (2, 3):
                continue
            needle = self.tokens[sp]
            nxt: List[Tuple[bool, object]] = []
            for special, val in frags:
                if special:
                    nxt.append((special, val)); continue
                parts = val.split(needle)
                for k, part in enumerate(parts):
                    if k:
                        nxt.append((True, sp))
                    if part:
                        nxt.append((False, part))
            frags = nxt
        for special, val in frags:
            if special:
                out.append(val); continue
            for w in self.rx.findall(val):
                self._bpe("".join(self.b2u[b] for b in w.encode("utf-8")), out)
        if add_special and self.add_eos and self.eos >= 0:
            out.append(self.eos)
        return out

    def split(self, text: str) -> List[str]:
        return self.rx.findall(text)

    def token_to_piece(self, tok: int, special: bool = True) -> bytes:
        t, text = self.types[tok], self.tokens[tok]
        if not special and t in (2, 3, 5):
            return b""
        if t in (2, 3, 4, 5):
            return text.encode("utf-8")
        if t != 1:
            return b""
        out = bytearray()
        for ch in text:
            if ch in self.u2b:
                out.append(self.u2b[ch])
            else:
                out += ch.encode("utf-8")
        return bytes(out)
