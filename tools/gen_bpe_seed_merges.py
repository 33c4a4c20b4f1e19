"""Trains a small byte-level BPE on an embedded text and writes blama_b200/bpe_seed_merges.txt: the first merges of every synthetic
vocabulary (blama_b200/gguf_synth.py synth_vocab), so that ordinary English text exercises real multi-step merges.  The rest of a
synthetic vocabulary is filled with random (valid) merges.  Run once; the file is committed."""
import os
import re
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from blama_b200.gguf_synth import bytes_to_unicode  # noqa: E402

TEXT = """The first man to walk on the moon was Neil Armstrong, in July 1969. President George W. Bush said that it's not what you
think; they're going to the market, and we've seen that before. I'm sure you'll agree that he'd rather stay. France has a long history
of art, science and philosophy. The quick brown fox jumps over the lazy dog 1234567890 times. In the beginning there was nothing, and
then there was something: a model of language that predicts the next token from the previous ones. Transformers use attention to mix
information between positions; every layer has a normalisation, a projection into queries, keys and values, rotary embeddings, a
softmax over the scores, and a feed-forward network with a gated activation. The weather in the mountains changes quickly: rain in
the morning, sunshine at noon, snow in the evening. She said, "Don't worry about it -- it's fine!" and smiled. What is the capital of
the United States? Washington, D.C. is the capital. Numbers such as 3.14159, 2,718 and 42 appear often; dates like 2024-10-18 too.
    def main(argv):
        for i in range(10):
            print(i, argv[i % len(argv)])
        return 0
Verification re-computes the logits of every generated token and compares the top ten with the prover's claim. été, naïve, café,
über, señor, 東京, 北京, Москва, Ελλάδα. The server answers with text and token data. \tTabs\tand   spaces\n\nnew lines.
"""


def main(n_merges=1400):
    b2u = bytes_to_unicode()
    words = re.findall(r"'s|'t|'re|'ve|'m|'ll|'d| ?[^\W\d_]+| ?\d+| ?[^\s\w]+|\s+", TEXT)
    # repeat the text so that counts are well above 1
    seqs = {}
    for w in words:
        key = tuple(b2u[b] for b in w.encode("utf-8"))
        seqs[key] = seqs.get(key, 0) + 1
    seqs = [[list(k), v] for k, v in seqs.items()]
    merges = []
    for _ in range(n_merges):
        counts = {}
        for sym, cnt in seqs:
            for a, b in zip(sym, sym[1:]):
                counts[(a, b)] = counts.get((a, b), 0) + cnt
        if not counts:
            break
        best = max(counts.items(), key=lambda kv: (kv[1], kv[0]))[0]
        merges.append(best)
        a, b = best
        for entry in seqs:
            sym = entry[0]
            out, i = [], 0
            while i < len(sym):
                if i + 1 < len(sym) and sym[i] == a and sym[i + 1] == b:
                    out.append(a + b); i += 2
                else:
                    out.append(sym[i]); i += 1
            entry[0] = out
    path = os.path.join(ROOT, "blama_b200", "bpe_seed_merges.txt")
    with open(path, "w", encoding="utf-8") as f:
        for a, b in merges:
            f.write(f"{a} {b}\n")
    print(path, len(merges), "merges")


if __name__ == "__main__":
    main()
