"""Noise floor of the reference arithmetic at a BASELINE size: the oracle against itself (ORC_MODE_GGML vs ORC_MODE_GGML_ALT: the same
integer dot products, fp32 partial sums added in the opposite order), teacher-forced on the same tokens.  CPU only.
    python tools/noise_floor.py llama-3.1-8b-q4km [n_tokens]"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from blama_b200 import gguf_synth as gs  # noqa: E402
from oracle import pyoracle as po  # noqa: E402


def main():
    shape = sys.argv[1] if len(sys.argv) > 1 else "llama-3.1-8b-q4km"
    n = int(sys.argv[2]) if len(sys.argv) > 2 else 14
    path = f"/dev/shm/blama_b200_{shape}.gguf"
    if not os.path.exists(path):
        gs.write_gguf(path, shape)
    m = po.Model(path)
    a, b = po.Ctx(m, 64, po.MODE_GGML), po.Ctx(m, 64, po.MODE_GGML_ALT)
    toks = [int(t) for t in gs.synth_prompt(shape, n, 3)]
    for i, t in enumerate(toks):
        wa, wb = a.decode([t])[0], b.decode([t])[0]
        d = np.abs(wa - wb)
        top = np.argsort(-wa)[:10]
        print(f"{i:3d}  max {d.max():.4f}  rms {np.sqrt((d ** 2).mean()):.5f}  top-10 max {d[top].max():.4f}  logit std {wa.std():.3f}", flush=True)


if __name__ == "__main__":
    main()
