"""A few batched decode steps (blk_decode_batch) for an ncu launch list:
    ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file out.csv python tools/profile_batch.py [shape] [n_seq] [ctx]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bench import ensure_model  # noqa: E402
from blama_b200 import capi, gguf_synth as gs  # noqa: E402

shape = sys.argv[1] if len(sys.argv) > 1 else "llama-3.1-8b-q4km"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 32
ctx_len = int(sys.argv[3]) if len(sys.argv) > 3 else 96
path = ensure_model(shape, 0, lambda: None)
m = capi.Model(path)
ctxs = [capi.Ctx(m, ctx_len + 64) for _ in range(n)]
for i, c in enumerate(ctxs):
    c.decode(gs.synth_prompt(shape, ctx_len, i))
import time
for step in range(4):
    t0 = time.perf_counter()
    capi.decode_batch(ctxs[0], ctxs, [int(t) for t in gs.synth_prompt(shape, n, 100 + step)], 40)
    print(f"step {step}: {(time.perf_counter() - t0) * 1e3:.2f} ms", flush=True)
