"""Decode tok/s of the persistent kernel against the context length (device-resident greedy loop, CUDA events, 64 tokens per point).
    python tools/decode_ctx_sweep.py [shape] [ctx ...]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bench import ensure_model  # noqa: E402
from blama_b200 import capi, gguf_synth as gs  # noqa: E402

args = sys.argv[1:]
shape = args[0] if args and not args[0].isdigit() else "llama-3.1-8b-q4km"
ctxs = [int(a) for a in args if a.isdigit()] or [64, 512, 1024, 1152, 1280, 2048, 4096, 8192]
path = ensure_model(shape, 0, lambda: None)
m = capi.Model(path)
c = capi.Ctx(m, max(ctxs) + 128)
for n in ctxs:
    c.clear(); c.decode(gs.synth_prompt(shape, n, 1))
    first = int(c.topk(1)["token"][0])
    c.decode_loop(first, 4)                                  # warm
    c.timer_start(); c.decode_loop(first, 64, wait=False); ms = c.timer_stop()
    print(f"{shape} context {n:5d}: {64 / ms * 1e3:7.1f} tok/s ({ms / 64 * 1e3:7.1f} us/token)", flush=True)
