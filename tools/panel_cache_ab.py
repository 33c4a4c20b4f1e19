"""Resident bf16 panels (the model's panel cache, BLK_PANEL_CACHE_GB) against the streamed / fused GEMM forms: prompt prefill time
vs token count and the time of one batched decode step vs the number of sequences.
    python tools/panel_cache_ab.py [shape]"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bench import ensure_model  # noqa: E402
from blama_b200 import capi, gguf_synth  # noqa: E402

shape = sys.argv[1] if len(sys.argv) > 1 else "llama-3.1-8b-q4km"
path = ensure_model(shape, 0, lambda: None)
for name, gb in (("no cache (fused <= 256 tokens, streamed panels beyond)", "0"), ("resident panels", None)):
    if gb is None:
        os.environ.pop("BLK_PANEL_CACHE_GB", None)
    else:
        os.environ["BLK_PANEL_CACHE_GB"] = gb
    m = capi.Model(path)
    c = capi.Ctx(m, 2304)
    row = []
    for T in (32, 64, 128, 256, 512, 1024, 2048):
        toks = gguf_synth.synth_prompt(shape, T, 3)
        c.clear(); c.decode(toks); c.topk(1)                    # warm (allocations, attributes, the cache itself)
        best = 1e9
        for _ in range(3):
            c.clear(); c.flush_l2()
            t0 = time.perf_counter(); c.decode(toks); c.topk(1); best = min(best, time.perf_counter() - t0)
        row.append(f"T={T}: {best * 1e3:6.2f}")
    print(f"{name}\n  prefill  " + " | ".join(row) + " ms", flush=True)
    c.close()
    ctxs = [capi.Ctx(m, 256) for _ in range(64)]
    for i, cx in enumerate(ctxs):
        cx.decode(gguf_synth.synth_prompt(shape, 96, i))
    row = []
    for n in (1, 4, 8, 16, 32, 64):
        best = 1e9
        for step in range(4):
            toks = [int(t) for t in gguf_synth.synth_prompt(shape, n, 100 + step)]
            t0 = time.perf_counter(); capi.decode_batch(ctxs[0], ctxs[:n], toks, 40); dt = time.perf_counter() - t0
            if step: best = min(best, dt)
        row.append(f"n={n}: {best * 1e3:5.2f}")
    print("  batched decode step  " + " | ".join(row) + " ms", flush=True)
    for cx in ctxs:
        cx.close()
    m.close()
