"""Stage-level timeline of the persistent decode kernel (BLK_MEGA_TRACE=1): where a token's time goes."""
import os, sys
os.environ["BLK_MEGA_TRACE"] = "1"
sys.path.insert(0, '.')
import numpy as np
from bench import ensure_model
from blama_b200 import capi, gguf_synth
shape = sys.argv[1] if len(sys.argv) > 1 else "llama-3.1-8b-q4km"
ctx_len = int(sys.argv[2]) if len(sys.argv) > 2 else 512
path = ensure_model(shape, 0, lambda: None)
m = capi.Model(path); c = capi.Ctx(m, 2048)
c.decode(gguf_synth.synth_prompt(shape, ctx_len, 1))
first = int(c.topk(1)["token"][0])
c.decode_loop(first, 4)
tr = c.debug_trace().astype(np.float64)
names = ["embed+sync"]
per_layer = ["qkv.prologue", "qkv.mac", "qkv.epilogue", "qkv.sync", "attn.scores", "attn.sync1", "attn.pv", "attn.sync2",
             "wo.prologue", "wo.mac", "wo.epilogue", "wo.sync", "gu.prologue", "gu.mac", "gu.epilogue", "gu.sync",
             "down.prologue", "down.mac", "down.epilogue", "down.sync"]
n_layer = m.n_layer
names += per_layer * n_layer + ["head.prologue", "head.mac", "head.epilogue", "head.sync"]
d = np.diff(tr[:, : len(names) + 1], axis=1)          # [cta][event]
agg = {}
for i, nme in enumerate(names):
    agg.setdefault(nme, []).append(d[:, i])
clk = 1.965e3   # cycles per us at the max SM clock
tot = 0.0
print(f"{'stage':16s} {'n':>4s} {'mean us':>9s} {'max-cta us':>10s} {'sum us':>9s}")
for nme, lst in agg.items():
    a = np.stack(lst, 1)                              # [cta][occurrence]
    mean = a.mean() / clk; mx = a.max(0).mean() / clk; s = a.mean(0).sum() / clk
    tot += s
    print(f"{nme:16s} {a.shape[1]:4d} {mean:9.2f} {mx:10.2f} {s:9.1f}")
print(f"total {tot:.1f} us")
