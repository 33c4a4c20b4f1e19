"""Stage-level timeline of the persistent decode kernel (BLK_MEGA_TRACE=1): where a token's time goes.
Every CTA's thread 0 records (SM clock, tag) at stage boundaries; intervals are attributed to the tag that ENDS them."""
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
raw = c.debug_trace()
KIND = ["qkv", "wo", "gu", "down", "head"]
TAG = {2: "(loop top)", 3: "prologue.tail", 4: "act regs", 5: "mac", 6: "epilogue"}
PRO = {30: "pro.wait x+sumsq", 31: "pro.sync+scale", 32: "pro.wait src blk0", 34: "pro.quantise"}
ATT = {40: "wo.mac.wait", 41: "wo.mac.dot", 42: "wo.mac.issue", 43: "wo.mac.butterfly", 44: "wo.mac.end", 17: "attn.pv.wait+max", 11: "attn.scores(+wait q)", 13: "attn.pv.stats", 14: "attn.pv.pv", 15: "attn.pv.partials", 16: "attn.pv.combine", 20: "final wait"}
clk = 1.965e3
agg = {}
for cta in range(raw.shape[0]):
    row = raw[cta]; row = row[row != 0]
    t = (row >> 8).astype(np.float64); tag = (row & 0xff).astype(int)
    kind = 0
    for i in range(1, len(t)):
        tg = int(tag[i])
        if tg in ATT: name = ATT[tg]
        elif tg in PRO: name = PRO[tg]
        else: name = KIND[tg >> 5] + "." + TAG.get(tg & 31, str(tg & 31))
        agg.setdefault(name, {}).setdefault(cta, []).append((t[i] - t[i - 1]) / clk)
tot = 0.0
print(f"{'stage':28s} {'n':>4s} {'mean us':>9s} {'max-cta us':>10s} {'sum us':>9s}")
for name, per in agg.items():
    a = np.array([per[k] for k in sorted(per) if len(per[k]) == len(per[0])])
    mean = a.mean(); mx = a.max(0).mean(); s = a.mean(0).sum(); tot += s
    print(f"{name:28s} {a.shape[1]:4d} {mean:9.2f} {mx:10.2f} {s:9.1f}")
print(f"total {tot:.1f} us")
