"""Cross-CTA timeline of the persistent decode kernel (BLK_MEGA_TRACE=2: stamps are %globaltimer, common to all SMs).
For one layer: when (us after the layer's first event) each stage ENDS, as min / mean / max over the CTAs -> the critical path."""
import os, sys
os.environ["BLK_MEGA_TRACE"] = "2"
sys.path.insert(0, '.')
import numpy as np
from bench import ensure_model
from blama_b200 import capi, gguf_synth
shape = sys.argv[1] if len(sys.argv) > 1 else "llama-3.1-8b-q4km"
ctx_len = int(sys.argv[2]) if len(sys.argv) > 2 else 512
layers = [int(x) for x in (sys.argv[3] if len(sys.argv) > 3 else "10,11").split(",")]
path = ensure_model(shape, 0, lambda: None)
m = capi.Model(path); c = capi.Ctx(m, 2048)
c.decode(gguf_synth.synth_prompt(shape, ctx_len, 1))
first = int(c.topk(1)["token"][0])
c.decode_loop(first, 4)
raw = c.debug_trace()
KIND = ["qkv", "wo", "gu", "down", "head"]
TAG = {2: "top", 3: "prologue done", 4: "act regs", 5: "mac", 6: "epilogue(published)"}
PRO = {30: "pro: x arrived+sumsq", 31: "pro: scale", 32: "pro: src blk0 arrived", 34: "pro: quantised"}
ATT = {17: "pv.wait+max", 11: "attn: scores published", 13: "attn: stats", 14: "attn: pv", 15: "attn: partials published", 16: "attn: combine published", 20: "final"}
per = {}
for cta in range(raw.shape[0]):
    row = raw[cta]; row = row[row != 0]
    t = (row >> 8).astype(np.float64) / 1e3; tag = (row & 0xff).astype(int)
    starts = [i for i in range(len(tag)) if tag[i] == 2]
    per[cta] = (t, tag, starts)
res = np.diff(np.unique(np.concatenate([per[c][0] for c in per])))
print(f"globaltimer resolution ~{res[res > 0].min() * 1e3:.0f} ns")
for layer in layers:
    t0 = min(per[c][0][per[c][2][layer]] for c in per)
    n_ev = per[0][2][layer + 1] - per[0][2][layer]
    print(f"\nlayer {layer}: stage end times, us after the first CTA enters the layer   (min / mean / max over CTAs, n)")
    # group events by (ordinal among same-length CTAs)
    groups = {}
    for cta, (t, tag, starts) in per.items():
        i0, i1 = starts[layer], starts[layer + 1]
        groups.setdefault(i1 - i0, []).append((cta, t[i0:i1 + 1] - t0, tag[i0:i1 + 1]))
    for n, lst in sorted(groups.items(), reverse=True):
        print(f"-- {len(lst)} CTAs with {n} events (e.g. CTA {lst[0][0]})")
        T = np.array([x[1] for x in lst]); tg = lst[0][2]
        for i in range(T.shape[1]):
            g = int(tg[i])
            name = ATT.get(g) or PRO.get(g) or (KIND[g >> 5] + ": " + TAG.get(g & 31, str(g & 31)))
            print(f"  {name:30s} {T[:, i].min():7.2f} {T[:, i].mean():7.2f} {T[:, i].max():7.2f}   argmax CTA {lst[int(T[:, i].argmax())][0]}")
