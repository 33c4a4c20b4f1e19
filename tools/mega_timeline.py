"""Per-CTA timeline of one layer of the persistent decode kernel (BLK_MEGA_TRACE=1): stage durations in us for a few CTAs."""
import os, sys
os.environ["BLK_MEGA_TRACE"] = "1"
sys.path.insert(0, '.')
import numpy as np
from bench import ensure_model
from blama_b200 import capi, gguf_synth
shape = sys.argv[1] if len(sys.argv) > 1 else "llama-3.1-8b-q4km"
ctx_len = int(sys.argv[2]) if len(sys.argv) > 2 else 512
layer = int(sys.argv[3]) if len(sys.argv) > 3 else 10
path = ensure_model(shape, 0, lambda: None)
m = capi.Model(path); c = capi.Ctx(m, 2048)
c.decode(gguf_synth.synth_prompt(shape, ctx_len, 1))
first = int(c.topk(1)["token"][0])
c.decode_loop(first, 4)
raw = c.debug_trace()
KIND = ["qkv", "wo", "gu", "down", "head"]
TAG = {2: "top", 3: "pro.tail", 4: "actregs", 5: "mac", 6: "epi"}
PRO = {30: "wait_x", 31: "sync+scale", 32: "wait_src", 34: "quant"}
ATT = {45: "actregs.rep", 40: "m.wait", 41: "m.dot", 42: "m.issue", 43: "m.bfly", 44: "m.end", 17: "pv.wait+max", 11: "scores", 13: "pv.stats", 14: "pv.pv", 15: "pv.partials", 16: "pv.combine", 20: "final"}
clk = 1.965e3
ctas = [0, 1, 2, 17, 18, 19, 73, 74, 100, 146, 147]
rows = {}
for cta in ctas:
    row = raw[cta]; row = row[row != 0]
    t = (row >> 8).astype(np.float64); tag = (row & 0xff).astype(int)
    # layer boundaries: tag (0<<5)|2 = qkv loop top
    starts = [i for i in range(len(tag)) if tag[i] == 2]
    i0, i1 = starts[layer], starts[layer + 1]
    seq = []
    for i in range(i0 + 1, i1 + 1):
        tg = int(tag[i])
        name = ATT.get(tg) or PRO.get(tg) or (KIND[tg >> 5] + "." + TAG.get(tg & 31, str(tg & 31)))
        seq.append((name, (t[i] - t[i - 1]) / clk))
    rows[cta] = seq
n = max(len(s) for s in rows.values())
print("stage".ljust(16) + "".join(f"cta{c:>4d} " for c in ctas))
for i in range(n):
    name = rows[ctas[0]][i][0] if i < len(rows[ctas[0]]) else "?"
    print(name.ljust(16) + "".join((f"{rows[c][i][1]:7.2f} " if i < len(rows[c]) and rows[c][i][0] == name else f"{'!' + rows[c][i][0][:5] if i < len(rows[c]) else '':>7s} ") for c in ctas))
print("layer total".ljust(16) + "".join(f"{sum(d for _, d in rows[c]):7.2f} " for c in ctas))
# per-layer matrix for one CTA
cta = int(os.environ.get("TL_CTA", "0"))
row = raw[cta]; row = row[row != 0]
t = (row >> 8).astype(np.float64); tag = (row & 0xff).astype(int)
starts = [i for i in range(len(tag)) if tag[i] == 2]
print(f"\nper-layer stage durations, CTA {cta}")
hdr = None
for l in range(len(starts) - 1):
    i0, i1 = starts[l], starts[l + 1]
    seq = []
    for i in range(i0 + 1, i1 + 1):
        tg = int(tag[i])
        name = ATT.get(tg) or PRO.get(tg) or (KIND[tg >> 5] + "." + TAG.get(tg & 31, str(tg & 31)))
        seq.append((name, (t[i] - t[i - 1]) / clk))
    if hdr is None:
        hdr = [n for n, _ in seq]; print("L   " + " ".join(f"{n[-7:]:>7s}" for n in hdr))
    print(f"{l:<3d} " + " ".join(f"{d:7.2f}" for _, d in seq) + f"  | {sum(d for _, d in seq):7.2f}")
