"""Per-kernel HBM throughput of the decode mat-vecs + device-resident decode loop rate (quick iteration aid)."""
import sys, time
sys.path.insert(0, '.')
import numpy as np
from bench import ensure_model
from blama_b200 import capi, gguf_synth
shape = sys.argv[1] if len(sys.argv) > 1 else "llama-3.1-8b-q4km"
path = ensure_model(shape, 0, lambda: None)
m = capi.Model(path); c = capi.Ctx(m, 1024)
print("persistent decode kernel:", c.persistent_decode)
names = ["gate_up", "down", "qkv", "wo", "lm_head"]
for w, n in enumerate(names):
    ms, b = c.bench_kernel(w, 64 if w < 4 else 16)
    print(f"{n:8s} {ms*1e3:8.2f} us  {b/1e6:8.2f} MB  {b/ms/1e6:8.1f} GB/s  ({b/ms/1e6/6553*100:.1f}% of 6553)")
prompt = gguf_synth.synth_prompt(shape, 512, 1)
c.decode(prompt)
for rep in range(3):
    first = int(c.topk(1)["token"][0])
    c.timer_start(); c.decode_loop(first, 128, wait=False); ms = c.timer_stop()
    print(f"decode loop: {128/ms*1e3:.1f} tok/s  ({ms/128*1e3:.1f} us/token) at ctx ~{c.n_past}")
print(c.profile_step(first))
print(c.profile_step(first))
