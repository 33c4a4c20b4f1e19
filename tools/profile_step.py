"""One short decode run for ncu launch lists / full captures: 8-token prompt, then N device-resident decode steps."""
import sys
sys.path.insert(0, '.')
from bench import ensure_model
from blama_b200 import capi, gguf_synth
shape = sys.argv[1] if len(sys.argv) > 1 else "llama-3.1-8b-q4km"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 3
path = ensure_model(shape, 0, lambda: None)
m = capi.Model(path); c = capi.Ctx(m, 4096)
npr = int(sys.argv[3]) if len(sys.argv) > 3 else 8
c.decode(gguf_synth.synth_prompt(shape, npr, 1))
first = int(c.topk(1)["token"][0])
c.timer_start(); c.decode_loop(first, n, wait=False); ms = c.timer_stop()
print(f"{n} steps: {ms/n*1e3:.1f} us/token")
if len(sys.argv) > 4:
    nv = int(sys.argv[4])
    c.clear(); c.decode(gguf_synth.synth_prompt(shape, 32, 1))
    toks = gguf_synth.synth_prompt(shape, nv, 2)
    print(c.profile_verify(toks))
    c.clear(); c.decode(gguf_synth.synth_prompt(shape, 32, 1))
    print(c.profile_verify(toks))
