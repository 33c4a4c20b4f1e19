"""One verify prefill of N tokens (for ncu launch lists / captures of the prefill kernels)."""
import sys
sys.path.insert(0, '.')
import numpy as np
from bench import ensure_model
from blama_b200 import capi, gguf_synth
shape = sys.argv[1] if len(sys.argv) > 1 else "llama-3.1-8b-q4km"
nv = int(sys.argv[2]) if len(sys.argv) > 2 else 2048
path = ensure_model(shape, 0, lambda: None)
m = capi.Model(path); c = capi.Ctx(m, nv + 64)
toks = gguf_synth.synth_prompt(shape, nv, 2)
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 2
for rep in range(reps):
    c.clear(); c.decode(gguf_synth.synth_prompt(shape, 32, 1))
    rpt = c.profile_verify(toks)
print(rpt)
