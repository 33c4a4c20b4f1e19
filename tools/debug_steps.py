import sys, numpy as np
sys.path.insert(0, '.')
from blama_b200 import capi, gguf_synth as gs
from oracle import pyoracle as po
for name in sys.argv[1:]:
    path = f"/tmp/{name}.gguf"; gs.write_gguf(path, name)
    toks = gs.synth_prompt(name, 6, 3)
    om = po.Model(path); oc = po.Ctx(om, 256, po.MODE_GGML, 4)
    m = capi.Model(path); c = capi.Ctx(m, 256)
    for i, t in enumerate(toks):
        want = oc.decode([t])[0]; c.decode([int(t)]); got = c.logits()
        print(name, 'step', i, 'max|d|', float(np.abs(got-want).max()), 'std', float(want.std()))
