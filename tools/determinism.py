"""Run-to-run determinism of the two hot paths on the bench model: the same request twice must give the same bits."""
import sys
sys.path.insert(0, '.')
import numpy as np
from bench import ensure_model
from blama_b200 import capi, gguf_synth, host_api
shape = sys.argv[1] if len(sys.argv) > 1 else "llama-3.1-8b-q4km"
path = ensure_model(shape, 0, lambda: None)
hm = host_api.Model(path); inst = host_api.Instance(hm, 2048 + 128)
prompt = gguf_synth.synth_prompt(shape, 32, 1)
runs = []
for rep in range(2):
    inst.start_session(seed=1); inst.set_initial_prompt(prompt)
    toks, top = inst.complete(1024)
    inst.stop_session()
    runs.append((np.asarray(toks), np.ascontiguousarray(top)))
print("complete(1024) twice: tokens equal", np.array_equal(runs[0][0], runs[1][0]), " top-10 logits bit-equal", runs[0][1].tobytes() == runs[1][1].tobytes())
scores = []
for rep in range(3):
    inst.start_session(seed=1); inst.set_initial_prompt(prompt)
    scores.append(inst.verify(runs[0][0], runs[0][1]))
    inst.stop_session()
print("verify x3 scores", [repr(s) for s in scores], "equal", len(set(scores)) == 1)
