"""Two builds of the library on the same decode workload (run each in its own process; BLAMA_B200_LIB selects the build): tok/s of the
bench loop and an md5 of the logits of a short teacher-forced run, so that a change meant to be bit-neutral can be checked.
    python tools/decode_lib_ab.py [shape] [lib.so ...]"""
import hashlib, json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CHILD = r'''
import hashlib, json, os, sys
sys.path.insert(0, %r)
import numpy as np
from blama_b200 import capi, gguf_synth as gs
shape = sys.argv[1]
path = f"/dev/shm/blama_b200_{shape}.gguf"
if not os.path.exists(path): gs.write_gguf(path, shape)
m = capi.Model(path); c = capi.Ctx(m, 1024)
prompt = gs.synth_prompt(shape, 512, 1)
rates = []
for rep in range(3):
    c.clear(); c.flush_l2(); c.decode(prompt)
    first = int(c.topk(1)["token"][0])
    c.timer_start(); c.decode_loop(first, 256, wait=False); ms = c.timer_stop()
    rates.append(round(256 / ms * 1e3, 1))
h = hashlib.md5()
c.clear(); c.decode(gs.synth_prompt(shape, 600, 3))
for t in gs.synth_prompt(shape, 6, 2):
    c.decode([int(t)]); h.update(np.ascontiguousarray(c.logits()).tobytes())
print(json.dumps({"lib": os.environ.get("BLAMA_B200_LIB", "default"), "tok_s": rates, "logits_md5": h.hexdigest()}))
''' % ROOT
args = sys.argv[1:]
shape = args[0] if args and not args[0].endswith(".so") else "llama-3.1-8b-q4km"
libs = [a for a in args if a.endswith(".so")] or [""]
for lib in libs:
    env = dict(os.environ)
    if lib: env["BLAMA_B200_LIB"] = os.path.abspath(lib)
    print(subprocess.run([sys.executable, "-c", CHILD, shape], env=env, capture_output=True, text=True).stdout.strip(), flush=True)
