"""A/B of decode-kernel variants on one GPU: for every configuration (a set of BLK_* environment switches, read at model load /
context creation) the decode tok/s of the bench workload (512-token prompt + 256 greedy tokens, device-resident loop, CUDA events)
and the deviation of its logits from the first configuration's on a short teacher-forced run.
    python tools/decode_ab.py [shape] base: local0:BLK_ATTN_LOCAL=0 qkv2:BLK_MEGA_W_QKV=2 ..."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from blama_b200 import capi, gguf_synth as gs  # noqa: E402


def main():
    args = sys.argv[1:]
    shape = "llama-3.1-8b-q4km"
    if args and ":" not in args[0]:
        shape, args = args[0], args[1:]
    configs = []
    for a in args or ["base:"]:
        name, _, envs = a.partition(":")
        configs.append((name, dict(kv.split("=", 1) for kv in envs.split(",") if kv)))
    path = f"/dev/shm/blama_b200_{shape}.gguf"
    if not os.path.exists(path):
        gs.write_gguf(path, shape)
    prompt = gs.synth_prompt(shape, 512, 1)
    probe = [int(t) for t in gs.synth_prompt(shape, 12, 2)]
    ref_rows = None
    keys = sorted({k for _, e in configs for k in e})
    for name, env in configs:
        for k in keys:
            os.environ.pop(k, None)
        os.environ.update(env)
        m = capi.Model(path)
        c = capi.Ctx(m, 1024)
        rates = []
        for rep in range(3):
            c.clear(); c.flush_l2()
            c.decode(prompt)
            first = int(c.topk(1)["token"][0])
            c.timer_start()
            c.decode_loop(first, 256, wait=False)
            ms = c.timer_stop()
            rates.append(256 / ms * 1e3)
        ms_k, nbytes = c.bench_kernel(5, 32) if c.persistent_decode else (0.0, 0)
        # parity probe: 12 single steps from an empty context, then 4 more behind a 600-token prefill (several attention splits)
        rows = []
        c.clear()
        for t in probe:
            c.decode([t]); rows.append(c.logits())
        c.clear(); c.decode(gs.synth_prompt(shape, 600, 3))
        for t in probe[:4]:
            c.decode([t]); rows.append(c.logits())
        if ref_rows is None:
            ref_rows = rows
        dev = [float(np.abs(a - b).max()) for a, b in zip(rows, ref_rows)]
        print(json.dumps({"config": name, "env": env, "persistent": bool(c.persistent_decode), "tok_s": [round(r, 1) for r in rates],
                          "kernel_ms": round(ms_k, 4), "kernel_gbs": round(nbytes / ms_k / 1e6, 1) if ms_k else None,
                          "max_dev_short": round(max(dev[:12]), 4), "max_dev_long": round(max(dev[12:]), 4)}), flush=True)
        c.close(); m.close()


if __name__ == "__main__":
    main()
