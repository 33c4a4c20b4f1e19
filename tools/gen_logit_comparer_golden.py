"""Generates tests/golden/logit_comparer_golden.json by running the REFERENCE's own LogitComparer (compiled from
/root/reference into oracle/_ref by oracle/Makefile) on seeded inputs.  Run in the build container only; the GPU box
has no /root/reference and uses the committed fixture."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import pyoracle as po  # noqa: E402


def f32(x):
    return float(np.float32(x))


def main():
    assert po.ref_lib() is not None, "oracle/_ref/libref_logitcomparer.so missing: run make -C oracle (needs /root/reference)"
    rng = np.random.default_rng(20261018)
    cases = []
    # the reference's own unit test (inference/test/t-LogitComparer.cpp:13-39)
    a = [(i, 17.5 - 0.5 * i) for i in range(10)]
    cases.append({"name": "t-LogitComparer compare - no model", "a": a, "b": a})
    cases.append({"name": "scaled x1.01", "a": a, "b": [(i, f32((17.5 - 0.5 * i) * 1.01)) for i in range(10)]})
    for k in range(200):
        n1, n2 = int(rng.integers(1, 11)), int(rng.integers(1, 11))
        ids1 = rng.choice(60, n1, replace=False)
        mode = k % 4
        if mode == 0:
            ids2 = rng.choice(60, n2, replace=False)
        elif mode == 1:
            ids2 = ids1[:n2] if n2 <= n1 else np.concatenate([ids1, 60 + np.arange(n2 - n1)])
        elif mode == 2:
            ids2 = rng.permutation(ids1)[:n2] if n2 <= n1 else np.concatenate([rng.permutation(ids1), 60 + np.arange(n2 - n1)])
        else:
            ids2 = ids1[:n2] if n2 <= n1 else np.concatenate([ids1, ids1[: n2 - n1]])   # duplicates
        l1 = np.sort(rng.normal(6, 3, n1).astype(np.float32))[::-1]
        base = np.resize(l1, len(ids2)).astype(np.float32)
        l2 = np.sort(base + rng.normal(0, 0.05 * (k % 5), len(ids2)).astype(np.float32))[::-1]
        cases.append({"name": f"random-{k}", "a": [(int(t), f32(l)) for t, l in zip(ids1, l1)], "b": [(int(t), f32(l)) for t, l in zip(ids2, l2)]})
    out = []
    metrics = []
    for c in cases:
        m = po.ref_compare(c["a"], c["b"])
        s = po.ref_similarity(c["a"], c["b"])
        metrics.append(m)
        out.append({**c, "top1Match": m[0], "distance": repr(m[1]), "jsd": repr(m[2]), "similarity": repr(s),
                    "metrics_hex": [np.float32(x).tobytes().hex() for x in m], "similarity_hex": np.float32(s).tobytes().hex()})
    finite = [m for m in metrics if all(np.isfinite(m))]
    scores = [np.float32(po.ref_score(finite[: n])).tobytes().hex() for n in (1, 2, 10, len(finite))]
    doc = {"generator": "tools/gen_logit_comparer_golden.py", "source": "reference inference/code/llama/LogitComparer.cpp compiled with g++ -O2 (oracle/Makefile target ref)",
           "cases": out, "score_prefix_lengths": [1, 2, 10, len(finite)], "score_hex": scores}
    path = os.path.join(ROOT, "tests", "golden", "logit_comparer_golden.json")
    with open(path, "w") as f:
        json.dump(doc, f, indent=0)
    print("wrote", path, len(out), "cases")


if __name__ == "__main__":
    main()
