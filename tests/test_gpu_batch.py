"""Continuous batching (not in the reference, whose Server serialises requests: Server.cpp:36).
  * blk_decode_batch: one forward pass for several sequences of different lengths, against the oracle's BF16 mode per sequence
    (the batched step runs in the bf16 operand arithmetic of the verify prefill);
  * a Server with max_batch > 1: concurrent /complete requests all finish with the requested length, each response verifies against a
    serial verifier at the reference's bars, a lone request is bit-identical to the unbatched Server, requests may join mid-flight."""
import numpy as np
import pytest

from blama_b200 import gguf_synth as gs

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", ["small-llama-q4km", "small-qwen2-q8", "tiny-llama-q8"])
def test_decode_batch_matches_oracle_bf16(name, gguf_path, oracle):
    from blama_b200 import capi

    path = gguf_path(name)
    m = capi.Model(path)
    om = oracle.Model(path)
    lens = [40, 33, 70, 100, 36]                      # all >= prefill_min: the prompts run the bf16 prefill too (like for like); 70 / 100 cross a KV page
    ctxs = [capi.Ctx(m, 256) for _ in lens]
    ocs = [oracle.Ctx(om, 256, oracle.MODE_BF16, 4) for _ in lens]
    ws = capi.Ctx(m, 256)                             # a workspace context that is not part of the batch
    for i, n in enumerate(lens):
        p = gs.synth_prompt(name, n, 50 + i)
        ctxs[i].decode(p); ocs[i].decode(p)
    rng = np.random.default_rng(4)
    worst = 0.0
    for step in range(5):
        toks = [int(t) for t in gs.synth_prompt(name, len(lens), 90 + step)]
        top = capi.decode_batch(ws, ctxs, toks, 10)
        for i in range(len(lens)):
            want = ocs[i].decode([toks[i]])[0]
            worst = max(worst, float(np.abs(top[i]["logit"] - want[top[i]["token"]]).max()))
            assert np.all(np.diff(top[i]["logit"]) <= 0)
            # the row's top-10 agrees with the reference wherever the reference's gaps pin a rank (parity_stats rule, d = 2 x deviation)
            ref_ids = np.argsort(-want, kind="stable")[:11]
            d = float(np.abs(top[i]["logit"] - want[top[i]["token"]]).max())
            for r in range(10):
                above = want[ref_ids[r - 1]] - want[ref_ids[r]] if r else np.inf
                if above > 2 * d + 1e-3 and want[ref_ids[r]] - want[ref_ids[r + 1]] > 2 * d + 1e-3:
                    assert int(top[i]["token"][r]) == int(ref_ids[r]), (name, step, i, r)
            assert ctxs[i].n_past == lens[i] + step + 1
        if step == 2:                                  # a subset of the sequences: the others simply do not advance
            sub = [ctxs[0], ctxs[3]]
            t2 = [int(t) for t in gs.synth_prompt(name, 2, 200)]
            top2 = capi.decode_batch(ctxs[0], sub, t2, 10)          # the workspace may be one of the batch
            for j, i in enumerate((0, 3)):
                want = ocs[i].decode([t2[j]])[0]
                worst = max(worst, float(np.abs(top2[j]["logit"] - want[top2[j]["token"]]).max()))
                lens[i] += 1
    print(f"\n[batched decode {name}] max |dlogit| on the top-10 vs oracle(BF16) = {worst:.3g}")
    assert worst <= 0.15, worst
    # the caches the batched steps wrote are the ones a single step continues from
    nxt = int(gs.synth_prompt(name, 1, 300)[0])
    ctxs[1].decode([nxt])
    want = ocs[1].decode([nxt])[0]
    assert float(np.abs(ctxs[1].logits() - want).max()) <= 0.5       # batch-1 kernel = int8 arithmetic on a bf16-built cache
    with pytest.raises(capi.BlkError):
        capi.decode_batch(ws, [ctxs[0], ctxs[0]], [1, 2], 10)         # the same context twice
    for c in ctxs + [ws]:
        c.close()
    for o in ocs:
        o.close()
    m.close(); om.close()


def test_batching_server(gguf_path):
    from blama_b200 import host_api as H

    name = "small-llama-q4km"
    H.lib()
    m = H.Model(gguf_path(name))
    plain = H.Server([m], ctx_size=512)
    batched = H.Server([m], ctx_size=512, max_batch=4)
    prompts = [gs.synth_prompt(name, 6 + 3 * k, 70 + k) for k in range(7)]
    # a lone request on the batching server takes the batch-1 kernel: bit-identical to the plain server
    a = plain.wait_complete(plain.submit_complete(prompts[0], 20, seed=3))
    b = batched.wait_complete(batched.submit_complete(prompts[0], 20, seed=3))
    assert np.array_equal(a[0], b[0]) and np.array_equal(a[1]["logit"], b[1]["logit"])
    # seven concurrent requests of different lengths on four slots: they join and leave between steps
    want_len = [24, 40, 8, 33, 16, 40, 5]
    tickets = [batched.submit_complete(p, n, seed=k) for k, (p, n) in enumerate(zip(prompts, want_len))]
    # ... and a verify request queued in the middle of them is served between two steps
    vt = batched.submit_verify(prompts[0], a[0], a[1], a[2], seed=3)
    got = [batched.wait_complete(t) for t in tickets]
    assert batched.wait_verify(vt) >= 0.95
    assert batched.last_worker_error() == ""
    for k, g in enumerate(got):
        assert len(g[0]) == want_len[k] and np.all(g[2] == 10)
        assert np.all(np.diff(g[1]["logit"], axis=1) <= 0)
        # every response, whichever mix of batched and single steps produced it, verifies at the reference's bars
        s = plain.wait_verify(plain.submit_verify(prompts[k], g[0], g[1], g[2], seed=k))
        assert s >= 0.95, (k, s)
    st = batched.stats()
    assert st[0]["requests"] >= 9 and st[0]["gpu_ms"] > 0
    # a request that fails (prompt longer than the context) answers empty and the others carry on
    t_bad = batched.submit_complete(gs.synth_prompt(name, 600, 1), 4)
    t_ok = batched.submit_complete(prompts[1], 6, seed=1)
    assert len(batched.wait_complete(t_bad)[0]) == 0 and len(batched.wait_complete(t_ok)[0]) == 6
    assert "Initial prompt too long" in batched.last_worker_error()
    batched.close(); plain.close(); m.close()
