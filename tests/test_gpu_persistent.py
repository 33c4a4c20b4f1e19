"""The persistent cooperative decode kernel (mega_decode.cuh) against the per-op CUDA-graph path (BLK_MEGA=0) and the oracle:
same arithmetic contract, different kernels.  Covers what the short model tests do not reach: contexts long enough for several
attention tiles per CTA, the path selection, a body-only step (no lm_head) followed by a full one, many tokens in one loop."""
import numpy as np
import pytest

from blama_b200 import gguf_synth as gs

pytestmark = pytest.mark.gpu

FLIP_TOL = 0.25        # see tests/test_gpu_model.py / blama_b200/parity_stats.py: bound of a flipped Q8_K / f16 rounding
CLEAN_TOL = 1e-4


def _two_paths(path, n_ctx, monkeypatch):
    """(persistent-kernel context, per-op context) on the same GGUF; BLK_MEGA is read when the model is loaded"""
    from blama_b200 import capi

    monkeypatch.delenv("BLK_MEGA", raising=False)
    m1 = capi.Model(path); c1 = capi.Ctx(m1, n_ctx)
    monkeypatch.setenv("BLK_MEGA", "0")
    m0 = capi.Model(path); c0 = capi.Ctx(m0, n_ctx)
    monkeypatch.delenv("BLK_MEGA", raising=False)
    return (m1, c1), (m0, c0)


@pytest.mark.parametrize("name", ["small-llama-q4km", "small-qwen2-q8", "small-llama70-q4km", "tiny-llama-q8",
                                  "small-llama-wideffn-q4km", "small-qwen2-wideffn-q8"])
def test_path_selection_and_agreement(name, gguf_path, monkeypatch):
    (m1, c1), (m0, c0) = _two_paths(gguf_path(name), 256, monkeypatch)
    assert c1.persistent_decode and not c0.persistent_decode
    clean = total = 0
    for seq in range(6):                      # independent short sequences: a flipped rounding only taints its own sequence
        c1.clear(); c0.clear()
        for t in gs.synth_prompt(name, 6, 40 + seq):
            c1.decode([int(t)]); c0.decode([int(t)])
            err = float(np.abs(c1.logits() - c0.logits()).max())
            assert err <= FLIP_TOL, (seq, err)
            clean += err <= CLEAN_TOL; total += 1
    assert clean >= 0.3 * total, (clean, total)      # the same arithmetic up to fp32 summation order
    for c, m in ((c1, m1), (c0, m0)):
        c.close(); m.close()


def test_f32_weights_take_the_per_op_path(gguf_path):
    from blama_b200 import capi

    m = capi.Model(gguf_path("tiny-llama-f32")); c = capi.Ctx(m, 64)
    assert not c.persistent_decode
    c.decode([1, 2, 3]); assert c.n_past == 3
    c.close(); m.close()


@pytest.mark.parametrize("n_fill", [2200, 4500])
def test_long_context_several_attention_tiles(n_fill, gguf_path, oracle, monkeypatch):
    """n_head_kv = 2 -> 32 context splits of 64-token tiles: past 2048 tokens a CTA walks several tiles of its slice with a running
    maximum (the first tile is prefetched with cp.async during the QKV phase, the later ones travel behind the computation; at 4500
    tokens three tiles, the token being decoded in the last one); also crosses KV page boundaries."""
    name = "small-llama-q4km"
    (m1, c1), (m0, c0) = _two_paths(gguf_path(name), n_fill + 200, monkeypatch)
    fill = gs.synth_prompt(name, n_fill, 11)
    c1.decode(fill); c0.decode(fill)                      # tcgen05 prefill on both (identical kernels)
    assert np.array_equal(c1.logits(), c0.logits())
    toks = gs.synth_prompt(name, 12, 12)
    worst = 0.0
    for t in toks:
        c1.decode([int(t)]); c0.decode([int(t)])
        worst = max(worst, float(np.abs(c1.logits() - c0.logits()).max()))
        assert np.array_equal(c1.topk(10)["logit"], np.sort(c1.logits())[::-1][:10])
    assert worst <= FLIP_TOL, worst
    for c, m in ((c1, m1), (c0, m0)):
        c.close(); m.close()


def test_body_steps_then_head_and_device_loop(gguf_path, monkeypatch):
    """blk_decode(n < prefill_min) = n-1 steps without the lm_head + one with; blk_decode_loop = greedy feedback on the device"""
    from blama_b200 import capi

    name = "small-llama-q4km"
    m = capi.Model(gguf_path(name)); a = capi.Ctx(m, 512); b = capi.Ctx(m, 512)
    toks = gs.synth_prompt(name, 9, 21)
    a.decode(toks)                                        # 8 body-only launches + 1 full
    for t in toks:
        b.decode([int(t)])                                # 9 full launches
    assert np.array_equal(a.logits(), b.logits())
    first = int(a.topk(1)["token"][0])
    last = a.decode_loop(first, 40)                       # 40 launches back to back, arg-max fed back on the device
    tok = first
    for _ in range(40):
        tok = int(b.decode_topk(tok, 1)["token"][0])      # the same through the host
    assert last == tok and a.n_past == b.n_past
    a.close(); b.close(); m.close()


def test_run_to_run_determinism(gguf_path):
    """the same request twice gives the same bits: decode (LL hand-offs, fixed summation orders) and batched verify"""
    from blama_b200 import host_api

    name = "small-llama-q4km"
    hm = host_api.Model(gguf_path(name)); inst = host_api.Instance(hm, 1400)
    prompt = gs.synth_prompt(name, 12, 5)
    runs = []
    for _ in range(2):
        inst.start_session(seed=3); inst.set_initial_prompt(prompt)
        toks, top = inst.complete(1100)                   # long enough for the two-pass GEMM form in the verify below
        inst.stop_session()
        runs.append((np.asarray(toks), np.ascontiguousarray(top)))
    assert np.array_equal(runs[0][0], runs[1][0]) and runs[0][1].tobytes() == runs[1][1].tobytes()
    scores = []
    for _ in range(2):
        inst.start_session(seed=3); inst.set_initial_prompt(prompt)
        scores.append(inst.verify(runs[0][0], runs[0][1]))
        inst.stop_session()
    assert scores[0] == scores[1] and scores[0] >= 0.98, scores
    inst.close(); hm.close()
