"""Multi-token prefill path (tcgen05 GEMMs + flash attention + per-row top-10 / gather) on whole random-init models.

Parity bars:
  * against the oracle in ORC_MODE_BF16 (same operand rounding as the tensor-core GEMMs): logits within PREFILL_TOL
  * against the decode path / the reference arithmetic (int8 activations): LogitComparer score >= 0.99, i.e. far inside the
    reference's own cross-backend bar (score >= 0.95, t-LogitComparer.cpp:76-78); verdict identical"""
import numpy as np
import pytest

from blama_b200 import gguf_synth as gs
from blama_b200 import parity_stats as ps

pytestmark = pytest.mark.gpu

PREFILL_TOL = 0.08          # absolute, logits have std ~2: bf16 operand rounding + f16 flash-attention ordering
MODELS = ["tiny-llama-q4km", "tiny-qwen2-q8", "small-llama-q4km", "small-qwen2-q8", "small-llama70-q4km", "tiny-llama-f32", "small-llama-gq3", "small-qwen2-gq7"]


@pytest.mark.parametrize("name", MODELS)
@pytest.mark.parametrize("form", ["resident", "two-pass", "fused"])
def test_prompt_prefill_matches_bf16_oracle(name, form, gguf_path, oracle, monkeypatch):
    """form: the GEMMs read the model's resident bf16 panels (default) / a streaming panel de-quantised once per matrix and request
    (default above 256 tokens for matrices that are not resident) / de-quantise inside the GEMM; all with the deterministic split-K
    that few-token batches get"""
    from blama_b200 import capi

    # (without resident panels the default switches form at 257 tokens: pick explicitly so that both are covered at this size)
    if form != "resident":
        monkeypatch.setenv("BLK_PANEL_CACHE_GB", "0")
        monkeypatch.setenv("BLK_PANEL_MIN", "0" if form == "fused" else "32")

    path = gguf_path(name)
    toks = gs.synth_prompt(name, 75, 11)                     # 75 >= prefill_min and not a multiple of any tile size
    om = oracle.Model(path)
    oc = oracle.Ctx(om, 256, oracle.MODE_BF16, 4)
    want = oc.decode(toks, all_logits=True)
    m = capi.Model(path)
    c = capi.Ctx(m, 256)
    c.decode(toks)
    got = c.logits()
    assert np.abs(got - want[-1]).max() <= PREFILL_TOL
    pbytes, n_res, n_mat = m.panel_cache()
    if form == "resident" and "f32" not in name:
        assert n_res == n_mat and pbytes > 0                 # every layer matrix and the lm_head
    if form != "resident":
        assert n_res == 0 and pbytes == 0
    top = c.topk(10)
    assert np.array_equal(top["logit"], np.sort(got)[::-1][:10])
    # the KV cache written by the prefill is usable by the decode path: next-token logits stay close to the oracle's
    nxt = int(toks[3])
    want2 = oc.decode([nxt])[0]
    c.decode([nxt])
    assert np.abs(c.logits() - want2).max() <= 0.5           # decode continues in the int8 arithmetic: quantisation-noise bound
    assert c.n_past == 76
    c.close(); m.close(); oc.close(); om.close()


@pytest.mark.parametrize("name", ["small-llama-q4km", "small-qwen2-q8", "tiny-llama-q4km"])
def test_batched_verify_against_oracle_and_prover(name, gguf_path, oracle):
    from blama_b200 import capi, host_api

    path = gguf_path(name)
    prompt = gs.synth_prompt(name, 9, 21)
    m = capi.Model(path)
    prover, verifier = capi.Ctx(m, 512), capi.Ctx(m, 512)
    prover.decode(prompt)
    cur = prover.topk(40)
    toks, tops = [], []
    for i in range(90):
        tok = int(cur["token"][(i * 7) % 5])
        cur = prover.decode_topk(tok, 40)
        toks.append(tok); tops.append(cur[:10].copy())
    claimed = np.stack([t["token"] for t in tops])
    verifier.decode(prompt)
    g, top = verifier.verify_prefill(toks, claimed)          # one causal prefill of the 90 response tokens
    assert verifier.n_past == len(prompt) + 90
    # (a) against the bf16-mode oracle at every position
    om = oracle.Model(path)
    oc = oracle.Ctx(om, 512, oracle.MODE_BF16, 4)
    oc.decode(prompt)
    want = oc.decode(toks, all_logits=True)
    for i in range(90):
        assert np.abs(g[i] - want[i][claimed[i]]).max() <= 0.15, i      # prompt KV came from the int8 decode path on the GPU side
        # the verifier's own top-10: identical ids at every rank the reference's gaps pin (2 x the row's measured deviation)
        d = max(float(np.abs(g[i] - want[i][claimed[i]]).max()), float(np.abs(top[i]["logit"] - want[i][top[i]["token"]]).max()))
        ids11, val11 = ps.top_sorted(want[i], 11)
        for r in range(10):
            above = val11[r - 1] - val11[r] if r else np.inf
            if above > 2 * d + 1e-3 and val11[r] - val11[r + 1] > 2 * d + 1e-3:
                assert int(top[i]["token"][r]) == int(ids11[r]), (i, r, d)
    # (b) prover (decode path) vs verifier (prefill path): the reference's own acceptance metric
    metrics, sims = [], []
    for i in range(90):
        mine = np.zeros(10, dtype=capi.TD_DTYPE)
        mine["token"] = claimed[i]; mine["logit"] = g[i]
        mine = mine[np.argsort(-mine["logit"], kind="stable")]
        metrics.append(host_api.lc_compare(tops[i], mine))
        sims.append(host_api.lc_similarity(tops[i], mine))
    score = host_api.lc_score(metrics)
    assert score >= 0.99 and np.mean(sims) >= 0.98, (score, np.mean(sims))
    # the verifier can keep generating after the fill
    nxt = verifier.topk(10)
    assert nxt["logit"][0] >= nxt["logit"][9]
    prover.close(); verifier.close(); m.close(); oc.close(); om.close()


def test_chunked_prefill_equals_single_chunk(gguf_path):
    from blama_b200 import capi

    name = "small-llama-q4km"
    path = gguf_path(name)
    toks = gs.synth_prompt(name, 200, 31)
    m = capi.Model(path)
    a, b = capi.Ctx(m, 512, 2048), capi.Ctx(m, 512, 64)     # n_batch 64 -> 4 chunks (64, 64, 64, 8 -> the tail falls below prefill_min)
    a.decode(toks); b.decode(toks)
    assert np.abs(a.logits() - b.logits()).max() <= 0.3
    assert a.n_past == b.n_past == 200
    a.close(); b.close(); m.close()


@pytest.mark.parametrize("name", ["small-llama-q4km", "small-qwen2-q8"])
def test_sparse_claimed_logits_equal_the_full_head(name, gguf_path):
    """without the verifier's own top-10 (what fillCtx needs) the vocabulary projection is evaluated only at the claimed ids:
    same bf16 x bf16 -> f32 arithmetic as the GEMM form, so the gathered logits agree to fp32 summation order; ragged
    n_claimed, ids outside the vocabulary and the state left behind (n_past, last-position top-k) behave the same"""
    from blama_b200 import capi

    m = capi.Model(gguf_path(name))
    a, b = capi.Ctx(m, 512), capi.Ctx(m, 512)
    rng = np.random.default_rng(5)
    prompt = gs.synth_prompt(name, 40, 1)
    toks = gs.synth_prompt(name, 300, 2)
    claimed = rng.integers(0, m.n_vocab, size=(len(toks), 10)).astype(np.int32)
    claimed[7, 3] = m.n_vocab + 5; claimed[9, 0] = -1                  # outside the vocabulary -> -inf
    n_claimed = rng.integers(0, 11, size=len(toks)).astype(np.int32)
    n_claimed[7] = n_claimed[9] = 10
    a.decode(prompt); b.decode(prompt)
    g_full, top = a.verify_prefill(toks, claimed, n_claimed, want_top=True)
    g_sparse, none = b.verify_prefill(toks, claimed, n_claimed, want_top=False)
    assert none is None and a.n_past == b.n_past == len(prompt) + len(toks)
    finite = np.isfinite(g_full)
    assert np.array_equal(finite, np.isfinite(g_sparse))
    assert np.isneginf(g_sparse[7, 3]) and np.isneginf(g_sparse[9, 0])
    assert np.abs(g_full[finite] - g_sparse[finite]).max() <= 2e-3 * max(1.0, np.abs(g_full[finite]).max())
    for i in range(len(toks)):
        assert np.all(g_sparse[i, n_claimed[i]:] == 0.0)
    # the last position's row: bf16 GEMM (full head) vs the decode mat-vec (sparse path): same top-1 unless it is a near tie
    ta, tb = a.topk(10), b.topk(10)
    assert ta["token"][0] == tb["token"][0] or abs(ta["logit"][0] - ta["logit"][1]) < 0.2
    a.close(); b.close(); m.close()


@pytest.mark.parametrize("name", ["small-llama-q4km", "small-qwen2-q8"])
@pytest.mark.parametrize("T", [48, 300])
def test_resident_panels_give_the_streamed_forms_logits(name, T, gguf_path, monkeypatch):
    """the resident panels hold exactly what the streamed two-pass form de-quantises per request: the logits of a prefill are
    bit-identical between the two, below and above the 256-token switch (same GEMM kernel, same tiles, same split-K)"""
    from blama_b200 import capi

    path = gguf_path(name)
    toks = gs.synth_prompt(name, T, 5)
    rows = []
    for cache in ("default", "0"):
        if cache == "0":
            monkeypatch.setenv("BLK_PANEL_CACHE_GB", "0")
            monkeypatch.setenv("BLK_PANEL_MIN", "32")
        m = capi.Model(path); c = capi.Ctx(m, 512)
        c.decode(toks); rows.append(c.logits().copy())
        c.clear(); c.decode(toks)                            # second pass: the cache is warm / the stream panels are reused
        assert np.array_equal(rows[-1], c.logits())
        c.close(); m.close()
    assert np.array_equal(rows[0], rows[1])
