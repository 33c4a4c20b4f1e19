"""Host sampler (top-k candidate list from the device) against the oracle's full-vocabulary restatement of the
llama.cpp chain blama configures (reference Sampler.cpp:30-95): same tokens, draw for draw."""
import numpy as np
import pytest


@pytest.fixture(scope="module")
def host():
    from blama_b200 import host_api

    host_api.lib()
    return host_api


@pytest.mark.parametrize("seed,temp,top_p", [(0, 0.8, 0.95), (1717, 0.8, 0.95), (3, 1.3, 0.5), (9, 0.2, 1.0), (5, 0.0, 0.95)])
def test_top40_shortcut_equals_full_vocab_chain(seed, temp, top_p, oracle, host):
    rng = np.random.default_rng(seed + 100)
    V = 5000
    for trial in range(8):
        logits = (rng.standard_normal(V) * 2.0).astype(np.float32)
        cand = oracle.topk(logits, 40)                       # what the device hands over
        n_draws = 50
        full = oracle.Sampler(seed, temp, top_p)
        want = [full.sample(logits) for _ in range(n_draws)]
        got = host.sampler_draw(cand, n_draws, seed=seed, temp=temp, top_p=top_p)
        assert got.tolist() == want
        # the oracle's own candidate entry point agrees too
        short = oracle.Sampler(seed, temp, top_p)
        assert [short.sample_candidates(cand) for _ in range(n_draws)] == want
        full.close(); short.close()


def test_reset_reseeds(oracle, host):
    rng = np.random.default_rng(1)
    logits = (rng.standard_normal(2000) * 2.0).astype(np.float32)
    s = oracle.Sampler(42, 0.8, 0.95)
    first = [s.sample(logits) for _ in range(10)]
    s.reset()
    assert [s.sample(logits) for _ in range(10)] == first
    cand = oracle.topk(logits, 40)
    assert host.sampler_draw(cand, 10, seed=42).tolist() == first


def test_wider_topk_uses_whole_list(oracle, host):
    rng = np.random.default_rng(2)
    logits = (rng.standard_normal(300) * 2.0).astype(np.float32)
    full = oracle.Sampler(7, 0.9, 0.9, top_k=0, min_p=0.0)
    want = [full.sample(logits) for _ in range(30)]
    allc = np.zeros(300, dtype=host.TD_DTYPE)
    allc["token"] = np.arange(300); allc["logit"] = logits
    got = host.sampler_draw(allc, 30, seed=7, temp=0.9, top_p=0.9, top_k=0, min_p=0.0, is_sorted=False)
    assert got.tolist() == want


def test_unsupported_configurations_fail_loudly(host):
    from blama_b200.host_api import HostError

    cand = np.zeros(4, dtype=host.TD_DTYPE)
    cand["token"] = np.arange(4); cand["logit"] = [4, 3, 2, 1]
    assert host.sampler_draw(cand, 3, temp=0.0).tolist() == [0, 0, 0]          # greedy
    with pytest.raises(HostError):
        host.sampler_draw(np.zeros(0, dtype=host.TD_DTYPE), 1)
