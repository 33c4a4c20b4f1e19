"""The reference-shaped host API (bl::llama::Model / Instance / Session over the C ABI) on a GPU:
the reference's own session tests restated (inference/test/t-integration.cpp:124-248) + parity with the oracle's Session."""
import numpy as np
import pytest

from blama_b200 import gguf_synth as gs

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def host():
    from blama_b200 import host_api

    host_api.lib()
    return host_api


@pytest.fixture(scope="module")
def hmodel(host, gguf_path):
    m = host.Model(gguf_path("small-llama-q4km"))
    yield m
    m.close()


def test_session_lifecycle_error_strings(host, hmodel):
    """exact exception strings of the reference (t-integration.cpp:137-217)"""
    inst = host.Instance(hmodel, 256)
    inst.warmup()
    inst.start_session()
    with pytest.raises(host.HostError, match="^Session hasn't started yet$"):
        inst.complete(1)
    with pytest.raises(host.HostError, match="^Session hasn't started yet$"):
        inst.stream(1)
    with pytest.raises(host.HostError, match="^Session hasn't started yet$"):
        inst.get_state()
    with pytest.raises(host.HostError, match="^Session is already started. Stop it to start a new one.$"):
        inst.start_session()
    inst.set_initial_prompt(gs.synth_prompt("small-llama-q4km", 5, 1))
    with pytest.raises(host.HostError, match="^Session already started$"):
        inst.set_state()
    with pytest.raises(host.HostError, match="^Session already started$"):
        inst.set_initial_prompt([1, 2])
    inst.stop_session()
    inst.start_session()
    with pytest.raises(host.HostError, match=r"^Initial prompt too long. Got 300 tokens, max: 252$"):
        inst.set_initial_prompt(list(range(1, 301)))
    inst.stop_session()
    inst.close()
    with pytest.raises(host.HostError):
        host.Model("/nonexistent/model.gguf")
    with pytest.raises(host.HostError, match="no CPU backend"):
        host.Model("/nonexistent/model.gguf", gpu=False)


def test_filling_ctx_is_bit_equal_in_sequential_mode(host, hmodel):
    """reference "filling ctx" test (t-integration.cpp:219-248): CHECK(l.token == l2.token); CHECK(l.logit == l2.logit)"""
    prompt = gs.synth_prompt("small-llama-q4km", 9, 4)
    a, b = host.Instance(hmodel, 256), host.Instance(hmodel, 256)
    a.start_session(seed=3).set_initial_prompt(prompt)
    b.start_session(seed=3, sequential_verify=True).set_initial_prompt(prompt)
    toks, top = a.complete(10)
    assert len(toks) == 10
    out, out_n = b.fill_ctx(toks, top)
    assert np.all(out_n == 10)
    assert np.array_equal(out["token"], top["token"]) and np.array_equal(out["logit"], top["logit"])
    metrics = [host.lc_compare(top[i], out[i]) for i in range(10)]
    assert host.lc_score(metrics) == 1.0
    a.close(); b.close()


def test_complete_matches_oracle_session(host, gguf_path, oracle):
    """same seed, same prompt: the CUDA Session and the oracle's restated Session emit the same tokens and top-10 while
    the arithmetic stays on the clean branch (see test_gpu_model for the flip discussion); logits always within FLIP_TOL"""
    name = "tiny-llama-q8"
    path = gguf_path(name)
    prompt = gs.synth_prompt(name, 10, 6)
    om = oracle.Model(path)
    oc = oracle.Ctx(om, 128, oracle.MODE_GGML, 2)
    want_t, want_top = oc.complete(prompt, 16, seed=5)
    hm = host.Model(path)
    inst = host.Instance(hm, 128)
    inst.start_session(seed=5).set_initial_prompt(prompt)
    got_t, got_top = inst.complete(16)
    n_same = 0
    for i in range(min(len(want_t), len(got_t))):
        if want_t[i] != got_t[i]:
            break
        n_same += 1
        assert np.abs(got_top[i]["logit"] - want_top[i]["logit"]).max() <= 0.25
    assert n_same >= 4, (want_t, got_t)
    inst.close(); hm.close(); oc.close(); om.close()


def test_streaming_equals_complete(host, hmodel):
    prompt = gs.synth_prompt("small-llama-q4km", 6, 8)
    a, b = host.Instance(hmodel, 128), host.Instance(hmodel, 128)
    a.start_session(seed=9).set_initial_prompt(prompt)
    b.start_session(seed=9).set_initial_prompt(prompt)
    toks, _ = a.complete(7)
    assert np.array_equal(b.stream(7), toks)
    # follow-up prompt on a live session (t-integration.cpp:172-183 shape)
    more, _ = a.complete(2, prompt=gs.synth_prompt("small-llama-q4km", 4, 10))
    assert len(more) == 2
    a.close(); b.close()


def test_tokenizer_round_trip_on_the_loaded_model(host, hmodel):
    # the BPE tokenizer itself is pinned host-side in tests/test_tokenizer.py; here: the same answers through a device-resident model
    text = "The first man to walk on the moon, in July 1969."
    ids = hmodel.tokenize(text, add_special=False)
    assert b"".join(hmodel.token_to_bytes(int(t)) for t in ids).decode() == text
    vo = host.Model(hmodel.path, vocab_only=True) if hasattr(hmodel, "path") else None
    if vo is not None:
        assert vo.tokenize(text, add_special=False).tolist() == ids.tolist()
        vo.close()
    with_bos = hmodel.tokenize(text, add_special=True)
    assert len(with_bos) == len(ids) + 1 and with_bos[1:].tolist() == ids.tolist()
