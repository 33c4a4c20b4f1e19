"""The request dispatcher (reference server/code/server/Server.cpp:23-161) with N workers, its JSON bodies
(server/code/http/HttpServerMain.cpp:37-94, 255-288) and the HTTP front end (:298-358), on a small random-init model.
  * N queued /complete + /verify_completion jobs over two workers give exactly what serial sessions give;
  * a request that throws answers its callback (empty response / NaN) and leaves the worker alive (the reference dies there);
  * server/code/http/test.rb:6-25 replayed: POST /complete, then POST /verify_completion with {request, response}; the score of the
    re-parsed JSON equals the in-memory one bit for bit."""
import http.client
import json
import math

import numpy as np
import pytest

from blama_b200 import gguf_synth as gs

pytestmark = pytest.mark.gpu
NAME = "small-llama-q4km"


@pytest.fixture(scope="module")
def srv(gguf_path):
    from blama_b200 import host_api as H

    H.lib()
    m = H.Model(gguf_path(NAME))
    s = H.Server([m, m], ctx_size=512)            # two workers (two Instances) on one device, sharing the model like t-integration.cpp:220-224
    yield H, m, s
    s.close(); m.close()


def serial_complete(H, m, prompt, n, seed):
    i = H.Instance(m, 512)
    i.start_session(seed=seed).set_initial_prompt(prompt)
    toks, top = i.complete(n)
    i.close()
    return toks, top


def test_queued_jobs_equal_serial_sessions(srv):
    H, m, s = srv
    assert s.workers() == 2
    prompts = [gs.synth_prompt(NAME, 5 + k, 40 + k) for k in range(6)]
    tickets = [s.submit_complete(p, 12, seed=k) for k, p in enumerate(prompts)]
    got = [s.wait_complete(t) for t in tickets]
    for k, p in enumerate(prompts):
        toks, top = serial_complete(H, m, p, 12, k)
        assert np.array_equal(got[k][0], toks)
        assert np.array_equal(got[k][1]["token"], top["token"]) and np.array_equal(got[k][1]["logit"], top["logit"])
    # verify jobs, both workers busy: every response scores what a serial session gives
    vt = [s.submit_verify(prompts[k], got[k][0], got[k][1], got[k][2], seed=k) for k in range(6)]
    scores = [s.wait_verify(t) for t in vt]
    for k in range(6):
        i = H.Instance(m, 512)
        i.start_session(seed=k).set_initial_prompt(prompts[k])
        assert i.verify(got[k][0], got[k][1], got[k][2]) == scores[k]
        i.close()
        assert scores[k] >= 0.95
    st = s.stats()
    assert sum(w["requests"] for w in st) >= 12 and all(w["gpu_ms"] > 0 for w in st if w["requests"])


def test_failing_request_answers_and_worker_survives(srv):
    H, m, s = srv
    p = gs.synth_prompt(NAME, 6, 3)
    toks, top = serial_complete(H, m, p, 8, 1)
    bad = toks.copy(); bad[3] = 10 ** 7                       # token id outside the vocabulary: fillCtx throws on the worker
    t_bad = s.submit_verify(p, bad, top, seed=1)
    t_ok = s.submit_verify(p, toks, top, seed=1)
    assert math.isnan(s.wait_verify(t_bad))                   # the callback was answered (the reference would have terminated)
    assert s.wait_verify(t_ok) >= 0.95                        # ... and the workers still serve
    assert "decode" in s.last_worker_error().lower() or "token" in s.last_worker_error().lower()
    # a prompt longer than the context: setInitialPrompt throws with the reference's text
    t_long = s.submit_complete(gs.synth_prompt(NAME, 600, 1), 4)
    assert len(s.wait_complete(t_long)[0]) == 0
    assert "Initial prompt too long" in s.last_worker_error()
    # more than 10 claimed logits per token cross the C ABI as an error before anything is queued
    with pytest.raises(H.HostError):
        s.submit_verify(p, toks, top, np.full(len(toks), 11, dtype=np.int32))
    # positions without claimed logits cannot be verified: worst-case metrics instead of the reference's out-of-bounds read
    t_empty = s.submit_verify(p, toks, top, np.zeros(len(toks), dtype=np.int32))
    assert s.wait_verify(t_empty) == 0.0


def test_json_round_trip_replays_test_rb(srv):
    H, m, s = srv
    body = json.dumps({"prompt": "The first man to", "max_tokens": 20})                 # server/code/http/test.rb:6-9
    answer = s.wait_complete_json(s.submit_complete_json(body))
    parsed = json.loads(answer)
    assert set(parsed) == {"text", "tokenData"} and len(parsed["tokenData"]) == 20
    assert parsed["text"] == "".join(t["str"] for t in parsed["tokenData"])
    assert all(set(t) == {"id", "logits", "str"} and len(t["logits"]) == 10 for t in parsed["tokenData"])
    verify_body = '{"request":' + body + ',"response":' + answer + "}"                # test.rb:17-20
    result = json.loads(s.wait_verify_json(s.submit_verify_json(verify_body)))
    # in memory: the same request through the token-level entry points
    prompt = m.tokenize("The first man to", add_special=True)
    toks, top, nl = s.wait_complete(s.submit_complete(prompt, 20, seed=0))
    assert [t["id"] for t in parsed["tokenData"]] == toks.tolist()
    mem = s.wait_verify(s.submit_verify(prompt, toks, top, nl, seed=0))
    assert np.float32(result["result"]) == np.float32(mem)                              # same verdict float across the wire
    assert mem >= 0.95
    with pytest.raises(H.HostError):
        s.submit_complete_json('{"max_tokens": 3}')                                     # no prompt


def test_http_front_end(srv):
    H, m, s = srv
    port = s.http_start("127.0.0.1", 0)
    try:
        def post(target, body, method="POST"):
            c = http.client.HTTPConnection("127.0.0.1", port, timeout=60)
            c.request(method, target, body=body, headers={"Content-Type": "text/json"})
            r = c.getresponse()
            data = r.read()
            hdrs = {k.lower(): v for k, v in r.getheaders()}
            c.close()
            return r.status, hdrs, data

        st, hd, data = post("/complete", json.dumps({"prompt": "The first man to", "max_tokens": 8, "seed": 3}))
        assert st == 200 and hd["server"] == "Beast" and hd["content-type"] == "text/json" and hd["access-control-allow-origin"] == "*"
        resp = json.loads(data)
        assert len(resp["tokenData"]) == 8
        st, hd, data = post("/verify_completion", json.dumps({"request": {"prompt": "The first man to", "max_tokens": 8, "seed": 3}, "response": resp}))
        assert st == 200 and hd["content-type"] == "text/json"
        assert json.loads(data)["result"] >= 0.95
        assert post("/complete", "", method="GET")[0] == 400                            # only POST is served (HttpServerMain.cpp:307-311)
        assert post("/nope", "{}")[0] == 404
        assert post("/chat/completions", "{}")[0] == 501                                # chat templates are out of scope
        st, hd, data = post("/complete", '{"prompt": ')
        assert st == 400 and "error" in json.loads(data)                                # the reference dies on a malformed body
    finally:
        s.http_stop()
