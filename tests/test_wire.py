"""JSON wire format of /complete and /verify_completion (reference server/code/http/HttpServerMain.cpp:37-94, 255-288).
Golden vectors were produced by nlohmann::json itself (tools/gen_json_golden.py): float text, key order, string escapes.
Host-only: no device needed."""
import json
import os
import struct

import numpy as np
import pytest

from blama_b200 import host_api as H

GOLD = os.path.join(os.path.dirname(__file__), "golden", "json_wire_golden.json")


@pytest.fixture(scope="module")
def gold():
    with open(GOLD) as f:
        return json.load(f)


def f32_of(bits_hex: str) -> np.float32:
    return np.frombuffer(struct.pack("<I", int(bits_hex, 16)), dtype=np.float32)[0]


def test_float_text_round_trips_and_matches_nlohmann(gold):
    """every float: our text parses back to the same float; it equals nlohmann's text except where Grisu2 is not shortest"""
    differ = 0
    for rec in gold["floats"]:
        f = f32_of(rec["bits"])
        body = H.wire_verify_json(float(f))
        assert body.startswith('{"result":') and body.endswith("}")
        text = body[len('{"result":'):-1]
        assert np.float32(float(text)) == f or (np.isnan(f) and text == "null")
        assert float(text) == float(f)                     # the widened double itself round-trips
        if text != rec["text"]:
            differ += 1
            assert float(rec["text"]) == float(text)       # same double, Grisu2 printed one digit more
    assert differ <= len(gold["floats"]) // 200            # Grisu2 is shortest in > 99.5 % of the cases


def test_layout_rules():
    cases = [(1.0, "1.0"), (-0.0, "-0.0"), (0.0, "0.0"), (17.5, "17.5"), (1e-4, None), (100000.0, "100000.0"), (0.5, "0.5"), (1e15, None), (1e16, None)]
    for v, want in cases:
        text = H.wire_verify_json(v)[len('{"result":'):-1]
        if want is not None:
            assert text == want
        assert np.float32(float(text)) == np.float32(v)
    assert H.wire_verify_json(float("nan")) == '{"result":null}'
    # exponent layout: below 1e-4 and above 1e15 the exponential form with two exponent digits
    assert H.wire_json_roundtrip("[1e-05, 0.0001, 1e14, 1e15, 123456789012345678]") == "[1e-05,0.0001,100000000000000.0,1e+15,123456789012345678]"


def test_complete_body_is_byte_identical_to_nlohmann(gold):
    toks, top, nl, strs = [], np.zeros((len(gold["token_records"]), 10), dtype=H.TD_DTYPE), [], []
    for i, rec in enumerate(gold["token_records"]):
        parts = rec.split()
        toks.append(int(parts[0])); n = int(parts[1]); nl.append(n)
        for j in range(n):
            top[i, j] = (int(parts[2 + 2 * j]), f32_of(parts[3 + 2 * j]))
        strs.append("" if parts[-1] == "-" else bytes.fromhex(parts[-1]).decode("utf-8"))
    # the body without a model carries empty token strings: compare through a parse of the golden with "str" blanked
    ours = H.wire_complete_json(toks, top, nl)
    want = json.loads(gold["complete_body"])
    want["text"] = ""
    for t in want["tokenData"]:
        t["str"] = ""
    got = json.loads(ours)
    assert got == want
    # key order is nlohmann's (std::map): id < logits < str, text < tokenData, id < logit
    assert ours.startswith('{"text":"","tokenData":[{"id":1000,"logits":[')
    # numbers: byte-identical text for every logit
    import re
    nums_ours = re.findall(r'"logit":([^,}\]]+)', ours)
    nums_gold = re.findall(r'"logit":([^,}\]]+)', gold["complete_body"])
    assert len(nums_ours) == len(nums_gold)
    assert sum(a != b for a, b in zip(nums_ours, nums_gold)) <= 1
    assert all(float(a) == float(b) for a, b in zip(nums_ours, nums_gold))
    assert H.wire_verify_json(float(np.float32(0.9973522424697876))) == gold["verify_body"]


def test_strings_and_key_order_round_trip(gold):
    # parse -> dump of nlohmann's own body reproduces it byte for byte (escapes, UTF-8 pass-through, key order, floats)
    again = H.wire_json_roundtrip(gold["complete_body"])
    if again != gold["complete_body"]:
        # only Grisu2's rare non-shortest digits may differ
        assert json.loads(again) == json.loads(gold["complete_body"])
        assert len(again) >= len(gold["complete_body"]) - 4
    shuffled = '{"tokenData": [], "text": "a\\u00e9\\n\\"", "zz": [true, false, null], "aa": {"b": -3, "a": 2.50}}'
    assert H.wire_json_roundtrip(shuffled) == '{"aa":{"a":2.5,"b":-3},"text":"aé\\n\\"","tokenData":[],"zz":[true,false,null]}'


def test_request_parsing_follows_the_reference():
    r = H.wire_parse_request('{"prompt": "The first man to", "max_tokens": 20}')       # server/code/http/test.rb:6-9
    assert r["prompt"] == "The first man to" and r["max_tokens"] == 20 and r["seed"] == 0
    assert r["temp"] == pytest.approx(0.8) and r["top_p"] == pytest.approx(0.95)        # Server.hpp:30-31 defaults
    r = H.wire_parse_request('{"prompt":"x","seed":7,"temp":0.5,"top_p":1,"suffix":"ignored","extra":[1,2]}')
    assert r["seed"] == 7 and r["temp"] == 0.5 and r["top_p"] == 1.0
    with pytest.raises(H.HostError):
        H.wire_parse_request('{"max_tokens": 3}')                                      # prompt is required (HttpServerMain.cpp:87)
    with pytest.raises(H.HostError):
        H.wire_parse_request('{"prompt": "x", ')                                       # malformed JSON
    with pytest.raises(H.HostError):
        H.wire_parse_request('{"prompt": 5}')


def test_verify_body_round_trip_keeps_every_float_bit():
    rng = np.random.default_rng(5)
    n = 64
    toks = rng.integers(0, 128256, n).astype(np.int32)
    top = np.zeros((n, 10), dtype=H.TD_DTYPE)
    top["token"] = rng.integers(0, 128256, (n, 10))
    top["logit"] = np.sort(rng.normal(0, 5, (n, 10)).astype(np.float32), axis=1)[:, ::-1]
    nl = rng.integers(0, 11, n).astype(np.int32)
    body = H.wire_complete_json(toks, top, nl)
    req = '{"prompt":"p","max_tokens":64}'
    vt, vc, vn = H.wire_parse_verify('{"request":' + req + ',"response":' + body + '}')
    assert np.array_equal(vt, toks) and np.array_equal(vn, nl)
    for i in range(n):
        assert np.array_equal(vc[i, : nl[i]]["token"], top[i, : nl[i]]["token"])
        assert vc[i, : nl[i]]["logit"].tobytes() == top[i, : nl[i]]["logit"].tobytes()      # bit-identical floats across the wire
