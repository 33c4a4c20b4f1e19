"""LogitComparer / MetricsAggregator: product (host C++) and oracle restatement against
  (1) the reference's own unit test   inference/test/t-LogitComparer.cpp:13-39,
  (2) golden vectors produced by the reference's own object code (tests/golden/logit_comparer_golden.json,
      generator tools/gen_logit_comparer_golden.py), bit for bit,
  (3) the reference library itself where it was built (oracle/_ref; absent on the GPU box)."""
import json
import os

import numpy as np
import pytest

GOLDEN = os.path.join(os.path.dirname(__file__), "golden", "logit_comparer_golden.json")


def _hex(x) -> str:
    return np.float32(x).tobytes().hex()


@pytest.fixture(scope="module")
def golden():
    with open(GOLDEN) as f:
        return json.load(f)


@pytest.fixture(scope="module")
def host():
    from blama_b200 import host_api

    host_api.lib()
    return host_api


def test_reference_unit_case(oracle, host):
    a = [(i, 17.5 - 0.5 * i) for i in range(10)]
    for impl in (oracle, host):
        assert impl.lc_similarity(a, a) == 1.0
        top1, dist, jsd = impl.lc_compare(a, a)
        assert (top1, dist, jsd) == (1.0, 0.0, 0.0)
        assert impl.lc_score([(top1, dist, jsd)]) == 1.0


@pytest.mark.parametrize("which", ["oracle", "product"])
def test_golden_vectors_bit_exact(which, golden, oracle, host):
    impl = oracle if which == "oracle" else host
    finite = []
    for c in golden["cases"]:
        a = [tuple(x) for x in c["a"]]
        b = [tuple(x) for x in c["b"]]
        m = impl.lc_compare(a, b)
        assert [_hex(x) for x in m] == c["metrics_hex"], c["name"]
        assert _hex(impl.lc_similarity(a, b)) == c["similarity_hex"], c["name"]
        if all(np.isfinite(m)):
            finite.append(m)
    for n, want in zip(golden["score_prefix_lengths"], golden["score_hex"]):
        assert _hex(impl.lc_score(finite[:n])) == want


def test_against_reference_library_when_present(oracle, host):
    if oracle.ref_lib() is None:
        pytest.skip("oracle/_ref not built here (no /root/reference): golden vectors cover this")
    rng = np.random.default_rng(7)
    ms = []
    for _ in range(3000):
        n1, n2 = int(rng.integers(1, 11)), int(rng.integers(1, 11))
        ids1 = rng.choice(40, n1, replace=False)
        ids2 = rng.choice(40, n2, replace=False)
        a = list(zip(ids1.tolist(), np.sort(rng.normal(5, 3, n1).astype(np.float32))[::-1].tolist()))
        b = list(zip(ids2.tolist(), np.sort(rng.normal(5, 3, n2).astype(np.float32))[::-1].tolist()))
        want = oracle.ref_compare(a, b)
        for impl in (oracle, host):
            got = impl.lc_compare(a, b)
            assert [_hex(x) for x in got] == [_hex(x) for x in want]
            assert _hex(impl.lc_similarity(a, b)) == _hex(oracle.ref_similarity(a, b))
        if all(np.isfinite(want)):
            ms.append(want)
    assert _hex(host.lc_score(ms)) == _hex(oracle.ref_score(ms)) == _hex(oracle.lc_score(ms))


def test_edge_cases(oracle, host):
    # single entry, disjoint ids (jsd over an empty intersection = 0), unequal lengths
    for a, b in [([(5, 3.0)], [(5, 3.0)]), ([(1, 2.0), (2, 1.0)], [(3, 2.0), (4, 1.0)]), ([(1, 9.0), (2, 8.0), (3, 1.0)], [(1, 9.5)])]:
        assert [_hex(x) for x in oracle.lc_compare(a, b)] == [_hex(x) for x in host.lc_compare(a, b)]
    top1, dist, jsd = host.lc_compare([(1, 2.0), (2, 1.0)], [(3, 2.0), (4, 1.0)])
    assert top1 == 0.0 and dist == 0.0 and jsd == 0.0
