"""GPU parity of the decode kernels against the CPU oracle, through the C ABI (blk_test_*)."""
import numpy as np
import pytest

from blama_b200 import gguf_synth as gs

pytestmark = pytest.mark.gpu

TYPES = [(gs.Q4_K, "Q4_K"), (gs.Q5_K, "Q5_K"), (gs.Q6_K, "Q6_K"), (gs.Q8_0, "Q8_0"), (gs.F32, "F32")]


@pytest.mark.parametrize("gtype,name", TYPES)
def test_retile_dequant_bit_exact(gtype, name, oracle):
    """device re-tiling keeps every bit: dequantised values equal ggml's dequantize_row_* bit for bit"""
    from blama_b200 import capi

    rng = np.random.default_rng(11)
    rows, k = 6, 1024
    blk = gs.random_blocks(rng, gtype, rows * k, 0.05)
    want = oracle.dequantize(gtype, blk, rows * k).reshape(rows, k)
    got = capi.test_dequant(gtype, blk, rows, k)
    assert np.array_equal(want.view(np.uint32), got.view(np.uint32))


@pytest.mark.parametrize("gtype,name", TYPES)
@pytest.mark.parametrize("rows,k", [(2, 256), (64, 1024), (130, 4096), (34, 14336), (6, 28672), (10, 18944)])
def test_gemv_matches_ggml_arithmetic(gtype, name, rows, k, oracle):
    """dequant-fused GEMV == oracle's ggml-cpu restatement (Q8_K/Q8_0 activations, integer dot): the integer partial sums
    are identical, so the only difference is fp32 summation order -> tight tolerance"""
    from blama_b200 import capi

    if gtype == gs.Q8_0 and k > 18944:
        pytest.skip("Q8_0 rows longer than 18944 (Qwen2.5-7B ffn) are not covered")
    if gtype == gs.F32 and k > 8192:
        pytest.skip("F32 weights are a test-only format: rows up to 8192 elements")
    rng = np.random.default_rng(rows * 7 + k)
    blk = gs.random_blocks(rng, gtype, rows * k, 1.0 / np.sqrt(k))
    x = rng.standard_normal(k).astype(np.float32)
    x[rng.integers(0, k, 5)] *= 8.0          # outliers exercise the per-block scales
    want = oracle.matvec(gtype, blk, rows, k, x, oracle.MODE_GGML)
    got = capi.test_gemv(gtype, blk, rows, k, x)
    scale = np.abs(want).max() + 1e-6
    assert np.abs(got - want).max() <= 2e-5 * scale, (np.abs(got - want).max(), scale)


def test_gemv_zero_and_ragged_q8_0(oracle):
    from blama_b200 import capi

    rng = np.random.default_rng(5)
    # all-zero activations (Q8_K d = 0 branch) and a Q8_0 row length that is not a multiple of 256
    for gtype, k in [(gs.Q4_K, 512), (gs.Q6_K, 512), (gs.Q8_0, 512), (gs.Q8_0, 32 * 9)]:
        rows = 8
        blk = gs.random_blocks(rng, gtype, rows * k, 0.1)
        x0 = np.zeros(k, dtype=np.float32)
        assert np.array_equal(capi.test_gemv(gtype, blk, rows, k, x0), np.zeros(rows, dtype=np.float32))
        x = rng.standard_normal(k).astype(np.float32)
        want = oracle.matvec(gtype, blk, rows, k, x, oracle.MODE_GGML)
        got = capi.test_gemv(gtype, blk, rows, k, x)
        assert np.abs(got - want).max() <= 2e-5 * (np.abs(want).max() + 1e-6)
