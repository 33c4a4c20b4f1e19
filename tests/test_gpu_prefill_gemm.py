"""tcgen05 prefill GEMM (dequant-fused, bf16 operands, f32 TMEM accumulation) against a float64 reference on the same
bf16-rounded operands: what remains is f32 accumulation noise, so the tolerance is tight and any layout / descriptor
mistake shows up as garbage."""
import numpy as np
import pytest

from blama_b200 import gguf_synth as gs

pytestmark = pytest.mark.gpu


def bf16_round(x: np.ndarray) -> np.ndarray:
    u = np.ascontiguousarray(x, dtype=np.float32).view(np.uint32).astype(np.uint64)
    u = (u + 0x7FFF + ((u >> 16) & 1)) & 0xFFFF0000
    return u.astype(np.uint32).view(np.float32)


CASES = [(256, 256, 256), (300, 384, 1024), (1, 256, 512), (77, 1000, 256), (512, 1024, 4096), (640, 512, 14336)]


@pytest.mark.parametrize("gtype,name", [(gs.Q4_K, "Q4_K"), (gs.Q6_K, "Q6_K"), (gs.Q8_0, "Q8_0"), (gs.Q5_K, "Q5_K"), (gs.F32, "F32")])
@pytest.mark.parametrize("T,N,K", CASES)
@pytest.mark.parametrize("form", ["fused", "panel"])
def test_prefill_gemm(gtype, name, T, N, K, form, oracle, monkeypatch):
    """form = fused: weights dequantised inside the GEMM; panel: dequantised once to a bf16 panel, both operands TMA-fed."""
    from blama_b200 import capi

    if form == "panel":
        if gtype == gs.F32:
            pytest.skip("the panel form covers the quantised types")
        monkeypatch.setenv("BLK_TEST_PANEL", "1")

    if gtype == gs.F32 and K > 4096:
        pytest.skip("F32 weights are a test-only format")
    rng = np.random.default_rng(T * 3 + N + K)
    blk = gs.random_blocks(rng, gtype, N * K, 1.0 / np.sqrt(K))
    x = rng.standard_normal((T, K)).astype(np.float32)
    w = bf16_round(oracle.dequantize(gtype, blk, N * K).reshape(N, K)).astype(np.float64)
    want = bf16_round(x).astype(np.float64) @ w.T
    got = capi.test_gemm(gtype, blk, N, K, x)
    err = np.abs(got - want).max()
    assert err <= 2e-4 * (np.abs(want).max() + 1e-6), (err, np.abs(want).max())
