"""tcgen05 prefill GEMM throughput on the 8B shapes (T = 2048)."""
import sys
sys.path.insert(0, '.')
import numpy as np
from blama_b200 import capi, gguf_synth as gs
rng = np.random.default_rng(0)
T = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
for name, gtype, N, K in [("wq Q4_K", gs.Q4_K, 4096, 4096), ("gate Q4_K", gs.Q4_K, 14336, 4096), ("down Q4_K", gs.Q4_K, 4096, 14336), ("down Q6_K", gs.Q6_K, 4096, 14336), ("lm_head Q6_K", gs.Q6_K, 32768, 4096), ("wq Q8_0", gs.Q8_0, 4096, 4096)]:
    blk = gs.random_blocks(rng, gtype, N * K, 0.02)
    ms = capi.bench_gemm(gtype, blk, N, K, T, 10)
    print(f"{name:14s} T={T} N={N} K={K}: {ms*1e3:8.1f} us  {2*T*N*K/ms/1e9:8.1f} TFLOP/s ({2*T*N*K/ms/1e9/1610.3*100:.1f}% of 1610)")
