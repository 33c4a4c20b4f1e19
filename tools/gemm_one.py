"""one tcgen05 prefill GEMM shape, for ncu captures"""
import sys
sys.path.insert(0, '.')
import numpy as np
from blama_b200 import capi, gguf_synth as gs
rng = np.random.default_rng(0)
T, N, K = 2048, 14336, 4096
gtype = {"q4k": gs.Q4_K, "q6k": gs.Q6_K, "q80": gs.Q8_0}[sys.argv[1] if len(sys.argv) > 1 else "q4k"]
blk = gs.random_blocks(rng, gtype, N * K, 0.02)
ms = capi.bench_gemm(gtype, blk, N, K, T, 2)
print(f"{ms*1e3:.1f} us  {2*T*N*K/ms/1e9:.1f} TFLOP/s")
