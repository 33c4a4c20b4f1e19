// prefill.hpp -- host-side entry points of the multi-token prefill path (implemented in prefill.cu).
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include "qweights.cuh"

namespace blk {

// C[T][N] (f32, row stride ldc) = or += X[T][K] (bf16, row-major, device) . W^T  through the tcgen05 GEMM
// mode: 0 store (+bias), 1 accumulate into C.  Returns a CUDA error (cudaErrorNotSupported if the driver lacks TMA).
cudaError_t prefill_gemm(const QMat& W, const __nv_bfloat16* X, int T, float* C, long long ldc, const float* bias, int mode, cudaStream_t st);

// y[i] = bf16(x[i])
cudaError_t convert_f32_to_bf16(const float* x, __nv_bfloat16* y, size_t n, cudaStream_t st);

} // namespace blk
