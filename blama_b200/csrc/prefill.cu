// prefill.cu -- launchers of the multi-token prefill path (tcgen05 GEMM + helpers).
#include "prefill.hpp"
#include "prefill_gemm.cuh"

#include <mutex>

namespace blk {
namespace {

using EncodeTiledFn = CUresult (*)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                   const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                   CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_tiled() {
    static EncodeTiledFn fn = nullptr;
    static std::once_flag once;
    std::call_once(once, [] {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
        else (void)cudaGetLastError();
    });
    return fn;
}

__global__ void f32_to_bf16_kernel(const float* __restrict__ x, __nv_bfloat16* __restrict__ y, size_t n) {
    const size_t i = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) * 4;
    if (i + 3 < n) {
        const float4 v = *reinterpret_cast<const float4*>(x + i);
        __nv_bfloat162 a = __floats2bfloat162_rn(v.x, v.y), b = __floats2bfloat162_rn(v.z, v.w);
        uint2 o; o.x = *reinterpret_cast<uint32_t*>(&a); o.y = *reinterpret_cast<uint32_t*>(&b);
        *reinterpret_cast<uint2*>(y + i) = o;
    } else {
        for (size_t j = i; j < n; j++) y[j] = __float2bfloat16_rn(x[j]);
    }
}

} // namespace

cudaError_t convert_f32_to_bf16(const float* x, __nv_bfloat16* y, size_t n, cudaStream_t st) {
    const size_t threads = (n + 3) / 4;
    f32_to_bf16_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, st>>>(x, y, n);
    return cudaGetLastError();
}

cudaError_t prefill_gemm(const QMat& W, const __nv_bfloat16* X, int T, float* C, long long ldc, const float* bias, int mode, cudaStream_t st) {
    EncodeTiledFn enc = encode_tiled();
    if (!enc) return cudaErrorNotSupported;
    if (W.K % PG_BK || T <= 0) return cudaErrorInvalidValue;
    CUtensorMap tmap;
    const cuuint64_t gdim[2] = {(cuuint64_t)W.K, (cuuint64_t)T};
    const cuuint64_t gstride[1] = {(cuuint64_t)W.K * sizeof(__nv_bfloat16)};
    const cuuint32_t box[2] = {(cuuint32_t)PG_BK, 128u};
    const cuuint32_t estr[2] = {1u, 1u};
    const CUresult r = enc(&tmap, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<__nv_bfloat16*>(X), gdim, gstride, box, estr,
                           CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                           CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return cudaErrorInvalidValue;
    void (*kernel)(const CUtensorMap, const PrefillGemmArgs) = nullptr;
    switch (W.type) {
        case QT_Q4_K: kernel = prefill_gemm_kernel<QT_Q4_K>; break;
        case QT_Q5_K: kernel = prefill_gemm_kernel<QT_Q5_K>; break;
        case QT_Q6_K: kernel = prefill_gemm_kernel<QT_Q6_K>; break;
        case QT_Q8_0: kernel = prefill_gemm_kernel<QT_Q8_0>; break;
        case QT_F32: kernel = prefill_gemm_kernel<QT_F32>; break;
        case QT_F16: kernel = prefill_gemm_kernel<QT_F16>; break;
        default: return cudaErrorInvalidValue;
    }
    int dev = 0, sms = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, PG_SMEM_BYTES);
    if (e != cudaSuccess) return e;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    PrefillGemmArgs a{};
    a.W = W; a.bias = bias; a.C = C; a.ldc = ldc; a.T = T; a.N = W.N; a.K = W.K; a.mode = mode;
    const int tiles = ((T + PG_BM - 1) / PG_BM) * ((W.N + PG_BN - 1) / PG_BN);
    const int grid = tiles < sms ? tiles : sms;
    kernel<<<grid, PG_THREADS, PG_SMEM_BYTES, st>>>(tmap, a);
    return cudaGetLastError();
}

} // namespace blk
