// gemv_swiglu.cu -- instantiations of the decode mat-vec for the EPI_SWIGLU epilogue (see decode_kernels.cuh).
#include "gemv_kernels.cuh"

namespace blk {
namespace {
template <int TA, int TB, int U>
cudaError_t go(const GemvArgs& a, const GemvPlan& p, cudaStream_t st) {
    return launch_pdl(gemv_pairs_kernel<EPI_SWIGLU, TA, TB, U>, dim3(p.ctas), dim3(GEMV_THREADS), 0, st, a, p.S);
}
template <int TA, int TB>
cudaError_t by_u(const GemvArgs& a, const GemvPlan& p, cudaStream_t st) {
    switch (p.U) {
        case 1: return go<TA, TB, 1>(a, p, st);
        case 2: return go<TA, TB, 2>(a, p, st);
        default: return cudaErrorInvalidValue;
    }
}
} // namespace

cudaError_t launch_gemv_swiglu(const GemvArgs& a, cudaStream_t st) {
    const int ta = a.seg[0].W.type, tb = (a.nseg > 2) ? a.seg[2].W.type : ta;
    if (a.nseg > 1 && a.seg[1].W.type != ta) return cudaErrorInvalidValue;
    const GemvPlan p = gemv_plan(a.seg[0].W.K, ta, a.total_pairs);
    if (p.U == 0 || p.ctas <= 0) return cudaErrorInvalidValue;
#define BLK_G(A, B) if (ta == A && tb == B) return by_u<A, B>(a, p, st);
    BLK_G(QT_Q4_K, QT_Q4_K) BLK_G(QT_Q6_K, QT_Q6_K) BLK_G(QT_Q8_0, QT_Q8_0) BLK_G(QT_F32, QT_F32) BLK_G(QT_Q5_K, QT_Q5_K)
#undef BLK_G
    return cudaErrorInvalidValue;
}
} // namespace blk
