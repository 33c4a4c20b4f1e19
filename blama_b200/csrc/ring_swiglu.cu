// ring_swiglu.cu -- instantiations of the TMA-ring decode mat-vec for the EPI_SWIGLU epilogue (see gemv_ring.cuh).
#include "gemv_ring.cuh"

namespace blk {
namespace {
template <int TA, int TB>
cudaError_t go(const RingArgs& a, cudaStream_t st) {
    static unsigned long long attr_done = 0;      // bit per device
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    if (!(attr_done >> (dev & 63) & 1ull)) {
        e = cudaFuncSetAttribute(gemv_ring_kernel<EPI_SWIGLU, TA, TB>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
        if (e != cudaSuccess) return e;
        attr_done |= 1ull << (dev & 63);
    }
    return launch_pdl(gemv_ring_kernel<EPI_SWIGLU, TA, TB>, dim3(a.plan.ctas), dim3(RING_THREADS), (size_t)a.plan.smem_bytes, st, a);
}
} // namespace

cudaError_t launch_ring_swiglu(const RingArgs& a, cudaStream_t st) {
    const GemvArgs& g = a.g;
    const int ta = g.seg[0].W.type, tb = (g.nseg > 2) ? g.seg[2].W.type : ta;
    if (g.nseg > 1 && g.seg[1].W.type != ta) return cudaErrorInvalidValue;
    if (!a.plan.ok) return cudaErrorInvalidValue;
#define BLK_G(A, B) if (ta == A && tb == B) return go<A, B>(a, st);
    BLK_G(QT_Q4_K, QT_Q4_K) BLK_G(QT_Q6_K, QT_Q6_K) BLK_G(QT_Q8_0, QT_Q8_0) BLK_G(QT_Q5_K, QT_Q5_K)
#undef BLK_G
    return cudaErrorInvalidValue;
}
} // namespace blk
