// gemv_ring.cuh -- persistent, TMA-fed decode mat-vec (the production decode path; gemv_kernels.cuh is the fallback for
// shapes whose planes cannot be bulk-copied, i.e. the tiny test models).
//
// One CTA per SM, 8 warps.  Each warp owns a private ring of shared-memory slots that its lane 0 keeps filled with
// cp.async.bulk copies (the 1-D TMA path: global -> shared, completion on an mbarrier), one "item" per slot:
//   item = one K-slice of one ROW PAIR = the slice's bytes of every plane of both rows (qs+hdr | ql+qh+sc | qs+d).
// A warp re-arms a slot the moment it has consumed it, so every SM always has 8 x NS slices in flight and HBM keeps
// streaming; nothing is staged through registers.  The first NS items are requested BEFORE griddepcontrol.wait, i.e.
// while the previous kernel of the step is still running (weights do not depend on it).
//
// Prologue (after the wait, once per CTA): the f32 input vector is turned into the activation format the weight type
// needs, in shared memory: optional RMSNorm * weight (ggml rms_norm + mul), then Q8_K / Q8_0 quantisation -- the
// arithmetic of ggml's quantize_row_q8_K / q8_0.  Doing it per CTA (148 x 16 KB of L2 reads) removes every separate
// norm / quantise kernel from the token's critical path.
//
// Epilogues are the ones of gemv_kernels.cuh: store (+ chunk max for the top-k), residual add, QKV (+bias, RoPE,
// KV-page write in f16), SwiGLU.
#pragma once
#include "gemv_kernels.cuh"

namespace blk {

constexpr int RING_WARPS = 8;
constexpr int RING_THREADS = RING_WARPS * 32;
constexpr int RING_MAX_SLOTS = 4;

// ---- PTX wrappers --------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE_%=;\n\t"
        "bra WAIT_%=;\n\t"
        "DONE_%=:\n\t}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
// 1-D bulk copy global -> shared (TMA engine), size a multiple of 16 B, both addresses 16 B aligned
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

// ---- plan -----------------------------------------------------------------------------------------------------------
struct RingPlan {
    int S = 1;            // K-slices per row pair (1, 2, 4, 8); a slice is a whole number of 256-element super-blocks
    int NS = 2;           // ring slots per warp
    int slot_bytes = 0;   // bytes of the largest item
    int act_bytes = 0;    // shared memory of the prepared activations
    int smem_bytes = 0;
    int ctas = 0;
    bool ok = false;
};

// bytes of one row of `nsb` super-blocks (256 elements) in the ring, per type
__host__ __device__ inline int ring_row_bytes(int type, int nsb) {
    switch (type) {
        case QT_Q4_K: return nsb * (128 + 16);
        case QT_Q5_K: return nsb * (128 + 16 + 32);
        case QT_Q6_K: return nsb * (128 + 64 + 16);      // d (2 B / super-block) is read straight from global
        case QT_Q8_0: return nsb * (256 + 16);
        default: return 0;
    }
}
__host__ __device__ inline int ring_act_bytes(int K, int fmt) {
    if (fmt == ACT_Q8_K) return K + ((K >> 8) * 4 + 15) / 16 * 16 + (K >> 4) * 2;
    if (fmt == ACT_Q8_0) return K + (K >> 5) * 4;
    return K * 4;
}

inline RingPlan ring_plan(int K, int type_a, int type_b, int total_pairs, int n_sms, int smem_budget = 108 * 1024) {
    RingPlan p;
    if (K % 256 || total_pairs <= 0) return p;
    if (ring_row_bytes(type_a, 1) == 0 || ring_row_bytes(type_b, 1) == 0) return p;
    const int nsb = K / 256;
    const int fmt = act_format_for(type_a);
    p.act_bytes = (ring_act_bytes(K, fmt) + 127) / 128 * 128;
    const int ring_budget = smem_budget - p.act_bytes - 1024;
    if (ring_budget < 16 * 1024) return p;
    const int per_warp = ring_budget / RING_WARPS;
    for (int S = 1; S <= 8; S *= 2) {
        if (S > nsb) break;
        const int nsb_slice = (nsb + S - 1) / S;
        const int item = 2 * std::max(ring_row_bytes(type_a, nsb_slice), ring_row_bytes(type_b, nsb_slice));
        const int slot = (item + 127) / 128 * 128;
        int ns = per_warp / slot;
        if (ns > RING_MAX_SLOTS) ns = RING_MAX_SLOTS;
        if (ns >= 2 && slot <= 6 * 1024) { p.S = S; p.NS = ns; p.slot_bytes = slot; p.ok = true; break; }
    }
    if (!p.ok) return p;
    p.smem_bytes = p.act_bytes + RING_WARPS * p.NS * p.slot_bytes + 1024;
    const int G = RING_WARPS / p.S;
    p.ctas = std::min(n_sms, (total_pairs + G - 1) / G);
    return p;
}

// ---- a unit read from the ring instead of from global -----------------------------------------------------------------
// row_base: start of this row's slice in the slot; nsb: super-blocks in the slice; ul: unit index inside the slice
template <int TYPE> struct RingUnit;
template <> struct RingUnit<QT_Q4_K> : RowUnit<QT_Q4_K> {
    __device__ __forceinline__ void load(const uint8_t* rb, int nsb, int ul, const QMat&, int64_t, int) {
        const uint4* q = reinterpret_cast<const uint4*>(rb + ul * 32);
        q0 = q[0]; q1 = q[1];
        hdr = *reinterpret_cast<const uint4*>(rb + nsb * 128 + (ul >> 2) * 16);
    }
};
template <> struct RingUnit<QT_Q5_K> : RowUnit<QT_Q5_K> {
    __device__ __forceinline__ void load(const uint8_t* rb, int nsb, int ul, const QMat&, int64_t, int) {
        const uint4* q = reinterpret_cast<const uint4*>(rb + ul * 32);
        q0 = q[0]; q1 = q[1];
        hdr = *reinterpret_cast<const uint4*>(rb + nsb * 128 + (ul >> 2) * 16);
        const uint4* h = reinterpret_cast<const uint4*>(rb + nsb * 144 + (ul >> 2) * 32);
        h0 = h[0]; h1 = h[1];
    }
};
template <> struct RingUnit<QT_Q6_K> : RowUnit<QT_Q6_K> {
    __device__ __forceinline__ void load(const uint8_t* rb, int nsb, int ul, const QMat& W, int64_t row, int u_global) {
        const int s = ul >> 2, hh = (ul >> 1) & 1, t = ul & 1;
        const uint8_t* ql = rb + s * 128 + hh * 64 + t * 16;
        l0 = *reinterpret_cast<const uint4*>(ql); l1 = *reinterpret_cast<const uint4*>(ql + 32);
        h = *reinterpret_cast<const uint4*>(rb + nsb * 128 + s * 64 + hh * 32 + t * 16);
        sc = *reinterpret_cast<const uint2*>(rb + nsb * 192 + s * 16 + hh * 8);
        dh = __ldg(reinterpret_cast<const uint16_t*>(W.p3) + (size_t)row * (W.K >> 8) + (u_global >> 2));
    }
};
template <> struct RingUnit<QT_Q8_0> : RowUnit<QT_Q8_0> {
    __device__ __forceinline__ void load(const uint8_t* rb, int nsb, int ul, const QMat&, int64_t, int) {
        const uint4* q = reinterpret_cast<const uint4*>(rb + ul * 32);
        q0 = q[0]; q1 = q[1];
        dh = *reinterpret_cast<const uint16_t*>(rb + nsb * 256 + ul * 2);
    }
};

// lane 0: request the bytes of super-blocks [sb0, sb0+nsb) of `row` into dst (layout = RingUnit's expectations)
template <int TYPE>
__device__ __forceinline__ void ring_issue_row(uint8_t* dst, const QMat& W, int64_t row, int sb0, int nsb, uint64_t* bar) {
    const size_t rsb = (size_t)row * (W.K >> 8) + sb0;       // first super-block of the slice, in super-block units
    if (TYPE == QT_Q4_K) {
        bulk_g2s(dst, W.p0 + rsb * 128, nsb * 128, bar);
        bulk_g2s(dst + nsb * 128, W.p1 + rsb * 16, nsb * 16, bar);
    } else if (TYPE == QT_Q5_K) {
        bulk_g2s(dst, W.p0 + rsb * 128, nsb * 128, bar);
        bulk_g2s(dst + nsb * 128, W.p1 + rsb * 16, nsb * 16, bar);
        bulk_g2s(dst + nsb * 144, W.p2 + rsb * 32, nsb * 32, bar);
    } else if (TYPE == QT_Q6_K) {
        bulk_g2s(dst, W.p0 + rsb * 128, nsb * 128, bar);
        bulk_g2s(dst + nsb * 128, W.p1 + rsb * 64, nsb * 64, bar);
        bulk_g2s(dst + nsb * 192, W.p2 + rsb * 16, nsb * 16, bar);
    } else {   // Q8_0: 8 blocks of 32 per "super-block"
        bulk_g2s(dst, W.p0 + rsb * 256, nsb * 256, bar);
        bulk_g2s(dst + nsb * 256, W.p1 + rsb * 16, nsb * 16, bar);
    }
}

struct RingArgs {
    GemvArgs g;                 // segments, epilogue parameters, output (g.act is unused: activations are made in the prologue)
    const float* in;            // f32 input vector [K]
    const float* norm_w;        // RMSNorm weight, or nullptr for "quantise only"
    float eps;
    RingPlan plan;
};

// which rows a pair index means (see gemv_kernels.cuh): returns segment index
template <int EPI>
__device__ __forceinline__ int ring_pair_rows(const GemvArgs& a, int pair, int& r0, int& r1) {
    if (EPI == EPI_SWIGLU) { r0 = r1 = pair; return 0; }
    int si = 0;
    if (a.nseg > 1 && pair >= a.seg[1].pair0) si = 1;
    if (a.nseg > 2 && pair >= a.seg[2].pair0) si = 2;
    const int p = pair - a.seg[si].pair0;
    r0 = 2 * p; r1 = 2 * p + 1;
    if (EPI == EPI_QKV && a.neox && a.seg[si].kind != 2) {
        const int hd = a.d_head >> 1;
        r0 = (p / hd) * a.d_head + (p % hd); r1 = r0 + hd;
    }
    return si;
}

template <int EPI, int TA, int TB>
__global__ void __launch_bounds__(RING_THREADS, 1) gemv_ring_kernel(const RingArgs ra) {
    extern __shared__ __align__(128) unsigned char ring_smem[];
    const GemvArgs& a = ra.g;
    const RingPlan& pl = ra.plan;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int S = pl.S, NS = pl.NS, G = RING_WARPS / S;
    const int g = w / S, ks = w % S;
    const int K = a.seg[0].W.K, nsb = K >> 8;
    const int fmt = act_format_for(TA);

    // shared memory map: [activations][ring: warp-major, slot-major][barriers][partials]
    unsigned char* act_base = ring_smem;
    uint8_t* my_ring = ring_smem + pl.act_bytes + (size_t)w * NS * pl.slot_bytes;
    uint64_t* bars = reinterpret_cast<uint64_t*>(ring_smem + pl.act_bytes + (size_t)RING_WARPS * NS * pl.slot_bytes);
    uint64_t* my_bar = bars + w * RING_MAX_SLOTS;
    float* part = reinterpret_cast<float*>(bars + RING_WARPS * RING_MAX_SLOTS);     // [2 parities][RING_WARPS][2]

    // this CTA's pairs, this warp group's share, this warp's slice
    const int p_begin = (int)((long long)a.total_pairs * blockIdx.x / gridDim.x);
    const int p_end = (int)((long long)a.total_pairs * (blockIdx.x + 1) / gridDim.x);
    const int n_items = (p_end - p_begin > g) ? (p_end - p_begin - g + G - 1) / G : 0;      // pairs p_begin + g + i*G
    const int sb0 = (int)((long long)nsb * ks / S), sb1 = (int)((long long)nsb * (ks + 1) / S);
    const int nsb_s = sb1 - sb0;

    auto issue = [&](int item) {          // lane 0 only
        const int pair = p_begin + g + item * G;
        int r0, r1;
        const int si = ring_pair_rows<EPI>(a, pair, r0, r1);
        const QMat& Wa = a.seg[si].W;
        const QMat& Wb = (EPI == EPI_SWIGLU) ? a.seg[1].W : Wa;
        uint8_t* dst = my_ring + (size_t)(item % NS) * pl.slot_bytes;
        uint64_t* bar = my_bar + (item % NS);
        if (TA != TB && si == 2) {
            const int rb = ring_row_bytes(TB, nsb_s);
            mbar_expect_tx(bar, 2 * rb);
            ring_issue_row<TB>(dst, Wa, r0, sb0, nsb_s, bar);
            ring_issue_row<TB>(dst + rb, Wb, r1, sb0, nsb_s, bar);
        } else {
            const int rb = ring_row_bytes(TA, nsb_s);
            mbar_expect_tx(bar, 2 * rb);
            ring_issue_row<TA>(dst, Wa, r0, sb0, nsb_s, bar);
            ring_issue_row<TA>(dst + rb, Wb, r1, sb0, nsb_s, bar);
        }
    };

    if (lane == 0) {
        for (int s = 0; s < NS; s++) mbar_init(my_bar + s, 1);
        mbar_fence_init();
    }
    __syncwarp();
    pdl_launch_dependents();
    // weights do not depend on the previous kernel: request the first NS items now
    if (lane == 0 && nsb_s > 0) for (int i = 0; i < NS && i < n_items; i++) issue(i);

    pdl_wait();

    // ---- prologue: f32 input -> (RMSNorm) -> activations in shared memory ----------------------------------------------
    int8_t* sq = reinterpret_cast<int8_t*>(act_base);
    float* sd = reinterpret_cast<float*>(act_base + K);
    int16_t* sbs = reinterpret_cast<int16_t*>(act_base + K + (((K >> 8) * 4 + 15) / 16) * 16);
    float* sf = reinterpret_cast<float*>(act_base);
    {
        __shared__ double s_red[RING_WARPS];
        __shared__ float s_scale;
        float scale = 1.0f;
        if (ra.norm_w) {
            double sum = 0.0;
            for (int i = threadIdx.x; i < K; i += RING_THREADS) { const float v = ra.in[i]; sum += (double)__fmul_rn(v, v); }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
            if (lane == 0) s_red[w] = sum;
            __syncthreads();
            if (threadIdx.x == 0) {
                double tot = 0.0;
                for (int i = 0; i < RING_WARPS; i++) tot += s_red[i];
                const float mean = (float)(tot / (double)K);
                s_scale = __fdiv_rn(1.0f, __fsqrt_rn(__fadd_rn(mean, ra.eps)));
            }
            __syncthreads();
            scale = s_scale;
        }
        if (fmt == ACT_F32) {
            for (int i = threadIdx.x; i < K; i += RING_THREADS) {
                float v = ra.in[i];
                if (ra.norm_w) v = __fmul_rn(__fmul_rn(v, scale), ra.norm_w[i]);
                sf[i] = v;
            }
        } else {
            for (int b = w; b < nsb; b += RING_WARPS) {
                if (ra.norm_w) quantize_256_warp<true, false>(ra.in + (size_t)b * 256, 256, fmt, sq + (size_t)b * 256, sd, sbs, b, scale, ra.norm_w + (size_t)b * 256);
                else quantize_256_warp<false, false>(ra.in + (size_t)b * 256, 256, fmt, sq + (size_t)b * 256, sd, sbs, b);
            }
        }
        __syncthreads();
    }
    const ActView A{sq, sd, sbs, sf};

    // ---- main loop -------------------------------------------------------------------------------------------------------
    for (int item = 0; item < n_items; item++) {
        const int pair = p_begin + g + item * G;
        int r0, r1;
        const int si = ring_pair_rows<EPI>(a, pair, r0, r1);
        const GemvSeg& Sg = a.seg[si];
        float v0 = 0.0f, v1 = 0.0f;
        if (nsb_s > 0) {
            const uint8_t* slot = my_ring + (size_t)(item % NS) * pl.slot_bytes;
            mbar_wait(my_bar + (item % NS), (uint32_t)((item / NS) & 1));
            const int units = nsb_s * ((TA == QT_Q8_0) ? 8 : 4);
            const int u_first = sb0 * ((TA == QT_Q8_0) ? 8 : 4);
            if (TA != TB && si == 2) {
                const int rb = ring_row_bytes(TB, nsb_s);
                for (int ul = lane; ul < units; ul += 32) {
                    RingUnit<TB> ua, ub;
                    ua.load(slot, nsb_s, ul, Sg.W, r0, u_first + ul); ub.load(slot + rb, nsb_s, ul, Sg.W, r1, u_first + ul);
                    v0 += ua.dot(A, u_first + ul); v1 += ub.dot(A, u_first + ul);
                }
            } else {
                const int rb = ring_row_bytes(TA, nsb_s);
                const QMat& Wb = (EPI == EPI_SWIGLU) ? a.seg[1].W : Sg.W;
                for (int ul = lane; ul < units; ul += 32) {
                    RingUnit<TA> ua, ub;
                    ua.load(slot, nsb_s, ul, Sg.W, r0, u_first + ul); ub.load(slot + rb, nsb_s, ul, Wb, r1, u_first + ul);
                    v0 += ua.dot(A, u_first + ul); v1 += ub.dot(A, u_first + ul);
                }
            }
            __syncwarp();
            if (lane == 0 && item + NS < n_items) issue(item + NS);      // slot consumed by every lane: refill it
            v0 = warp_sum(v0); v1 = warp_sum(v1);
        }
        if (S > 1) {
            float* pp = part + ((item & 1) * RING_WARPS + w) * 2;
            if (lane == 0) { pp[0] = v0; pp[1] = v1; }
            // the S warps of a group meet on a named barrier (ids 1..G); groups run independently of each other
            asm volatile("bar.sync %0, %1;" ::"r"(g + 1), "r"(S * 32) : "memory");
            if (ks != 0) continue;
            v0 = pp[0]; v1 = pp[1];
            for (int k = 1; k < S; k++) { v0 += pp[2 * k]; v1 += pp[2 * k + 1]; }
        }
        if (lane != 0) continue;
        // ---- epilogue (identical to gemv_kernels.cuh) ----
        if (EPI == EPI_SWIGLU) { a.out[pair] = (v0 / (1.0f + expf(-v0))) * v1; continue; }
        if (Sg.bias) { v0 += Sg.bias[r0]; v1 += Sg.bias[r1]; }
        if (EPI == EPI_STORE) {
            a.out[r0] = v0; a.out[r1] = v1;
            if (a.tail.kind == TAIL_CHUNKMAX) atomicMax(a.tail.chunk_max + (r0 >> a.tail.chunk_shift), float_order_key(fmaxf(v0, v1)));
        } else if (EPI == EPI_RESID) {
            a.out[r0] += v0; a.out[r1] += v1;
        } else if (EPI == EPI_QKV) {
            if (Sg.kind != 2) {
                const int i = a.neox ? (r0 % a.d_head) : ((r0 % a.d_head) >> 1);
                const float2 cs = a.rope_cs[i];
                const float x0 = v0, x1 = v1;
                v0 = x0 * cs.x - x1 * cs.y;
                v1 = x0 * cs.y + x1 * cs.x;
            }
            if (Sg.kind == 0) { a.out[r0] = v0; a.out[r1] = v1; }
            else {
                const int pos = a.pos[0];
                const size_t base = ((size_t)a.page_table[pos / KV_PAGE] * KV_PAGE + (pos % KV_PAGE)) * a.kv_dim;
                __half* dst = (Sg.kind == 1) ? a.k_pool : a.v_pool;
                dst[base + r0] = __float2half_rn(v0);
                dst[base + r1] = __float2half_rn(v1);
            }
        }
    }
}

// implemented in ring_*.cu; cudaErrorInvalidValue when no instantiation exists for the type pair
cudaError_t launch_ring_store(const RingArgs& a, cudaStream_t st);
cudaError_t launch_ring_resid(const RingArgs& a, cudaStream_t st);
cudaError_t launch_ring_qkv(const RingArgs& a, cudaStream_t st);
cudaError_t launch_ring_swiglu(const RingArgs& a, cudaStream_t st);

} // namespace blk
