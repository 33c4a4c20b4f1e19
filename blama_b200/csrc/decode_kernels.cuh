// decode_kernels.cuh -- batch-1 decode path (Session::getToken -> llama_decode(1 token), reference
// inference/code/llama/Session.cpp:169-190, 395-401) as hand-written sm_100a kernels.
//
// Arithmetic follows the reference's CPU backend (upstream ggml-cpu): the activation vector of every W.x is
// quantised row-wise to Q8_K (K-quant weights) or Q8_0 (Q8_0 weights) and multiplied with the still-quantised weights
// through integer dot products (dp4a) whose partial sums are scaled in fp32.  All integer partial sums are therefore
// bit-identical to ggml_vec_dot_q4_K_q8_K & co; only the order of the fp32 additions differs.
//
// All kernels here are HBM-bound byte/integer work: coalesced 128-bit loads (two per lane per step), warp-shuffle
// reductions, grids sized in multiples of the SM count.  No tensor cores (decode -> HBM roofline).
#pragma once
#include "qweights.cuh"

namespace blk {

constexpr int KV_PAGE = 64;          // tokens per KV page
constexpr int GEMV_THREADS = 256;    // 8 warps per CTA
constexpr int MAX_GQ = 8;            // query heads per KV head (70B: 8, Qwen2.5-7B: 7)
constexpr int TOPK_MAX = 64;
constexpr int TOPK_CHUNK = 1024;     // logits per CTA in the first top-k stage

// activations prepared for a quantised mat-vec
struct ActBuf {
    float* f32 = nullptr;     // [K]   (always written: residual / debug / F32 weights)
    int8_t* q = nullptr;      // [K]
    float* d = nullptr;       // [K/256] (Q8_K) or [K/32] (Q8_0; value already rounded through fp16)
    int16_t* bs = nullptr;    // [K/16]  (Q8_K only)
};

__device__ __forceinline__ uint4 ldg_stream(const void* p) {
    uint4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
    return r;
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

// =================================================================================================================
// weight re-tiling (load time): ggml blocks -> split planes (qweights.cuh)
// =================================================================================================================
__global__ void retile_q4k_kernel(const uint8_t* __restrict__ src, uint8_t* qs, uint8_t* hdr, int64_t nblk) {
    const int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= nblk) return;
    const uint8_t* s = src + b * 144;
    for (int i = 0; i < 16; i++) hdr[b * 16 + i] = s[i];
    for (int i = 0; i < 128; i++) qs[b * 128 + i] = s[16 + i];
}
__global__ void retile_q5k_kernel(const uint8_t* __restrict__ src, uint8_t* qs, uint8_t* hdr, uint8_t* qh, int64_t nblk) {
    const int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= nblk) return;
    const uint8_t* s = src + b * 176;
    for (int i = 0; i < 16; i++) hdr[b * 16 + i] = s[i];
    for (int i = 0; i < 32; i++) qh[b * 32 + i] = s[16 + i];
    for (int i = 0; i < 128; i++) qs[b * 128 + i] = s[48 + i];
}
__global__ void retile_q6k_kernel(const uint8_t* __restrict__ src, uint8_t* ql, uint8_t* qh, uint8_t* sc, uint8_t* d, int64_t nblk) {
    const int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= nblk) return;
    const uint8_t* s = src + b * 210;
    for (int i = 0; i < 128; i++) ql[b * 128 + i] = s[i];
    for (int i = 0; i < 64; i++) qh[b * 64 + i] = s[128 + i];
    for (int i = 0; i < 16; i++) sc[b * 16 + i] = s[192 + i];
    d[b * 2] = s[208]; d[b * 2 + 1] = s[209];
}
__global__ void retile_q80_kernel(const uint8_t* __restrict__ src, uint8_t* qs, uint8_t* d, int64_t nblk) {
    const int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= nblk) return;
    const uint8_t* s = src + b * 34;
    d[b * 2] = s[0]; d[b * 2 + 1] = s[1];
    for (int i = 0; i < 32; i++) qs[b * 32 + i] = s[2 + i];
}

// =================================================================================================================
// embedding gather (get_rows with dequant) + per-step RoPE table
// =================================================================================================================
// rope table: cs[i] = {cos(theta_i), sin(theta_i)}, theta_i = pos * theta_scale^i / freq_factor_i, theta built by
// repeated multiplication exactly as ggml_rope_cache_init does
__device__ inline void rope_table_fill(float2* cs, int half_rot, int pos, float theta_scale, const float* freq_factors) {
    for (int i = threadIdx.x; i < half_rot; i += blockDim.x) {
        float theta = (float)pos;
        for (int k = 0; k < i; k++) theta *= theta_scale;
        const float ff = freq_factors ? freq_factors[i] : 1.0f;
        const float th = theta / ff;
        float s, c; sincosf(th, &s, &c);
        cs[i] = make_float2(c, s);
    }
}

// grid = n_tok CTAs.  x[t] = dequant(token_embd[tok[t]]); CTA 0.. also fill the rope table rows for their token
__global__ void embed_kernel(QMat E, const int32_t* __restrict__ tokens, const int32_t* __restrict__ pos0, float* x,
                             float2* rope_cs, int half_rot, float theta_scale, const float* freq_factors) {
    const int t = blockIdx.x;
    dequant_row_cta(E, tokens[t], x + (size_t)t * E.K);
    rope_table_fill(rope_cs + (size_t)t * half_rot, half_rot, pos0[0] + t, theta_scale, freq_factors);
}

// =================================================================================================================
// activation preparation: (optional RMSNorm * weight) -> f32 copy + Q8_K / Q8_0 quantisation
//   rms_norm: upstream ggml-cpu ops.cpp ggml_compute_forward_rms_norm_f32 (row sum of squares in double)
//   Q8_K    : upstream ggml-quants.c quantize_row_q8_K_ref   (iscale = -127/max, first max wins, nearest_int)
//   Q8_0    : upstream ggml-quants.c quantize_row_q8_0_ref   (d = amax/127 stored as f16, roundf)
// grid = n_rows (tokens), block = 512.  K % 32 == 0.
// =================================================================================================================
__device__ __forceinline__ void quantize_256_warp(const float* y /*smem or global, 256-aligned block*/, int valid, int fmt,
                                                   int8_t* q, float* dq, int16_t* bs, int blk_index) {
    const int lane = threadIdx.x & 31;
    float v[8];
#pragma unroll
    for (int i = 0; i < 8; i++) v[i] = (lane * 8 + i < valid) ? y[lane * 8 + i] : 0.0f;
    int qi[8];
    if (fmt == ACT_Q8_K) {
        // first element attaining the max |x| decides the sign of the scale
        float amax = 0.0f; int idx = 0x7fffffff;
#pragma unroll
        for (int i = 0; i < 8; i++) { const float ax = fabsf(v[i]); if (ax > amax) { amax = ax; idx = lane * 8 + i; } }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const float oa = __shfl_xor_sync(0xffffffffu, amax, o);
            const int oi = __shfl_xor_sync(0xffffffffu, idx, o);
            if (oa > amax || (oa == amax && oi < idx)) { amax = oa; idx = oi; }
        }
        float mx = 0.0f;
#pragma unroll
        for (int i = 0; i < 8; i++) if (lane * 8 + i == idx) mx = v[i];
        mx = __shfl_sync(0xffffffffu, mx, (idx == 0x7fffffff ? 0 : idx) >> 3);
        if (amax == 0.0f) {
#pragma unroll
            for (int i = 0; i < 8; i++) qi[i] = 0;
            if (lane == 0) dq[blk_index] = 0.0f;
        } else {
            const float iscale = __fdiv_rn(-127.0f, mx);
#pragma unroll
            for (int i = 0; i < 8; i++) qi[i] = min(127, __float2int_rn(__fmul_rn(iscale, v[i])));
            if (lane == 0) dq[blk_index] = __fdiv_rn(1.0f, iscale);
        }
        int s = 0;
#pragma unroll
        for (int i = 0; i < 8; i++) s += qi[i];
        s += __shfl_xor_sync(0xffffffffu, s, 1);
        if ((lane & 1) == 0) bs[blk_index * 16 + (lane >> 1)] = (int16_t)s;
    } else {   // ACT_Q8_0: 32-element blocks = 4 lanes
        float amax = 0.0f;
#pragma unroll
        for (int i = 0; i < 8; i++) amax = fmaxf(amax, fabsf(v[i]));
        amax = fmaxf(amax, __shfl_xor_sync(0xffffffffu, amax, 1));
        amax = fmaxf(amax, __shfl_xor_sync(0xffffffffu, amax, 2));
        const float d = __fdiv_rn(amax, 127.0f);
        const float id = d != 0.0f ? __fdiv_rn(1.0f, d) : 0.0f;
#pragma unroll
        for (int i = 0; i < 8; i++) qi[i] = (int)roundf(__fmul_rn(v[i], id));
        if ((lane & 3) == 0 && lane * 8 < valid) dq[blk_index * 8 + (lane >> 2)] = __half2float(__float2half_rn(d));
    }
    if (lane * 8 < valid) {
        uint2 pk;
        pk.x = (uint32_t)(qi[0] & 0xff) | ((uint32_t)(qi[1] & 0xff) << 8) | ((uint32_t)(qi[2] & 0xff) << 16) | ((uint32_t)(qi[3] & 0xff) << 24);
        pk.y = (uint32_t)(qi[4] & 0xff) | ((uint32_t)(qi[5] & 0xff) << 8) | ((uint32_t)(qi[6] & 0xff) << 16) | ((uint32_t)(qi[7] & 0xff) << 24);
        *reinterpret_cast<uint2*>(q + lane * 8) = pk;
    }
}

template <bool NORM>
__global__ void __launch_bounds__(512) act_prepare_kernel(const float* __restrict__ x, const float* __restrict__ w, int K, float eps,
                                                          int fmt, ActBuf out, int64_t row_stride_q, int64_t row_stride_d, int64_t row_stride_bs) {
    // row-strided outputs so the same kernel serves the batch (prefill) case: row = blockIdx.x
    const int row = blockIdx.x;
    x += (size_t)row * K;
    float* of = out.f32 ? out.f32 + (size_t)row * K : nullptr;
    int8_t* oq = out.q ? out.q + (size_t)row * row_stride_q : nullptr;
    float* od = out.d ? out.d + (size_t)row * row_stride_d : nullptr;
    int16_t* ob = out.bs ? out.bs + (size_t)row * row_stride_bs : nullptr;
    __shared__ double red[16];
    __shared__ float s_scale;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    float scale = 1.0f;
    if (NORM) {
        double sum = 0.0;
        for (int i = tid; i < K; i += 512) { const float v = x[i]; sum += (double)__fmul_rn(v, v); }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
        if (lane == 0) red[wid] = sum;
        __syncthreads();
        if (tid == 0) {
            double tot = 0.0;
            for (int i = 0; i < 16; i++) tot += red[i];
            const float mean = (float)(tot / (double)K);
            s_scale = __fdiv_rn(1.0f, __fsqrt_rn(__fadd_rn(mean, eps)));
        }
        __syncthreads();
        scale = s_scale;
    }
    extern __shared__ float ybuf[];   // [K] normalised row
    for (int i = tid; i < K; i += 512) {
        float v = x[i];
        if (NORM) v = __fmul_rn(__fmul_rn(v, scale), w[i]);
        ybuf[i] = v;
        if (of) of[i] = v;
    }
    __syncthreads();
    if (fmt == ACT_F32) return;
    const int nblk = (K + 255) >> 8;
    for (int b = wid; b < nblk; b += 16)
        quantize_256_warp(ybuf + b * 256, min(256, K - b * 256), fmt, oq + b * 256, od, ob, b);
}

// =================================================================================================================
// dequant-fused mat-vec: one warp owns a PAIR of rows at a time; each lane streams 2 x 128 bit of quantised weights
// per row per step and multiplies them with the int8 activations held in shared memory.
// =================================================================================================================
struct ActView { const int8_t* q; const float* d; const int16_t* bs; const float* f32; };

template <int TYPE> struct RowUnit;
template <> struct RowUnit<QT_Q4_K> {
    uint4 q0, q1, hdr;
    __device__ __forceinline__ void load(const QMat& W, int64_t row, int u) {
        const uint8_t* q = W.p0 + (size_t)row * (W.K >> 1) + (size_t)u * 32;
        q0 = ldg_stream(q); q1 = ldg_stream(q + 16);
        hdr = __ldg(reinterpret_cast<const uint4*>(W.p1 + ((size_t)row * (W.K >> 8) + (u >> 2)) * 16));
    }
    __device__ __forceinline__ float dot(const ActView& A, int u) const {
        const int4* a = reinterpret_cast<const int4*>(A.q + (size_t)u * 64);
        const int4 l0 = a[0], l1 = a[1], h0 = a[2], h1 = a[3];
        int ilo = 0, ihi = 0;
        const uint32_t M = 0x0F0F0F0Fu;
        ilo = __dp4a((int)(q0.x & M), l0.x, ilo); ihi = __dp4a((int)((q0.x >> 4) & M), h0.x, ihi);
        ilo = __dp4a((int)(q0.y & M), l0.y, ilo); ihi = __dp4a((int)((q0.y >> 4) & M), h0.y, ihi);
        ilo = __dp4a((int)(q0.z & M), l0.z, ilo); ihi = __dp4a((int)((q0.z >> 4) & M), h0.z, ihi);
        ilo = __dp4a((int)(q0.w & M), l0.w, ilo); ihi = __dp4a((int)((q0.w >> 4) & M), h0.w, ihi);
        ilo = __dp4a((int)(q1.x & M), l1.x, ilo); ihi = __dp4a((int)((q1.x >> 4) & M), h1.x, ihi);
        ilo = __dp4a((int)(q1.y & M), l1.y, ilo); ihi = __dp4a((int)((q1.y >> 4) & M), h1.y, ihi);
        ilo = __dp4a((int)(q1.z & M), l1.z, ilo); ihi = __dp4a((int)((q1.z >> 4) & M), h1.z, ihi);
        ilo = __dp4a((int)(q1.w & M), l1.w, ilo); ihi = __dp4a((int)((q1.w >> 4) & M), h1.w, ihi);
        uint32_t sc2, mn2; k4_scale_min_pair(hdr, u & 3, sc2, mn2);
        const uint2 bsw = *reinterpret_cast<const uint2*>(A.bs + (size_t)u * 4);   // 4 x int16 sums of 16
        const int blo = (int)(int16_t)(bsw.x & 0xffff) + (int)(int16_t)(bsw.x >> 16);
        const int bhi = (int)(int16_t)(bsw.y & 0xffff) + (int)(int16_t)(bsw.y >> 16);
        const int p = (int)(sc2 & 0xff) * ilo + (int)(sc2 >> 8) * ihi;
        const int pm = (int)(mn2 & 0xff) * blo + (int)(mn2 >> 8) * bhi;
        const float2 dm = hdr_d_dmin(hdr);
        const float ad = A.d[u >> 2];
        return (dm.x * ad) * (float)p - (dm.y * ad) * (float)pm;
    }
};
template <> struct RowUnit<QT_Q5_K> {
    uint4 q0, q1, hdr, h0, h1;
    __device__ __forceinline__ void load(const QMat& W, int64_t row, int u) {
        const uint8_t* q = W.p0 + (size_t)row * (W.K >> 1) + (size_t)u * 32;
        q0 = ldg_stream(q); q1 = ldg_stream(q + 16);
        const size_t sb = (size_t)row * (W.K >> 8) + (u >> 2);
        hdr = __ldg(reinterpret_cast<const uint4*>(W.p1 + sb * 16));
        h0 = __ldg(reinterpret_cast<const uint4*>(W.p2 + sb * 32));
        h1 = __ldg(reinterpret_cast<const uint4*>(W.p2 + sb * 32 + 16));
    }
    __device__ __forceinline__ float dot(const ActView& A, int u) const {
        const int4* a = reinterpret_cast<const int4*>(A.q + (size_t)u * 64);
        const int4 l0 = a[0], l1 = a[1], g0 = a[2], g1 = a[3];
        const int j = u & 3;
        const uint32_t M = 0x0F0F0F0Fu, B = 0x01010101u;
        int ilo = 0, ihi = 0;
#define BLK_Q5_STEP(QW, HW, AL, AH)                                                                  \
        ilo = __dp4a((int)(((QW) & M) | ((((HW) >> (2 * j)) & B) << 4)), (AL), ilo);                 \
        ihi = __dp4a((int)((((QW) >> 4) & M) | ((((HW) >> (2 * j + 1)) & B) << 4)), (AH), ihi);
        BLK_Q5_STEP(q0.x, h0.x, l0.x, g0.x) BLK_Q5_STEP(q0.y, h0.y, l0.y, g0.y)
        BLK_Q5_STEP(q0.z, h0.z, l0.z, g0.z) BLK_Q5_STEP(q0.w, h0.w, l0.w, g0.w)
        BLK_Q5_STEP(q1.x, h1.x, l1.x, g1.x) BLK_Q5_STEP(q1.y, h1.y, l1.y, g1.y)
        BLK_Q5_STEP(q1.z, h1.z, l1.z, g1.z) BLK_Q5_STEP(q1.w, h1.w, l1.w, g1.w)
#undef BLK_Q5_STEP
        uint32_t sc2, mn2; k4_scale_min_pair(hdr, j, sc2, mn2);
        const uint2 bsw = *reinterpret_cast<const uint2*>(A.bs + (size_t)u * 4);
        const int blo = (int)(int16_t)(bsw.x & 0xffff) + (int)(int16_t)(bsw.x >> 16);
        const int bhi = (int)(int16_t)(bsw.y & 0xffff) + (int)(int16_t)(bsw.y >> 16);
        const int p = (int)(sc2 & 0xff) * ilo + (int)(sc2 >> 8) * ihi;
        const int pm = (int)(mn2 & 0xff) * blo + (int)(mn2 >> 8) * bhi;
        const float2 dm = hdr_d_dmin(hdr);
        const float ad = A.d[u >> 2];
        return (dm.x * ad) * (float)p - (dm.y * ad) * (float)pm;
    }
};
template <> struct RowUnit<QT_Q6_K> {
    uint4 l0, l1, h; uint2 sc; uint16_t dh;
    __device__ __forceinline__ void load(const QMat& W, int64_t row, int u) {
        const int s = u >> 2, hh = (u >> 1) & 1, t = u & 1;
        const uint8_t* ql = W.p0 + (size_t)row * (W.K >> 1) + (size_t)s * 128 + hh * 64 + t * 16;
        l0 = ldg_stream(ql); l1 = ldg_stream(ql + 32);
        h = ldg_stream(W.p1 + (size_t)row * (W.K >> 2) + (size_t)s * 64 + hh * 32 + t * 16);
        sc = __ldg(reinterpret_cast<const uint2*>(W.p2 + (size_t)row * (W.K >> 4) + s * 16 + hh * 8));
        dh = __ldg(reinterpret_cast<const uint16_t*>(W.p3) + (size_t)row * (W.K >> 8) + s);
    }
    __device__ __forceinline__ float dot(const ActView& A, int u) const {
        const int s = u >> 2, hh = (u >> 1) & 1, t = u & 1;
        const int e0 = 256 * s + 128 * hh + 16 * t;
        const int4 a0 = *reinterpret_cast<const int4*>(A.q + e0);
        const int4 a1 = *reinterpret_cast<const int4*>(A.q + e0 + 32);
        const int4 a2 = *reinterpret_cast<const int4*>(A.q + e0 + 64);
        const int4 a3 = *reinterpret_cast<const int4*>(A.q + e0 + 96);
        const uint32_t M = 0x0F0F0F0Fu, H = 0x30303030u;
        int i0 = 0, i1 = 0, i2 = 0, i3 = 0;
#define BLK_Q6_STEP(LA, LB, HW, A0, A1, A2, A3)                                     \
        i0 = __dp4a((int)(((LA) & M) | (((HW) << 4) & H)), (A0), i0);               \
        i1 = __dp4a((int)(((LB) & M) | (((HW) << 2) & H)), (A1), i1);               \
        i2 = __dp4a((int)((((LA) >> 4) & M) | ((HW) & H)), (A2), i2);               \
        i3 = __dp4a((int)((((LB) >> 4) & M) | (((HW) >> 2) & H)), (A3), i3);
        BLK_Q6_STEP(l0.x, l1.x, h.x, a0.x, a1.x, a2.x, a3.x)
        BLK_Q6_STEP(l0.y, l1.y, h.y, a0.y, a1.y, a2.y, a3.y)
        BLK_Q6_STEP(l0.z, l1.z, h.z, a0.z, a1.z, a2.z, a3.z)
        BLK_Q6_STEP(l0.w, l1.w, h.w, a0.w, a1.w, a2.w, a3.w)
#undef BLK_Q6_STEP
        // 16-element sums of the activations give the "-32" offset: sum (q-32) a = sum q a - 32 sum a
        const int bi = 16 * s + 8 * hh + t;
        const int b0 = A.bs[bi], b1 = A.bs[bi + 2], b2 = A.bs[bi + 4], b3 = A.bs[bi + 6];
        const uint32_t sl = t ? (sc.x >> 8) : sc.x, sh = t ? (sc.y >> 8) : sc.y;   // bytes t, t+2 | t+4, t+6
        const int s0 = (int)(int8_t)(sl & 0xff), s1 = (int)(int8_t)((sl >> 16) & 0xff);
        const int s2 = (int)(int8_t)(sh & 0xff), s3 = (int)(int8_t)((sh >> 16) & 0xff);
        const int p = s0 * (i0 - 32 * b0) + s1 * (i1 - 32 * b1) + s2 * (i2 - 32 * b2) + s3 * (i3 - 32 * b3);
        const float d = __half2float(__ushort_as_half(dh));
        return (d * A.d[s]) * (float)p;
    }
};
template <> struct RowUnit<QT_Q8_0> {
    uint4 q0, q1; uint16_t dh;
    __device__ __forceinline__ void load(const QMat& W, int64_t row, int u) {
        const uint8_t* q = W.p0 + (size_t)row * W.K + (size_t)u * 32;
        q0 = ldg_stream(q); q1 = ldg_stream(q + 16);
        dh = __ldg(reinterpret_cast<const uint16_t*>(W.p1) + (size_t)row * (W.K >> 5) + u);
    }
    __device__ __forceinline__ float dot(const ActView& A, int u) const {
        const int4* a = reinterpret_cast<const int4*>(A.q + (size_t)u * 32);
        const int4 a0 = a[0], a1 = a[1];
        int i = 0;
        i = __dp4a((int)q0.x, a0.x, i); i = __dp4a((int)q0.y, a0.y, i); i = __dp4a((int)q0.z, a0.z, i); i = __dp4a((int)q0.w, a0.w, i);
        i = __dp4a((int)q1.x, a1.x, i); i = __dp4a((int)q1.y, a1.y, i); i = __dp4a((int)q1.z, a1.z, i); i = __dp4a((int)q1.w, a1.w, i);
        return (float)i * (__half2float(__ushort_as_half(dh)) * A.d[u]);
    }
};
template <> struct RowUnit<QT_F32> {
    float4 w0, w1;
    __device__ __forceinline__ void load(const QMat& W, int64_t row, int u) {
        const float4* p = reinterpret_cast<const float4*>(W.p0) + ((size_t)row * W.K + (size_t)u * 8) / 4;
        w0 = __ldg(p); w1 = __ldg(p + 1);
    }
    __device__ __forceinline__ float dot(const ActView& A, int u) const {
        const float4* x = reinterpret_cast<const float4*>(A.f32 + (size_t)u * 8);
        const float4 x0 = x[0], x1 = x[1];
        return w0.x * x0.x + w0.y * x0.y + w0.z * x0.z + w0.w * x0.w + w1.x * x1.x + w1.y * x1.y + w1.z * x1.z + w1.w * x1.w;
    }
};
template <> struct RowUnit<QT_F16> {
    uint4 w;
    __device__ __forceinline__ void load(const QMat& W, int64_t row, int u) {
        w = __ldg(reinterpret_cast<const uint4*>(W.p0 + ((size_t)row * W.K + (size_t)u * 8) * 2));
    }
    __device__ __forceinline__ float dot(const ActView& A, int u) const {
        // ggml converts the activations to f16 for F16 weights (vec_dot_type F16) and accumulates in f32
        const float* x = A.f32 + (size_t)u * 8;
        const __half2* h = reinterpret_cast<const __half2*>(&w);
        float acc = 0.0f;
#pragma unroll
        for (int i = 0; i < 4; i++) {
            const float2 wf = __half22float2(h[i]);
            acc += wf.x * __half2float(__float2half_rn(x[2 * i])) + wf.y * __half2float(__float2half_rn(x[2 * i + 1]));
        }
        return acc;
    }
};

// two rows (possibly of two different matrices of the same type) against the same activations
template <int TYPE>
__device__ __forceinline__ void dot_pair(const QMat& Wa, int64_t ra, const QMat& Wb, int64_t rb, const ActView& A, float& oa, float& ob) {
    const int lane = threadIdx.x & 31;
    const int units = Wa.K / qmat_unit_elems(TYPE);
    float sa = 0.0f, sb = 0.0f;
    int u = lane;
    // two steps in flight: 8 x 128-bit loads per lane outstanding
    for (; u + 32 < units; u += 64) {
        RowUnit<TYPE> a0, b0, a1, b1;
        a0.load(Wa, ra, u); b0.load(Wb, rb, u); a1.load(Wa, ra, u + 32); b1.load(Wb, rb, u + 32);
        sa += a0.dot(A, u); sb += b0.dot(A, u); sa += a1.dot(A, u + 32); sb += b1.dot(A, u + 32);
    }
    if (u < units) {
        RowUnit<TYPE> a0, b0;
        a0.load(Wa, ra, u); b0.load(Wb, rb, u);
        sa += a0.dot(A, u); sb += b0.dot(A, u);
    }
    oa = warp_sum(sa); ob = warp_sum(sb);
}

__device__ __forceinline__ void dot_pair_any(const QMat& Wa, int64_t ra, const QMat& Wb, int64_t rb, const ActView& A, float& oa, float& ob) {
    switch (Wa.type) {
        case QT_Q4_K: dot_pair<QT_Q4_K>(Wa, ra, Wb, rb, A, oa, ob); break;
        case QT_Q6_K: dot_pair<QT_Q6_K>(Wa, ra, Wb, rb, A, oa, ob); break;
        case QT_Q8_0: dot_pair<QT_Q8_0>(Wa, ra, Wb, rb, A, oa, ob); break;
        case QT_Q5_K: dot_pair<QT_Q5_K>(Wa, ra, Wb, rb, A, oa, ob); break;
        case QT_F32: dot_pair<QT_F32>(Wa, ra, Wb, rb, A, oa, ob); break;
        default: dot_pair<QT_F16>(Wa, ra, Wb, rb, A, oa, ob); break;
    }
}

enum : int { EPI_STORE = 0, EPI_RESID = 1, EPI_QKV = 2, EPI_SWIGLU = 3 };

struct GemvSeg {
    QMat W;                 // rows of this segment
    const float* bias;      // optional [N]
    int pair0;              // first pair index of this segment
    int kind;               // EPI_QKV only: 0 = q, 1 = k, 2 = v
};

struct GemvArgs {
    GemvSeg seg[3];
    int nseg;
    int total_pairs;
    ActBuf act;             // global-memory activations (prepared by act_prepare_kernel)
    int act_fmt;
    float* out;             // EPI_STORE / EPI_RESID / EPI_SWIGLU destination; EPI_QKV: q vector [n_head*d_head] f32
    // EPI_QKV
    int d_head, neox;
    const float2* rope_cs;  // [d_head/2] {cos, sin} of this step's position
    const int32_t* pos;     // device scalar: position of the token being decoded
    __half* k_pool; __half* v_pool;   // this layer's KV pages [n_pages][KV_PAGE][n_kv*d_head]
    const int32_t* page_table;
    int kv_dim;             // n_head_kv * d_head
};

// stage the prepared activations into shared memory (int8 + scales + 16-sums); F32 activations stay in global
__device__ __forceinline__ ActView stage_activations(const GemvArgs& a, int K, unsigned char* smem) {
    ActView v{nullptr, nullptr, nullptr, a.act.f32};
    if (a.act_fmt == ACT_F32) return v;
    const int tid = threadIdx.x, nt = blockDim.x;
    int8_t* sq = reinterpret_cast<int8_t*>(smem);
    const int nd = (a.act_fmt == ACT_Q8_K) ? (K >> 8) : (K >> 5);
    float* sd = reinterpret_cast<float*>(smem + K);
    int16_t* sb = reinterpret_cast<int16_t*>(smem + K + ((nd * 4 + 15) & ~15));
    for (int i = tid; i < (K >> 4); i += nt) reinterpret_cast<uint4*>(sq)[i] = reinterpret_cast<const uint4*>(a.act.q)[i];
    for (int i = tid; i < nd; i += nt) sd[i] = a.act.d[i];
    if (a.act_fmt == ACT_Q8_K) for (int i = tid; i < (K >> 5); i += nt) reinterpret_cast<uint32_t*>(sb)[i] = reinterpret_cast<const uint32_t*>(a.act.bs)[i];
    v.q = sq; v.d = sd; v.bs = sb;
    return v;
}
__host__ __device__ inline size_t gemv_smem_bytes(int K, int fmt) {
    if (fmt == ACT_F32) return 0;
    const int nd = (fmt == ACT_Q8_K) ? (K >> 8) : (K >> 5);
    return (size_t)K + (size_t)((nd * 4 + 15) & ~15) + (fmt == ACT_Q8_K ? (size_t)(K >> 4) * 2 : 0) + 16;
}

template <int EPI>
__global__ void __launch_bounds__(GEMV_THREADS) gemv_pairs_kernel(const GemvArgs a) {
    extern __shared__ __align__(16) unsigned char gemv_smem[];
    const int K = a.seg[0].W.K;
    const ActView A = stage_activations(a, K, gemv_smem);
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const int warp = (blockIdx.x * GEMV_THREADS + threadIdx.x) >> 5;
    const int nwarps = (gridDim.x * GEMV_THREADS) >> 5;
    for (int pair = warp; pair < a.total_pairs; pair += nwarps) {
        if (EPI == EPI_SWIGLU) {
            // seg[0] = gate, seg[1] = up: pair p = (gate row p, up row p)
            float g, u;
            dot_pair_any(a.seg[0].W, pair, a.seg[1].W, pair, A, g, u);
            if (lane == 0) a.out[pair] = (g / (1.0f + expf(-g))) * u;     // ggml_silu_f32 then ggml_mul
            continue;
        }
        int si = 0;
        if (a.nseg > 1 && pair >= a.seg[1].pair0) si = 1;
        if (a.nseg > 2 && pair >= a.seg[2].pair0) si = 2;
        const GemvSeg& S = a.seg[si];
        const int p = pair - S.pair0;
        int r0 = 2 * p, r1 = 2 * p + 1;
        if (EPI == EPI_QKV && a.neox && S.kind != 2) {
            const int hd = a.d_head >> 1;
            r0 = (p / hd) * a.d_head + (p % hd); r1 = r0 + hd;
        }
        float v0, v1;
        dot_pair_any(S.W, r0, S.W, r1, A, v0, v1);
        if (lane != 0) continue;
        if (S.bias) { v0 += S.bias[r0]; v1 += S.bias[r1]; }
        if (EPI == EPI_STORE) { a.out[r0] = v0; a.out[r1] = v1; }
        else if (EPI == EPI_RESID) { a.out[r0] += v0; a.out[r1] += v1; }
        else if (EPI == EPI_QKV) {
            if (S.kind != 2) {     // rotary embedding on the (r0, r1) pair: ggml rope NORM / NEOX
                const int i = a.neox ? (r0 % a.d_head) : ((r0 % a.d_head) >> 1);
                const float2 cs = a.rope_cs[i];
                const float x0 = v0, x1 = v1;
                v0 = x0 * cs.x - x1 * cs.y;
                v1 = x0 * cs.y + x1 * cs.x;
            }
            if (S.kind == 0) { a.out[r0] = v0; a.out[r1] = v1; }
            else {
                const int pos = a.pos[0];
                const size_t base = ((size_t)a.page_table[pos / KV_PAGE] * KV_PAGE + (pos % KV_PAGE)) * a.kv_dim;
                __half* dst = (S.kind == 1) ? a.k_pool : a.v_pool;
                dst[base + r0] = __float2half_rn(v0);      // ggml_cpy f32 -> f16 into the cache
                dst[base + r1] = __float2half_rn(v1);
            }
        }
    }
}

// =================================================================================================================
// decode attention over the paged f16 KV cache, split along the context.
//   follows llama-graph.cpp build_attn_mha with flash_attn = false, in ggml's own order of operations:
//     kq = K.q (q rounded to f16, f32 accumulate) * scale ; soft_max_ext: max, expf, sum in double, p = e * (1/sum) ;
//     p rounded to f16 ; out = V.p (f32 accumulate).
//   The probabilities must be normalised BEFORE the f16 rounding: the Q8_K quantisation that follows amplifies a
//   2^-11 relative change of the attention output into ~0.1 logit differences.
//   kernel 1 (scores): grid (n_head_kv, n_split), block d_head: thread = token; writes scaled scores.
//   kernel 2 (pv)    : grid (n_head_kv, n_split), block d_head: every CTA reduces max / sum over the whole row
//                      (scores are L2 resident), then its own token slice: thread = output dim.
//   kernel 3 (combine): sums the split partials in fixed order and quantises for the Wo mat-vec.
// =================================================================================================================
struct AttnArgs {
    const float* q;            // [n_head][d_head] f32 (post-RoPE)
    const __half* k_pool; const __half* v_pool; const int32_t* page_table;
    const int32_t* pos;        // device scalar: n_kv = pos + 1
    int n_head, n_head_kv, d_head, kv_dim, n_split;
    float scale;
    float* scores;             // [n_head][score_stride]
    int score_stride;
    float* part_o;             // [n_head][n_split][d_head]
};

template <int DH>
__global__ void __launch_bounds__(DH) attn_scores_kernel(const AttnArgs a) {
    const int hk = blockIdx.x, split = blockIdx.y, tid = threadIdx.x;
    const int gq = a.n_head / a.n_head_kv;
    const int n_kv = a.pos[0] + 1;
    const int per = (n_kv + a.n_split - 1) / a.n_split;
    const int t_begin = split * per, t_end = min(n_kv, t_begin + per);
    if (t_begin >= t_end) return;
    __shared__ __half sq[MAX_GQ * DH];
    for (int i = tid; i < gq * DH; i += DH) sq[i] = __float2half_rn(a.q[(size_t)(hk * gq) * DH + i]);
    __syncthreads();
    for (int t = t_begin + tid; t < t_end; t += DH) {
        const __half* kr = a.k_pool + ((size_t)a.page_table[t / KV_PAGE] * KV_PAGE + (t % KV_PAGE)) * a.kv_dim + (size_t)hk * DH;
        float s[MAX_GQ];
#pragma unroll
        for (int g = 0; g < MAX_GQ; g++) s[g] = 0.0f;
#pragma unroll 4
        for (int c = 0; c < DH / 8; c++) {
            const uint4 kv = *reinterpret_cast<const uint4*>(kr + c * 8);
            const __half2* kh = reinterpret_cast<const __half2*>(&kv);
            float kf[8];
#pragma unroll
            for (int i = 0; i < 4; i++) { const float2 f = __half22float2(kh[i]); kf[2 * i] = f.x; kf[2 * i + 1] = f.y; }
#pragma unroll
            for (int g = 0; g < MAX_GQ; g++) if (g < gq) {
                const __half2* qh = reinterpret_cast<const __half2*>(sq + g * DH + c * 8);
#pragma unroll
                for (int i = 0; i < 4; i++) { const float2 f = __half22float2(qh[i]); s[g] += kf[2 * i] * f.x; s[g] += kf[2 * i + 1] * f.y; }
            }
        }
#pragma unroll
        for (int g = 0; g < MAX_GQ; g++) if (g < gq) a.scores[(size_t)(hk * gq + g) * a.score_stride + t] = __fmul_rn(s[g], a.scale);
    }
}

template <int DH>
__global__ void __launch_bounds__(DH) attn_pv_kernel(const AttnArgs a) {
    const int hk = blockIdx.x, split = blockIdx.y, tid = threadIdx.x;
    const int gq = a.n_head / a.n_head_kv;
    const int n_kv = a.pos[0] + 1;
    const int per = (n_kv + a.n_split - 1) / a.n_split;
    const int t_begin = split * per, t_end = min(n_kv, t_begin + per);
    const int lane = tid & 31, wid = tid >> 5;
    constexpr int NW = DH / 32;
    float* po = a.part_o + ((size_t)(hk * gq) * a.n_split + split) * DH;
    if (t_begin >= t_end) {
        for (int g = 0; g < gq; g++) po[(size_t)g * a.n_split * DH + tid] = 0.0f;
        return;
    }
    __shared__ float sred[MAX_GQ][NW];
    __shared__ double dred[MAX_GQ][NW];
    __shared__ float sp[MAX_GQ][DH];
    float M[MAX_GQ], inv[MAX_GQ];
    // row max
#pragma unroll
    for (int g = 0; g < MAX_GQ; g++) if (g < gq) {
        const float* sr = a.scores + (size_t)(hk * gq + g) * a.score_stride;
        float mx = -INFINITY;
        for (int t = tid; t < n_kv; t += DH) mx = fmaxf(mx, sr[t]);
        mx = warp_max(mx);
        if (lane == 0) sred[g][wid] = mx;
    }
    __syncthreads();
#pragma unroll
    for (int g = 0; g < MAX_GQ; g++) if (g < gq) {
        float mx = sred[g][0];
#pragma unroll
        for (int w = 1; w < NW; w++) mx = fmaxf(mx, sred[g][w]);
        M[g] = mx;
    }
    // row sum of expf(s - max) in double (ggml_float)
#pragma unroll
    for (int g = 0; g < MAX_GQ; g++) if (g < gq) {
        const float* sr = a.scores + (size_t)(hk * gq + g) * a.score_stride;
        double sum = 0.0;
        for (int t = tid; t < n_kv; t += DH) sum += (double)expf(sr[t] - M[g]);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
        if (lane == 0) dred[g][wid] = sum;
    }
    __syncthreads();
#pragma unroll
    for (int g = 0; g < MAX_GQ; g++) if (g < gq) {
        double sum = 0.0;
#pragma unroll
        for (int w = 0; w < NW; w++) sum += dred[g][w];
        inv[g] = (float)(1.0 / sum);
    }
    float acc[MAX_GQ];
#pragma unroll
    for (int g = 0; g < MAX_GQ; g++) acc[g] = 0.0f;
    for (int t0 = t_begin; t0 < t_end; t0 += DH) {
        const int t = t0 + tid;
#pragma unroll
        for (int g = 0; g < MAX_GQ; g++) if (g < gq) {
            float p = 0.0f;
            if (t < t_end) p = __fmul_rn(expf(a.scores[(size_t)(hk * gq + g) * a.score_stride + t] - M[g]), inv[g]);
            sp[g][tid] = __half2float(__float2half_rn(p));
        }
        __syncthreads();
        const int nt = min(DH, t_end - t0);
        for (int j = 0; j < nt; j++) {
            const int tj = t0 + j;
            const float v = __half2float(a.v_pool[((size_t)a.page_table[tj / KV_PAGE] * KV_PAGE + (tj % KV_PAGE)) * a.kv_dim + (size_t)hk * DH + tid]);
#pragma unroll
            for (int g = 0; g < MAX_GQ; g++) if (g < gq) acc[g] += sp[g][j] * v;
        }
        __syncthreads();
    }
#pragma unroll
    for (int g = 0; g < MAX_GQ; g++) if (g < gq) po[(size_t)g * a.n_split * DH + tid] = acc[g];
}

// sum the split partials (fixed order) and quantise the attention output for the Wo mat-vec.
// grid = n_head*d_head/256 CTAs, block = 256 (one Q8_K super-block of the output each)
__global__ void __launch_bounds__(256) attn_combine_kernel(const float* __restrict__ part_o, int d_head, int n_split, int fmt, ActBuf out) {
    const int e = blockIdx.x * 256 + threadIdx.x;
    const int h = e / d_head, dd = e % d_head;
    float o = 0.0f;
    for (int s = 0; s < n_split; s++) o += part_o[((size_t)h * n_split + s) * d_head + dd];
    __shared__ float y[256];
    y[threadIdx.x] = o;
    if (out.f32) out.f32[e] = o;
    __syncthreads();
    if (fmt == ACT_F32) return;
    if (threadIdx.x < 32) quantize_256_warp(y, 256, fmt, out.q + blockIdx.x * 256, out.d, out.bs, blockIdx.x);
}

// quantise an f32 vector (no norm): grid = ceil(K/256/8), block = 256 (8 warps, one super-block each)
__global__ void __launch_bounds__(256) act_quant_kernel(const float* __restrict__ x, int K, int fmt, ActBuf out) {
    const int b = blockIdx.x * 8 + (threadIdx.x >> 5);
    if (b * 256 >= K) return;
    quantize_256_warp(x + (size_t)b * 256, min(256, K - b * 256), fmt, out.q + (size_t)b * 256, out.d, out.bs, b);
}

// =================================================================================================================
// top-k of the vocabulary logits (Session.cpp:246-261 does a full std::sort of n_vocab pairs per token on the host)
// stage 1: each CTA bitonic-sorts TOPK_CHUNK logits in shared memory and emits its best TOPK_MAX
// stage 2: one CTA merges the per-chunk candidates.  Order: logit descending, ties by lower id.
// =================================================================================================================
__device__ __forceinline__ bool td_before(float la, int ia, float lb, int ib) { return la > lb || (la == lb && ia < ib); }

template <int N, int THREADS>
__device__ inline void bitonic_sort_desc(float* key, int* idx) {
    for (int k = 2; k <= N; k <<= 1) {
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int i = threadIdx.x; i < N; i += THREADS) {
                const int p = i ^ j;
                if (p > i) {
                    const bool up = ((i & k) == 0);     // "up" segments hold descending order
                    const float a = key[i], b = key[p]; const int ia = idx[i], ib = idx[p];
                    const bool a_first = td_before(a, ia, b, ib);
                    if (up ? !a_first : a_first) { key[i] = b; key[p] = a; idx[i] = ib; idx[p] = ia; }
                }
            }
            __syncthreads();
        }
    }
}

__global__ void __launch_bounds__(256) topk_stage1_kernel(const float* __restrict__ logits, int n, float* cand_l, int* cand_i) {
    __shared__ float key[TOPK_CHUNK];
    __shared__ int idx[TOPK_CHUNK];
    const int base = blockIdx.x * TOPK_CHUNK;
    for (int i = threadIdx.x; i < TOPK_CHUNK; i += 256) {
        const int g = base + i;
        key[i] = g < n ? logits[g] : -INFINITY;
        idx[i] = g < n ? g : 0x7fffffff;
    }
    __syncthreads();
    bitonic_sort_desc<TOPK_CHUNK, 256>(key, idx);
    for (int i = threadIdx.x; i < TOPK_MAX; i += 256) { cand_l[blockIdx.x * TOPK_MAX + i] = key[i]; cand_i[blockIdx.x * TOPK_MAX + i] = idx[i]; }
}

// one CTA, 1024 threads; n_cand <= 16384 candidates processed in rounds of 2048 keeping the best TOPK_MAX
__global__ void __launch_bounds__(1024) topk_stage2_kernel(const float* __restrict__ cand_l, const int* __restrict__ cand_i, int n_cand,
                                                          int k, int32_t* out_ids, float* out_logits) {
    __shared__ float key[2048];
    __shared__ int idx[2048];
    for (int i = threadIdx.x; i < TOPK_MAX; i += 1024) { key[i] = -INFINITY; idx[i] = 0x7fffffff; }
    for (int base = 0; base < n_cand; base += 2048 - TOPK_MAX) {
        for (int i = threadIdx.x; i < 2048 - TOPK_MAX; i += 1024) {
            const int g = base + i;
            key[TOPK_MAX + i] = g < n_cand ? cand_l[g] : -INFINITY;
            idx[TOPK_MAX + i] = g < n_cand ? cand_i[g] : 0x7fffffff;
        }
        __syncthreads();
        bitonic_sort_desc<2048, 1024>(key, idx);
    }
    for (int i = threadIdx.x; i < k; i += 1024) { out_ids[i] = idx[i]; out_logits[i] = key[i]; }
}

__global__ void gather_logits_kernel(const float* __restrict__ logits, int n_vocab, const int32_t* __restrict__ ids, int n, float* out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) { const int id = ids[i]; out[i] = (id >= 0 && id < n_vocab) ? logits[id] : -INFINITY; }
}

__global__ void advance_pos_kernel(int32_t* pos, int n) { if (threadIdx.x == 0 && blockIdx.x == 0) pos[0] += n; }

} // namespace blk
