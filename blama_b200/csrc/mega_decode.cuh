// mega_decode.cuh -- the whole batch-1 decode step (Session::getToken -> llama_decode(1 token), reference
// inference/code/llama/Session.cpp:169-190, 395-401) as ONE persistent cooperative kernel.
//
// Why: a decode step of an 8B model is ~290 dependent operations of 2-20 us each.  As separate kernels (even with
// PDL + CUDA graph) every boundary drains the memory system; measured 34 % of the HBM roofline.  Here one CTA per SM
// stays resident for the whole token and HBM never stops streaming:
//
//   * weights live in a second, "stream" copy laid out at load time as [pair][K-slice][row a | row b] chunks: the bytes one
//     warp needs for one row slice are contiguous (2-4 KB).  Q4_K / Q5_K chunks are ggml's own blocks (144 / 176 B are
//     16 B multiples); Q6_K / Q8_0 are re-tiled to 208 B / 272 B super-blocks.
//   * every warp owns a private 3-slot shared-memory ring that its lane 0 keeps full with cp.async.bulk (TMA 1-D) copies,
//     completion on an mbarrier.  The chunk sequence of a warp is a pure function of (cta, warp): it never depends on
//     activations, so the ring runs AHEAD across phase boundaries and across the grid barriers: while the CTAs
//     synchronise and re-quantise activations, the next phase's weights are already landing.
//   * one lane owns one HALF super-block (128 weights) of a row slice; the 128 int8 activations it needs are the same for
//     every row of the phase, so they sit in 32 registers (loaded once per phase) -- the inner loop reads only weights
//     from shared memory: ~0.5 instructions per weight for Q4_K.
//   * phases of a layer: QKV (+bias, RoPE, KV-page write) | attention scores | softmax + V.p (+ split combine by the
//     last CTA of each KV head) | Wo (+residual) | gate/up (SwiGLU) | down (+residual); then final norm + lm_head.
//     Phases are separated by a grid barrier (one atomic + spin per CTA); activations are exchanged through L2
//     (ld.global.cg) and re-quantised redundantly by every CTA in the prologue of the phase that consumes them.
//
// Arithmetic contract: identical to the per-kernel path (decode_kernels.cuh): activations quantised exactly as ggml's
// quantize_row_q8_K / q8_0, integer dot products bit-identical to ggml_vec_dot_*, attention in ggml's order
// (max -> expf -> sum in double -> p * (1/sum) -> f16 -> V.p).  Only the order of fp32 additions differs.
#pragma once
#include "gemv_ring.cuh"     // mbarrier / bulk-copy PTX wrappers
#include "mega_decode.hpp"

namespace blk {

// ---- small PTX helpers ------------------------------------------------------------------------------------------------
__device__ __forceinline__ int dp4a_us(uint32_t a_u8, int b_s8, int c) {
    int d; asm("dp4a.u32.s32 %0, %1, %2, %3;" : "=r"(d) : "r"(a_u8), "r"(b_s8), "r"(c)); return d;
}
__device__ __forceinline__ uint4 lds128(const void* p) { return *reinterpret_cast<const uint4*>(p); }
__device__ __forceinline__ unsigned int ld_acquire_u32(const unsigned int* p) {
    unsigned int v; asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory"); return v;
}
__device__ __forceinline__ int sext8(uint32_t w, int k) { return (int)(w << (24 - 8 * k)) >> 24; }

// activations of one lane: the 128 int8 values of its half super-block plus their scales / partial sums
struct LaneAct {
    int q[32];
    float d0;            // Q8_K: the block scale
    uint32_t aux[6];     // Q8_K: [0..3] 8 x int16 sums of 16 consecutive activations, [4..5] 4 x int16 sums of 32
                         // Q8_0: [0..3] the four 32-element block scales (float bits)
};

// ---- per-type dot product of one half super-block (weights in wb, loaded from the ring) ---------------------------------
__device__ __forceinline__ float mg_dot_q4k(const uint4* wb, int hf, const LaneAct& A) {
    const uint4 hdr = wb[0];
    uint32_t sc4, mn4;
    if (hf == 0) { sc4 = hdr.y & 0x3F3F3F3Fu; mn4 = hdr.z & 0x3F3F3F3Fu; }
    else { sc4 = (hdr.w & 0x0F0F0F0Fu) | ((hdr.y >> 2) & 0x30303030u); mn4 = ((hdr.w >> 4) & 0x0F0F0F0Fu) | ((hdr.z >> 2) & 0x30303030u); }
    const uint32_t* qw = reinterpret_cast<const uint32_t*>(wb + 1);
    int i0 = 0, i1 = 0, i2 = 0, i3 = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) {
        i0 = __dp4a((int)(qw[i] & 0x0F0F0F0Fu), A.q[i], i0);
        i1 = dp4a_us(qw[i] & 0xF0F0F0F0u, A.q[8 + i], i1);            // 16 x the high-nibble sum (exact)
        i2 = __dp4a((int)(qw[8 + i] & 0x0F0F0F0Fu), A.q[16 + i], i2);
        i3 = dp4a_us(qw[8 + i] & 0xF0F0F0F0u, A.q[24 + i], i3);
    }
    i1 >>= 4; i3 >>= 4;
    const int p = (int)(sc4 & 0xff) * i0 + (int)((sc4 >> 8) & 0xff) * i1 + (int)((sc4 >> 16) & 0xff) * i2 + (int)(sc4 >> 24) * i3;
    int pm = __dp2a_lo((int)A.aux[4], (int)mn4, 0);
    pm = __dp2a_hi((int)A.aux[5], (int)mn4, pm);
    const float2 dm = __half22float2(*reinterpret_cast<const __half2*>(&hdr.x));
    return (dm.x * A.d0) * (float)p - (dm.y * A.d0) * (float)pm;
}

__device__ __forceinline__ float mg_dot_q5k(const uint4* wb, int hf, const LaneAct& A) {
    const uint4 hdr = wb[0];
    uint32_t sc4, mn4;
    if (hf == 0) { sc4 = hdr.y & 0x3F3F3F3Fu; mn4 = hdr.z & 0x3F3F3F3Fu; }
    else { sc4 = (hdr.w & 0x0F0F0F0Fu) | ((hdr.y >> 2) & 0x30303030u); mn4 = ((hdr.w >> 4) & 0x0F0F0F0Fu) | ((hdr.z >> 2) & 0x30303030u); }
    const uint32_t* hw = reinterpret_cast<const uint32_t*>(wb + 1);      // 8 words: high bits of elements l = 4i..4i+3
    const uint32_t* qw = reinterpret_cast<const uint32_t*>(wb + 3);
    const uint32_t M = 0x0F0F0F0Fu, B = 0x01010101u;
    const int j0 = 4 * hf;                                              // bit of the first 32-element sub-block of this half
    int i0 = 0, i1 = 0, i2 = 0, i3 = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) {
        const uint32_t h = hw[i] >> j0;
        i0 = __dp4a((int)((qw[i] & M) | ((h & B) << 4)), A.q[i], i0);
        i1 = __dp4a((int)(((qw[i] >> 4) & M) | (((h >> 1) & B) << 4)), A.q[8 + i], i1);
        i2 = __dp4a((int)((qw[8 + i] & M) | (((h >> 2) & B) << 4)), A.q[16 + i], i2);
        i3 = __dp4a((int)(((qw[8 + i] >> 4) & M) | (((h >> 3) & B) << 4)), A.q[24 + i], i3);
    }
    const int p = (int)(sc4 & 0xff) * i0 + (int)((sc4 >> 8) & 0xff) * i1 + (int)((sc4 >> 16) & 0xff) * i2 + (int)(sc4 >> 24) * i3;
    int pm = __dp2a_lo((int)A.aux[4], (int)mn4, 0);
    pm = __dp2a_hi((int)A.aux[5], (int)mn4, pm);
    const float2 dm = __half22float2(*reinterpret_cast<const __half2*>(&hdr.x));
    return (dm.x * A.d0) * (float)p - (dm.y * A.d0) * (float)pm;
}

// Q6_K half: ql words l0[i] = wb[0..1], l1[i] = wb[2..3] (elements l and l + 32), qh words wb[4..5], scales wb[6].xy, d wb[6].z
__device__ __forceinline__ float mg_dot_q6k(const uint4* wb, const LaneAct& A) {
    const uint32_t* l0 = reinterpret_cast<const uint32_t*>(wb);
    const uint32_t* l1 = reinterpret_cast<const uint32_t*>(wb + 2);
    const uint32_t* hw = reinterpret_cast<const uint32_t*>(wb + 4);
    const uint32_t M = 0x0F0F0F0Fu, H = 0x30303030u;
    int is[8];
#pragma unroll
    for (int k = 0; k < 8; k++) is[k] = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) {
        const int t = i >> 2;
        const uint32_t h = hw[i];
        is[0 + t] = __dp4a((int)((l0[i] & M) | ((h << 4) & H)), A.q[i], is[0 + t]);
        is[2 + t] = __dp4a((int)((l1[i] & M) | ((h << 2) & H)), A.q[8 + i], is[2 + t]);
        is[4 + t] = __dp4a((int)(((l0[i] >> 4) & M) | (h & H)), A.q[16 + i], is[4 + t]);
        is[6 + t] = __dp4a((int)(((l1[i] >> 4) & M) | ((h >> 2) & H)), A.q[24 + i], is[6 + t]);
    }
    const uint32_t sx = wb[6].x, sy = wb[6].y;
    int p = 0;
#pragma unroll
    for (int k = 0; k < 4; k++) { p += sext8(sx, k) * is[k]; p += sext8(sy, k) * is[4 + k]; }
    // sum (q - 32) a = sum q a - 32 sum a, with the 16-element activation sums
    int t = __dp2a_lo((int)A.aux[0], (int)sx, 0);
    t = __dp2a_hi((int)A.aux[1], (int)sx, t);
    t = __dp2a_lo((int)A.aux[2], (int)sy, t);
    t = __dp2a_hi((int)A.aux[3], (int)sy, t);
    p -= 32 * t;
    const float d = __half2float(__ushort_as_half((unsigned short)(wb[6].z & 0xffff)));
    return (d * A.d0) * (float)p;
}

// Q8_0 half super-block = four 32-element blocks: int8 weights wb[0..7], f16 scales wb[8].xy
__device__ __forceinline__ float mg_dot_q80(const uint4* wb, const LaneAct& A) {
    const int* qw = reinterpret_cast<const int*>(wb);
    const __half2 d01 = *reinterpret_cast<const __half2*>(&wb[8].x), d23 = *reinterpret_cast<const __half2*>(&wb[8].y);
    const float dw[4] = {__low2float(d01), __high2float(d01), __low2float(d23), __high2float(d23)};
    float v = 0.0f;
#pragma unroll
    for (int b = 0; b < 4; b++) {
        int s = 0;
#pragma unroll
        for (int j = 0; j < 8; j++) s = __dp4a(qw[8 * b + j], A.q[8 * b + j], s);
        v += (float)s * (dw[b] * __uint_as_float(A.aux[b]));
    }
    return v;
}

// ---- activation quantisation of one 256-element block by one warp, into the swizzled shared-memory layout ---------------
// Same arithmetic as quantize_256_warp (gemv_kernels.cuh).  The int8 values of half super-block hs live at
// sq + hs*128, 16-byte chunk c stored at chunk position c ^ (hs & 7): the per-lane register loads (stride 128 B between
// lanes) are then bank-conflict free.
template <bool NORM>
__device__ __forceinline__ void mg_quantize_block(const float* __restrict__ y /*global, block start*/, int fmt, int b,
                                                  int8_t* sq, float* sd, int16_t* sbs, float scale, const float* __restrict__ w) {
    const int lane = threadIdx.x & 31;
    float v[8];
    {
        const float4 a = __ldcg(reinterpret_cast<const float4*>(y) + lane * 2), c = __ldcg(reinterpret_cast<const float4*>(y) + lane * 2 + 1);
        v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = c.x; v[5] = c.y; v[6] = c.z; v[7] = c.w;
        if (NORM) {
            const float4 wa = __ldg(reinterpret_cast<const float4*>(w) + lane * 2), wc = __ldg(reinterpret_cast<const float4*>(w) + lane * 2 + 1);
            const float ww[8] = {wa.x, wa.y, wa.z, wa.w, wc.x, wc.y, wc.z, wc.w};
#pragma unroll
            for (int i = 0; i < 8; i++) v[i] = __fmul_rn(__fmul_rn(v[i], scale), ww[i]);
        }
    }
    int qi[8];
    if (fmt == ACT_Q8_K) {
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
            if (lane == 0) sd[b] = 0.0f;
        } else {
            const float iscale = __fdiv_rn(-127.0f, mx);
#pragma unroll
            for (int i = 0; i < 8; i++) qi[i] = min(127, __float2int_rn(__fmul_rn(iscale, v[i])));
            if (lane == 0) sd[b] = __fdiv_rn(1.0f, iscale);
        }
        int s = 0;
#pragma unroll
        for (int i = 0; i < 8; i++) s += qi[i];
        s += __shfl_xor_sync(0xffffffffu, s, 1);
        if ((lane & 1) == 0) sbs[b * 16 + (lane >> 1)] = (int16_t)s;
    } else {   // ACT_Q8_0
        float amax = 0.0f;
#pragma unroll
        for (int i = 0; i < 8; i++) amax = fmaxf(amax, fabsf(v[i]));
        amax = fmaxf(amax, __shfl_xor_sync(0xffffffffu, amax, 1));
        amax = fmaxf(amax, __shfl_xor_sync(0xffffffffu, amax, 2));
        const float d = __fdiv_rn(amax, 127.0f);
        const float id = d != 0.0f ? __fdiv_rn(1.0f, d) : 0.0f;
#pragma unroll
        for (int i = 0; i < 8; i++) qi[i] = (int)roundf(__fmul_rn(v[i], id));
        if ((lane & 3) == 0) sd[b * 8 + (lane >> 2)] = __half2float(__float2half_rn(d));
    }
    uint2 pk;
    pk.x = (uint32_t)(qi[0] & 0xff) | ((uint32_t)(qi[1] & 0xff) << 8) | ((uint32_t)(qi[2] & 0xff) << 16) | ((uint32_t)(qi[3] & 0xff) << 24);
    pk.y = (uint32_t)(qi[4] & 0xff) | ((uint32_t)(qi[5] & 0xff) << 8) | ((uint32_t)(qi[6] & 0xff) << 16) | ((uint32_t)(qi[7] & 0xff) << 24);
    const int hs = 2 * b + (lane >> 4), c = (lane >> 1) & 7;
    *reinterpret_cast<uint2*>(sq + hs * 128 + ((c ^ (hs & 7)) << 4) + (lane & 1) * 8) = pk;
}

// ---- stream cursor: the chunk sequence of one warp over the whole token ---------------------------------------------
// Used once per (model, device) by mg_chunk_list_kernel to write every warp's chunk descriptors {address, bytes} to global
// memory; the decode kernel's producer then just walks its list (one 16 B load per 2-4 KB chunk).
struct MgCursor {
    int ph, s, k, c;          // phase, segment, item (row pair), chunk of the item
    int n;                    // items of this warp's group in the current segment
    int cpp;                  // chunks per item (1 if the chunk holds both rows, else 2)
    uint32_t bytes;           // bytes per chunk
    const uint8_t* addr0;     // chunk (item 0, 0)
    size_t item_stride;
};

// group bookkeeping of a warp in a phase
struct MgGroup { int W, L, rpc, NG, NGtot, gg, ws; bool active; };
__device__ __forceinline__ MgGroup mg_group(const MegaPhase* ph, int n_cta, int cta, int warp) {
    MgGroup g;
    g.W = ph->W; g.L = ph->L; g.rpc = ph->rpc;
    g.NG = MG_WARPS / g.W; g.NGtot = n_cta * g.NG;
    g.active = warp < g.NG * g.W;
    g.gg = cta * g.NG + warp / g.W;
    g.ws = warp % g.W;
    return g;
}
__device__ __forceinline__ void mg_seg_items(const MgGroup& g, const MegaSeg* sg, int& vg, int& n) {
    vg = (g.gg + sg->rot) % g.NGtot;
    n = (g.active && vg < sg->n_pairs) ? (sg->n_pairs - 1 - vg) / g.NGtot + 1 : 0;
}
__device__ inline void mg_cursor_load(MgCursor& cu, const MegaPhase* phases, int n_cta, int cta, int warp) {
    const MegaPhase* ph = phases + cu.ph;
    const MgGroup g = mg_group(ph, n_cta, cta, warp);
    const MegaSeg* sg = ph->seg + cu.s;
    int vg; mg_seg_items(g, sg, vg, cu.n);
    cu.cpp = 2 / g.rpc;
    cu.bytes = (uint32_t)(sg->slice_bytes * g.rpc);
    cu.addr0 = sg->base + ((size_t)vg * g.W + g.ws) * 2 * (size_t)sg->slice_bytes;
    cu.item_stride = (size_t)g.NGtot * g.W * 2 * (size_t)sg->slice_bytes;
    cu.k = 0; cu.c = 0;
}
// move to the first chunk at or after (ph, s) that exists; returns false at the end of the token
__device__ inline bool mg_cursor_seek(MgCursor& cu, const MegaPhase* phases, int n_phases, int n_cta, int cta, int warp) {
    while (cu.ph < n_phases) {
        if (cu.s < phases[cu.ph].nseg) {
            mg_cursor_load(cu, phases, n_cta, cta, warp);
            if (cu.n > 0) return true;
            cu.s++;
        } else { cu.ph++; cu.s = 0; }
    }
    return false;
}
__device__ inline bool mg_cursor_next(MgCursor& cu, const MegaPhase* phases, int n_phases, int n_cta, int cta, int warp) {
    if (++cu.c < cu.cpp) return true;
    cu.c = 0;
    if (++cu.k < cu.n) return true;
    cu.s++;
    return mg_cursor_seek(cu, phases, n_phases, n_cta, cta, warp);
}

// one thread per (cta, warp): list == nullptr -> only count; else write descriptors {addr.lo, addr.hi, bytes, 0}
__global__ void mg_chunk_list_kernel(const MegaPhase* phases, int n_phases, int n_cta, uint4* list, int list_stride, int* counts) {
    const int id = blockIdx.x * blockDim.x + threadIdx.x;
    if (id >= n_cta * MG_WARPS) return;
    const int cta = id / MG_WARPS, warp = id % MG_WARPS;
    MgCursor cu; cu.ph = 0; cu.s = 0; cu.k = 0; cu.c = 0; cu.n = 0; cu.cpp = 1; cu.bytes = 0; cu.addr0 = nullptr; cu.item_stride = 0;
    bool more = mg_cursor_seek(cu, phases, n_phases, n_cta, cta, warp);
    int n = 0;
    while (more) {
        if (list) {
            const unsigned long long a = (unsigned long long)(cu.addr0 + (size_t)cu.k * cu.item_stride + (size_t)cu.c * cu.bytes);
            list[(size_t)id * list_stride + n] = make_uint4((uint32_t)a, (uint32_t)(a >> 32), cu.bytes, 0u);
        }
        n++;
        more = mg_cursor_next(cu, phases, n_phases, n_cta, cta, warp);
    }
    counts[id] = n;
}

// ---- grid barrier ---------------------------------------------------------------------------------------------------
__device__ __forceinline__ void mg_grid_arrive(unsigned int* bar) {
    __syncthreads();
    if (threadIdx.x == 0) { __threadfence(); atomicAdd(bar, 1u); }
}
__device__ __forceinline__ void mg_grid_wait(const unsigned int* bar, unsigned int target) {
    if (threadIdx.x == 0) { while (ld_acquire_u32(bar) < target) { } __threadfence(); }
    __syncthreads();
}

// ---- step 0: embedding row -> x (rows spread over the CTAs) -------------------------------------------------------------
__device__ __noinline__ void mg_embed(const MegaParams& P) {
    const int tid = threadIdx.x;
    const int tok = P.tok[0];
    const int per = (P.n_embd / 64 + P.n_cta - 1) / P.n_cta;      // 64-element units per CTA
    const int u0 = (int)blockIdx.x * per, u1 = min(P.n_embd / 64, u0 + per);
    const QMat& E = P.tok_embd;
    if (E.type == QT_Q4_K || E.type == QT_Q5_K || E.type == QT_Q6_K) {
        for (int u = u0 + tid; u < u1; u += MG_THREADS) {
            float v[64];
            if (E.type == QT_Q4_K) dequant_unit_q4k(E, tok, u, v); else if (E.type == QT_Q5_K) dequant_unit_q5k(E, tok, u, v); else dequant_unit_q6k(E, tok, u, v);
            if (E.type == QT_Q6_K) { for (int l = 0; l < 64; l++) P.x[q6k_unit_elem(u, l)] = v[l]; }
            else { for (int l = 0; l < 64; l++) P.x[u * 64 + l] = v[l]; }
        }
    } else if (E.type == QT_Q8_0) {
        for (int u = 2 * u0 + tid; u < 2 * u1; u += MG_THREADS) { float v[32]; dequant_unit_q80(E, tok, u, v); for (int l = 0; l < 32; l++) P.x[u * 32 + l] = v[l]; }
    } else if (E.type == QT_F32) {
        const float* src = reinterpret_cast<const float*>(E.p0) + (size_t)tok * E.K;
        for (int i = u0 * 64 + tid; i < u1 * 64; i += MG_THREADS) P.x[i] = src[i];
    } else {
        const __half* src = reinterpret_cast<const __half*>(E.p0) + (size_t)tok * E.K;
        for (int i = u0 * 64 + tid; i < u1 * 64; i += MG_THREADS) P.x[i] = __half2float(src[i]);
    }
}

// ---- attention of one layer, two grid-synchronised stages ------------------------------------------------------------------
// CTA c serves KV head c % n_head_kv, context split c / n_head_kv.
struct MgAttn { int hk, split, n_split, t0, nt; bool on; };
__device__ __forceinline__ MgAttn mg_attn_setup(const MegaParams& P, int n_kv) {
    MgAttn a;
    a.n_split = max(1, min(P.max_split, (n_kv + 63) / 64));
    a.hk = (int)blockIdx.x % P.n_head_kv; a.split = (int)blockIdx.x / P.n_head_kv;
    a.on = a.split < a.n_split;
    const int per = (n_kv + a.n_split - 1) / a.n_split;
    a.t0 = a.split * per;
    a.nt = max(0, min(n_kv, a.t0 + per) - a.t0);
    return a;
}
__device__ __forceinline__ size_t mg_kv_row(const MegaParams& P, int hk, int t) {
    return ((size_t)P.page_table[t / KV_PAGE] * KV_PAGE + (t % KV_PAGE)) * P.kv_dim + (size_t)hk * P.d_head;
}

// stage 1: scaled scores of this CTA's token slice for the gq query heads of its KV head -> global
__device__ __noinline__ void mg_attn_scores(const MegaParams& P, int layer, int n_kv, unsigned char* s_act) {
    const MgAttn a = mg_attn_setup(P, n_kv);
    if (!a.on || a.nt == 0) return;
    const int tid = threadIdx.x, dh = P.d_head, gq = P.n_head / P.n_head_kv;
    __half* sqh = reinterpret_cast<__half*>(s_act);                 // [gq][dh]
    for (int i = tid; i < gq * dh; i += MG_THREADS) sqh[i] = __float2half_rn(__ldcg(P.qbuf + (size_t)(a.hk * gq) * dh + i));
    __syncthreads();
    const __half* kp = P.k_pools[layer];
    const int qd = tid & 3, QD = dh >> 2;                           // 4 lanes per token, a quarter of the head dim each
    for (int tl0 = 0; tl0 < a.nt; tl0 += MG_THREADS / 4) {
        const int tl = tl0 + (tid >> 2);
        float s[MAX_GQ];
#pragma unroll
        for (int gI = 0; gI < MAX_GQ; gI++) s[gI] = 0.0f;
        if (tl < a.nt) {
            const uint4* kr = reinterpret_cast<const uint4*>(kp + mg_kv_row(P, a.hk, a.t0 + tl) + qd * QD);
            for (int c = 0; c < QD / 8; c++) {
                const uint4 kv = __ldcg(kr + c);
                const __half2* kh = reinterpret_cast<const __half2*>(&kv);
                float kf[8];
#pragma unroll
                for (int i = 0; i < 4; i++) { const float2 f = __half22float2(kh[i]); kf[2 * i] = f.x; kf[2 * i + 1] = f.y; }
#pragma unroll
                for (int gI = 0; gI < MAX_GQ; gI++) if (gI < gq) {
                    const __half2* qh = reinterpret_cast<const __half2*>(sqh + gI * dh + qd * QD + c * 8);
#pragma unroll
                    for (int i = 0; i < 4; i++) { const float2 f = __half22float2(qh[i]); s[gI] += kf[2 * i] * f.x; s[gI] += kf[2 * i + 1] * f.y; }
                }
            }
        }
#pragma unroll
        for (int gI = 0; gI < MAX_GQ; gI++) if (gI < gq) {
            float v = s[gI];
            v += __shfl_xor_sync(0xffffffffu, v, 1);
            v += __shfl_xor_sync(0xffffffffu, v, 2);
            if (tl < a.nt && qd == 0) P.scores[(size_t)(a.hk * gq + gI) * P.score_stride + a.t0 + tl] = __fmul_rn(v, P.attn_scale);
        }
    }
}

// stage 2: soft-max statistics over the whole context (redundantly per CTA), probabilities of the own slice rounded to f16,
// partial V.p; the last CTA of the KV head to finish sums the split partials in split order.
__device__ __noinline__ void mg_attn_pv(const MegaParams& P, int layer, int n_kv, unsigned char* s_act, float* s_redf, double* s_redd, float* s_stat) {
    const MgAttn a = mg_attn_setup(P, n_kv);
    if (!a.on) return;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, dh = P.d_head, gq = P.n_head / P.n_head_kv;
    float* sp_p = reinterpret_cast<float*>(s_act);                           // [gq][MG_PCAP] probabilities (f16-rounded)
    float* s_red = sp_p + gq * MG_PCAP;                                      // [TG][gq][dh] partial outputs
    const int TG = MG_THREADS / dh;                                          // token groups
    const int d = tid % dh, tg = tid / dh;
    float acc[MAX_GQ];
#pragma unroll
    for (int gI = 0; gI < MAX_GQ; gI++) acc[gI] = 0.0f;
    if (a.nt > 0) {
        float M[MAX_GQ], inv[MAX_GQ];
#pragma unroll
        for (int gI = 0; gI < MAX_GQ; gI++) if (gI < gq) {
            const float* sr = P.scores + (size_t)(a.hk * gq + gI) * P.score_stride;
            float mx = -INFINITY;
            for (int t = tid; t < n_kv; t += MG_THREADS) mx = fmaxf(mx, __ldcg(sr + t));
            mx = warp_max(mx);
            if (lane == 0) s_redf[gI * MG_WARPS + warp] = mx;
        }
        __syncthreads();
#pragma unroll
        for (int gI = 0; gI < MAX_GQ; gI++) if (gI < gq) {
            float mx = s_redf[gI * MG_WARPS];
            for (int w = 1; w < MG_WARPS; w++) mx = fmaxf(mx, s_redf[gI * MG_WARPS + w]);
            M[gI] = mx;
        }
#pragma unroll
        for (int gI = 0; gI < MAX_GQ; gI++) if (gI < gq) {
            const float* sr = P.scores + (size_t)(a.hk * gq + gI) * P.score_stride;
            double sum = 0.0;
            for (int t = tid; t < n_kv; t += MG_THREADS) sum += (double)expf(__ldcg(sr + t) - M[gI]);
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
            if (lane == 0) s_redd[gI * MG_WARPS + warp] = sum;
        }
        __syncthreads();
#pragma unroll
        for (int gI = 0; gI < MAX_GQ; gI++) if (gI < gq) {
            double sum = 0.0;
            for (int w = 0; w < MG_WARPS; w++) sum += s_redd[gI * MG_WARPS + w];
            inv[gI] = (float)(1.0 / sum);
        }
        const __half* vp = P.v_pools[layer];
        for (int c0 = 0; c0 < a.nt; c0 += MG_PCAP) {
            const int cn = min(MG_PCAP, a.nt - c0);
            __syncthreads();
            for (int i = tid; i < gq * cn; i += MG_THREADS) {
                const int gI = i / cn, tl = i - gI * cn;
                const float sc = __ldcg(P.scores + (size_t)(a.hk * gq + gI) * P.score_stride + a.t0 + c0 + tl);
                float m = M[0], iv = inv[0];
#pragma unroll
                for (int g2 = 1; g2 < MAX_GQ; g2++) if (g2 == gI) { m = M[g2]; iv = inv[g2]; }
                sp_p[gI * MG_PCAP + tl] = __half2float(__float2half_rn(__fmul_rn(expf(sc - m), iv)));
            }
            __syncthreads();
            for (int tl = tg; tl < cn; tl += TG) {
                const float v = __half2float(__ushort_as_half(__ldcg(reinterpret_cast<const unsigned short*>(vp + mg_kv_row(P, a.hk, a.t0 + c0 + tl) + d))));
#pragma unroll
                for (int gI = 0; gI < MAX_GQ; gI++) if (gI < gq) acc[gI] += sp_p[gI * MG_PCAP + tl] * v;
            }
        }
    }
    // token groups meet in shared memory, summed in fixed order
    __syncthreads();
#pragma unroll
    for (int gI = 0; gI < MAX_GQ; gI++) if (gI < gq) s_red[(tg * gq + gI) * dh + d] = acc[gI];
    __syncthreads();
    for (int e = tid; e < gq * dh; e += MG_THREADS) {
        float o = 0.0f;
        for (int k = 0; k < TG; k++) o += s_red[k * gq * dh + e];
        P.part_o[((size_t)(a.hk * gq * dh + e)) * P.max_split + a.split] = o;
    }
    // the last CTA of this KV head sums the split partials in split order
    __threadfence();
    __syncthreads();
    if (tid == 0) {
        const unsigned int old = atomicAdd(P.sync + 4 + a.hk, 1u);
        const bool last = old == (unsigned int)(a.n_split - 1);
        s_stat[0] = last ? 1.0f : 0.0f;
        if (last) P.sync[4 + a.hk] = 0u;
    }
    __syncthreads();
    if (s_stat[0] != 0.0f) {
        __threadfence();
        for (int e = tid; e < gq * dh; e += MG_THREADS) {
            const float* pp = P.part_o + ((size_t)(a.hk * gq * dh + e)) * P.max_split;
            float o = 0.0f;
            for (int s = 0; s < a.n_split; s++) o += __ldcg(pp + s);
            P.attn_out[(size_t)a.hk * gq * dh + e] = o;
        }
    }
}

// ---- epilogue of a stream phase: combine the K-slice partials of every row pair of this CTA and finish the rows ---------
__device__ __noinline__ void mg_epilogue(const MegaParams& P, const MegaPhase* ph, int pos, const float2* s_part, const float2* s_rope) {
    const int tid = threadIdx.x, dh = P.d_head;
    const MgGroup g = mg_group(ph, P.n_cta, (int)blockIdx.x, 0);
    for (int e = tid; e < g.NG * P.max_items; e += MG_THREADS) {
        const int gl = e % g.NG, ks = e / g.NG;
        const int gg = (int)blockIdx.x * g.NG + gl;
        int rem = ks, s = 0, vg = 0, n = 0;
        const MegaSeg* sg = nullptr;
        for (; s < ph->nseg; s++) {
            sg = ph->seg + s;
            vg = (gg + sg->rot) % g.NGtot;
            n = vg < sg->n_pairs ? (sg->n_pairs - 1 - vg) / g.NGtot + 1 : 0;
            if (rem < n) break;
            rem -= n;
        }
        if (s >= ph->nseg) continue;
        const int p = vg + rem * g.NGtot;
        float v0 = 0.0f, v1 = 0.0f;
        for (int w = 0; w < g.W; w++) { const float2 t = s_part[ks * MG_WARPS + gl * g.W + w]; v0 += t.x; v1 += t.y; }
        const int kind = sg->kind;
        if (kind == MK_SWIGLU) { P.hbuf[p] = (v0 / (1.0f + expf(-v0))) * v1; continue; }     // ggml_silu_f32 then ggml_mul
        int r0 = 2 * p, r1 = 2 * p + 1;
        if ((kind == MK_Q || kind == MK_K) && P.neox) { const int hd = dh >> 1; r0 = (p / hd) * dh + (p % hd); r1 = r0 + hd; }
        if (sg->bias) { v0 += sg->bias[r0]; v1 += sg->bias[r1]; }
        if (kind == MK_RESID) { P.x[r0] = __ldcg(P.x + r0) + v0; P.x[r1] = __ldcg(P.x + r1) + v1; }
        else if (kind == MK_LOGITS) {
            P.logits[r0] = v0; P.logits[r1] = v1;
            atomicMax(P.chunk_max + (r0 >> P.chunk_shift), float_order_key(fmaxf(v0, v1)));
        } else {
            if (kind != MK_V) {                                   // rotary embedding on the pair: ggml rope NORM / NEOX
                const int i = P.neox ? (r0 % dh) : ((r0 % dh) >> 1);
                const float2 cs = s_rope[i];
                const float x0 = v0, x1 = v1;
                v0 = x0 * cs.x - x1 * cs.y;
                v1 = x0 * cs.y + x1 * cs.x;
            }
            if (kind == MK_Q) { P.qbuf[r0] = v0; P.qbuf[r1] = v1; }
            else {
                const size_t base = ((size_t)P.page_table[pos / KV_PAGE] * KV_PAGE + (pos % KV_PAGE)) * P.kv_dim;
                __half* dst = (kind == MK_K ? P.k_pools : P.v_pools)[ph->layer];
                dst[base + r0] = __float2half_rn(v0);           // ggml_cpy f32 -> f16 into the cache
                dst[base + r1] = __float2half_rn(v1);
            }
        }
    }
}

// ---- prologue of a stream phase: (RMSNorm *) quantise the source vector into shared memory, redundantly per CTA -------------
__device__ __noinline__ void mg_prologue(const MegaParams& P, const MegaPhase* ph, unsigned char* s_act, double* s_redd) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int K = ph->K, fmt = ph->act_fmt;
    int8_t* sq = reinterpret_cast<int8_t*>(s_act);
    float* sd = reinterpret_cast<float*>(s_act + K);
    int16_t* sbs = reinterpret_cast<int16_t*>(s_act + K + (K >> 5) * 4);
    const float* src = ph->src == MSRC_X ? P.x : (ph->src == MSRC_ATTN ? P.attn_out : P.hbuf);
    float scale = 1.0f;
    if (ph->norm_w) {
        double sum = 0.0;
        for (int i = tid; i < (K >> 2); i += MG_THREADS) {
            const float4 v = __ldcg(reinterpret_cast<const float4*>(src) + i);
            sum += (double)__fmul_rn(v.x, v.x); sum += (double)__fmul_rn(v.y, v.y); sum += (double)__fmul_rn(v.z, v.z); sum += (double)__fmul_rn(v.w, v.w);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
        if (lane == 0) s_redd[warp] = sum;
        __syncthreads();
        double tot = 0.0;
#pragma unroll
        for (int i = 0; i < MG_WARPS; i++) tot += s_redd[i];
        const float mean = (float)(tot / (double)K);
        scale = __fdiv_rn(1.0f, __fsqrt_rn(__fadd_rn(mean, P.eps)));
    }
    for (int b = warp; b < (K >> 8); b += MG_WARPS) {
        if (ph->norm_w) mg_quantize_block<true>(src + b * 256, fmt, b, sq, sd, sbs, scale, ph->norm_w + b * 256);
        else mg_quantize_block<false>(src + b * 256, fmt, b, sq, sd, sbs, 1.0f, nullptr);
    }
    __syncthreads();
}

// =================================================================================================================
// the kernel
// =================================================================================================================
__global__ void __launch_bounds__(MG_THREADS, 1) mega_decode_kernel(const __grid_constant__ MegaParams P) {
    extern __shared__ __align__(128) unsigned char smem[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

    // ---- shared memory carve-up ----
    unsigned char* ring = smem + (size_t)warp * MG_SLOTS * P.slot_bytes;
    unsigned char* sp = smem + (size_t)MG_WARPS * MG_SLOTS * P.slot_bytes;
    uint64_t* my_bar = reinterpret_cast<uint64_t*>(sp) + warp * MG_SLOTS; sp += MG_WARPS * MG_SLOTS * 8;
    float2* s_part = reinterpret_cast<float2*>(sp); sp += (size_t)P.max_items * MG_WARPS * sizeof(float2);
    float2* s_rope = reinterpret_cast<float2*>(sp); sp += (size_t)(P.d_head / 2) * sizeof(float2);
    double* s_redd = reinterpret_cast<double*>(sp); sp += 8 * MG_WARPS * sizeof(double);
    float* s_redf = reinterpret_cast<float*>(sp); sp += 8 * MG_WARPS * sizeof(float);
    float* s_stat = reinterpret_cast<float*>(sp); sp += 32 * sizeof(float);
    unsigned char* s_act = sp;                                      // P.act_bytes: activations | attention scratch

    if (lane == 0) { for (int s = 0; s < MG_SLOTS; s++) mbar_init(my_bar + s, 1); mbar_fence_init(); }
    __syncwarp();

    // ---- producer: this warp's chunk list; prime the ring (weights do not depend on anything) ----
    const uint4* plist = P.chunk_list + ((size_t)blockIdx.x * MG_WARPS + warp) * P.list_stride;
    const int p_total = P.chunk_counts[blockIdx.x * MG_WARPS + warp];
    int p_issued = 0, c_done = 0;
    uint4 p_desc = p_total > 0 ? __ldg(plist) : make_uint4(0, 0, 0, 0);
    auto issue_next = [&]() {
        if (p_issued < p_total) {
            if (lane == 0) {
                const int slot = p_issued % MG_SLOTS;
                mbar_expect_tx(my_bar + slot, p_desc.z);
                bulk_g2s(ring + (size_t)slot * P.slot_bytes, reinterpret_cast<const void*>(((unsigned long long)p_desc.y << 32) | p_desc.x), p_desc.z, my_bar + slot);
            }
            p_issued++;
            if (p_issued < p_total) p_desc = __ldg(plist + p_issued);
        }
    };
    for (int s = 0; s < MG_SLOTS; s++) issue_next();

    const int pos = P.pos[0];
    const int n_kv = pos + 1;
    unsigned int bar_target = 0;
    // optional per-CTA event trace (SM clock of thread 0 at every stage boundary), see tools/mega_trace.py
    int tr_i = 0;
    auto TR = [&]() { if (P.trace && tid == 0) { if (tr_i < P.trace_cap) P.trace[(size_t)blockIdx.x * P.trace_cap + tr_i] = clock64(); tr_i++; } };
    auto grid_sync = [&]() { mg_grid_arrive(P.sync); bar_target += (unsigned int)P.n_cta; mg_grid_wait(P.sync, bar_target); };

    TR();
    mg_embed(P);
    rope_table_fill(s_rope, P.d_head / 2, pos, P.theta_scale, P.rope_freqs);
    grid_sync();
    TR();

    // ================================= one stream phase (mat-vec group) =================================
    auto run_phase = [&](int phi) {
        const MegaPhase* ph = P.phases + phi;
        mg_prologue(P, ph, s_act, s_redd);
        TR();
        const MgGroup g = mg_group(ph, P.n_cta, (int)blockIdx.x, warp);
        const int K = ph->K, fmt = ph->act_fmt;
        // -- this lane's activations --
        const int sub = lane / g.L;                       // row of the chunk this lane works on
        const int hsl = lane - sub * g.L;                 // half super-block inside the slice
        const int hs = g.ws * g.L + hsl;                  // ... inside the row
        const bool lane_on = g.active && sub < g.rpc && hs < (K >> 7);
        LaneAct A;
        if (lane_on) {
            const int8_t* sq = reinterpret_cast<const int8_t*>(s_act);
            const float* sd = reinterpret_cast<const float*>(s_act + K);
            const int16_t* sbs = reinterpret_cast<const int16_t*>(s_act + K + (K >> 5) * 4);
#pragma unroll
            for (int c = 0; c < 8; c++) {
                const uint4 t = lds128(sq + hs * 128 + ((c ^ (hs & 7)) << 4));
                A.q[4 * c] = (int)t.x; A.q[4 * c + 1] = (int)t.y; A.q[4 * c + 2] = (int)t.z; A.q[4 * c + 3] = (int)t.w;
            }
            if (fmt == ACT_Q8_K) {
                A.d0 = sd[hs >> 1];
                const uint4 t = lds128(sbs + hs * 8);
                A.aux[0] = t.x; A.aux[1] = t.y; A.aux[2] = t.z; A.aux[3] = t.w;
                auto pair_sum = [](uint32_t w) -> int { return (int)(int16_t)(w & 0xffff) + (int)(int16_t)(w >> 16); };
                A.aux[4] = (uint32_t)(pair_sum(t.x) & 0xffff) | ((uint32_t)pair_sum(t.y) << 16);
                A.aux[5] = (uint32_t)(pair_sum(t.z) & 0xffff) | ((uint32_t)pair_sum(t.w) << 16);
            } else {
                A.d0 = 0.0f;
#pragma unroll
                for (int b = 0; b < 4; b++) A.aux[b] = __float_as_uint(sd[hs * 4 + b]);
                A.aux[4] = A.aux[5] = 0u;
            }
        } else {
#pragma unroll
            for (int i = 0; i < 32; i++) A.q[i] = 0;
            A.d0 = 0.0f;
#pragma unroll
            for (int i = 0; i < 6; i++) A.aux[i] = 0u;
        }
        // -- main loop: every chunk of this warp in this phase --
        const int sb = hsl >> 1, hf = hsl & 1;
        const int cpp = 2 / g.rpc;
        int slot_i = 0;                                   // running item index of this group inside the phase
        for (int s = 0; s < ph->nseg; s++) {
            const MegaSeg* sg = ph->seg + s;
            int vg, n; mg_seg_items(g, sg, vg, n);
            const int type = sg->type;
            const int sub_off = sub * sg->slice_bytes;
            const int tail_off = (g.L >> 1) * 208 + sb * 2;     // Q6_K: this super-block's d in the slice tail
            for (int k = 0; k < n; k++, slot_i++) {
                float v0 = 0.0f, v1 = 0.0f;
                for (int c = 0; c < cpp; c++) {
                    const int slot = c_done % MG_SLOTS;
                    mbar_wait(my_bar + slot, (uint32_t)((c_done / MG_SLOTS) & 1));
                    const unsigned char* sl = ring + (size_t)slot * P.slot_bytes + sub_off;
                    float v = 0.0f;
                    if (lane_on) {
                        uint4 wb[9];
                        if (type == QT_Q4_K) {
                            const unsigned char* b = sl + sb * 144;
                            wb[0] = lds128(b);
#pragma unroll
                            for (int i = 0; i < 4; i++) wb[1 + i] = lds128(b + 16 + hf * 64 + 16 * i);
                            v = mg_dot_q4k(wb, hf, A);
                        } else if (type == QT_Q6_K) {
                            const unsigned char* b = sl + sb * 208;
#pragma unroll
                            for (int i = 0; i < 2; i++) { wb[i] = lds128(b + hf * 64 + 16 * i); wb[2 + i] = lds128(b + hf * 64 + 32 + 16 * i); }
#pragma unroll
                            for (int i = 0; i < 2; i++) wb[4 + i] = lds128(b + 128 + hf * 32 + 16 * i);
                            const uint2 sc = *reinterpret_cast<const uint2*>(b + 192 + hf * 8);
                            wb[6].x = sc.x; wb[6].y = sc.y;
                            wb[6].z = *reinterpret_cast<const unsigned short*>(sl + tail_off);
                            v = mg_dot_q6k(wb, A);
                        } else if (type == QT_Q8_0) {
                            const unsigned char* b = sl + sb * 272;
#pragma unroll
                            for (int i = 0; i < 8; i++) wb[i] = lds128(b + hf * 128 + 16 * i);
                            const uint2 dd = *reinterpret_cast<const uint2*>(b + 256 + hf * 8);
                            wb[8].x = dd.x; wb[8].y = dd.y;
                            v = mg_dot_q80(wb, A);
                        } else {   // QT_Q5_K
                            const unsigned char* b = sl + sb * 176;
                            wb[0] = lds128(b); wb[1] = lds128(b + 16); wb[2] = lds128(b + 32);
#pragma unroll
                            for (int i = 0; i < 4; i++) wb[3 + i] = lds128(b + 48 + hf * 64 + 16 * i);
                            v = mg_dot_q5k(wb, hf, A);
                        }
                    }
                    __syncwarp();                         // every lane has read the slot: refill it
                    c_done++;
                    issue_next();
                    if (g.rpc == 2) { v0 = warp_sum(sub == 0 ? v : 0.0f); v1 = warp_sum(sub == 1 ? v : 0.0f); }
                    else if (c == 0) v0 = warp_sum(v); else v1 = warp_sum(v);
                }
                if (lane == 0) s_part[slot_i * MG_WARPS + warp] = make_float2(v0, v1);
            }
        }
        __syncthreads();
        TR();
        mg_epilogue(P, ph, pos, s_part, s_rope);
        TR();
    };

    // ================================= the token =================================
    int phase = 0;
    for (int l = 0; l < P.n_layer; l++) {
        run_phase(phase++);            // QKV (+bias, RoPE, KV write)
        grid_sync(); TR();
        mg_attn_scores(P, l, n_kv, s_act);
        TR(); grid_sync(); TR();
        mg_attn_pv(P, l, n_kv, s_act, s_redf, s_redd, s_stat);
        TR(); grid_sync(); TR();
        run_phase(phase++);            // Wo + residual
        grid_sync(); TR();
        run_phase(phase++);            // gate / up + SwiGLU
        grid_sync(); TR();
        run_phase(phase++);            // down + residual
        grid_sync(); TR();
    }
    if (P.with_head) { run_phase(phase++); grid_sync(); TR(); }

    // ---- exit: the last CTA out resets the barrier state for the next launch ----
    if (tid == 0) {
        if (blockIdx.x == 0 && P.advance_pos) P.pos[0] = pos + 1;
        const unsigned int old = atomicAdd(P.sync + 1, 1u);
        if (old == (unsigned int)(P.n_cta - 1)) { P.sync[0] = 0u; P.sync[1] = 0u; __threadfence(); }
    }
}

} // namespace blk
