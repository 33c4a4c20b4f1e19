// batch_decode.cuh -- kernels of the batched decode step (blk_decode_batch): ONE forward pass for n sequences, one new token each.
//
// Continuous batching is not in the reference (its Server serialises requests, server/code/server/Server.cpp:36); SURVEY.md lists
// it as the follow-up that turns the batch-1 mat-vec into a skinny GEMM (section 8f item 4).  The step reuses the prefill machinery --
// tcgen05 GEMMs with the weights de-quantised inside the kernel (prefill_gemm.cuh, bf16 operands, f32 accumulation), RMSNorm,
// SwiGLU epilogue -- and adds what a batch of INDEPENDENT sequences needs: per-row positions for the rotary table, K / V rows
// written into every sequence's own paged cache, and attention of every row over its own cache.
#pragma once
#include "prefill_kernels.cuh"

namespace blk {

constexpr int BATCH_MAX = 64;          // sequences per batched step

// embedding rows + the rotary table of every row's own position
__global__ void __launch_bounds__(256) embed_rows_kernel(QMat E, const int32_t* __restrict__ tokens, const int32_t* __restrict__ pos, float* x,
                                                         float2* rope_cs, int half_rot, float theta_scale, const float* freq_factors) {
    const int t = blockIdx.x;
    dequant_row_cta(E, tokens[t], x + (size_t)t * E.K);
    rope_table_fill(rope_cs + (size_t)t * half_rot, half_rot, pos[t], theta_scale, freq_factors);
}

// RoPE on q, k; q -> f16 [n][dq]; k, v -> the f16 cache row of position pos[t] in sequence t's own pools (one CTA per row)
struct QkvPostBatchArgs {
    QkvPostArgs base;                              // qkv, ld, rope_cs, q_out, dims (pos0 / pools / page_table unused)
    const int32_t* pos;                            // [n]
    const int32_t* const* page_table;              // [n]
    __half* const* const* k_pools;                 // [n] -> [n_layer]
    __half* const* const* v_pools;
    int layer;
};
__global__ void __launch_bounds__(256) qkv_post_batch_kernel(const QkvPostBatchArgs b) {
    const int t = blockIdx.x;
    qkv_post_row(b.base, t, b.pos[t], b.k_pools[t][b.layer], b.v_pools[t][b.layer], b.page_table[t]);
}

// attention of every row over its own cache: grid (n_head_kv, n), 256 threads.  Tiles of BD_TK tokens: scores with one thread per
// token (q broadcast from shared memory), tile maximum / sum per query head, P.V with one thread per (query head, dimension) pair,
// running rescale across tiles (flash order, everything in f32; K and V are the f16 cache rows).
constexpr int BD_TK = 512;
struct BatchAttnArgs {
    const __half* q;                               // [n][n_head * DH]
    const int32_t* pos;                            // [n]: the new token's position; the row attends to cells 0 .. pos
    const int32_t* const* page_table;
    __half* const* const* k_pools; __half* const* const* v_pools;
    __nv_bfloat16* out;                            // [n][n_head * DH]
    int layer, n_head, n_head_kv, kv_dim;
    float scale;
};
template <int DH>
__global__ void __launch_bounds__(256) decode_attn_batch_kernel(const BatchAttnArgs a) {
    constexpr int HP = DH / 2;                       // dimension pairs per head
    constexpr int SLOTS = 256 / HP;                  // query heads served per pass of the P.V stage (4 for d_head 128, 8 for 64)
    constexpr int NH = (MAX_GQ + SLOTS - 1) / SLOTS; // passes
    __shared__ float q_s[MAX_GQ * DH];
    __shared__ float s_s[MAX_GQ][BD_TK];
    __shared__ int off_s[BD_TK];                     // cache row offset (in halfs) of every token of the tile
    __shared__ float red[MAX_GQ][8];
    __shared__ float m_run[MAX_GQ], l_run[MAX_GQ], corr[MAX_GQ];
    const int hk = blockIdx.x, b = blockIdx.y, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int gq = a.n_head / a.n_head_kv, n_kv = a.pos[b] + 1;
    const int32_t* pt = a.page_table[b];
    const __half* kp = a.k_pools[b][a.layer] + (size_t)hk * DH;
    const __half* vp = a.v_pools[b][a.layer] + (size_t)hk * DH;
    for (int i = tid; i < gq * DH; i += 256) q_s[i] = __half2float(a.q[(size_t)b * a.n_head * DH + (size_t)hk * gq * DH + i]);
    if (tid < MAX_GQ) { m_run[tid] = -INFINITY; l_run[tid] = 0.0f; }
    const int dp = tid % HP, g0 = tid / HP;          // this thread's dimension pair and first query head
    float2 acc[NH];
#pragma unroll
    for (int i = 0; i < NH; i++) acc[i] = make_float2(0.0f, 0.0f);
    __syncthreads();
    for (int t0 = 0; t0 < n_kv; t0 += BD_TK) {
        const int cn = min(BD_TK, n_kv - t0);
        // scores of the tile: one thread per token
        for (int j = tid; j < cn; j += 256) {
            const int t = t0 + j;
            const int off = (pt[t / KV_PAGE] * KV_PAGE + (t % KV_PAGE)) * a.kv_dim;
            off_s[j] = off;
            const uint4* kr = reinterpret_cast<const uint4*>(kp + off);
            float sc[MAX_GQ];
#pragma unroll
            for (int g = 0; g < MAX_GQ; g++) sc[g] = 0.0f;
#pragma unroll 4
            for (int c = 0; c < DH / 8; c++) {
                const uint4 kv = kr[c];
                const __half2* kh = reinterpret_cast<const __half2*>(&kv);
                float kf[8];
#pragma unroll
                for (int i = 0; i < 4; i++) { const float2 f = __half22float2(kh[i]); kf[2 * i] = f.x; kf[2 * i + 1] = f.y; }
#pragma unroll
                for (int g = 0; g < MAX_GQ; g++) if (g < gq) {
#pragma unroll
                    for (int i = 0; i < 8; i++) sc[g] += kf[i] * q_s[g * DH + c * 8 + i];
                }
            }
#pragma unroll
            for (int g = 0; g < MAX_GQ; g++) if (g < gq) s_s[g][j] = sc[g] * a.scale;
        }
        __syncthreads();
        // tile maximum per head -> new running maximum, correction of what was accumulated so far
        for (int g = 0; g < gq; g++) {
            float m = -INFINITY;
            for (int j = tid; j < cn; j += 256) m = fmaxf(m, s_s[g][j]);
            m = warp_max(m);
            if (lane == 0) red[g][warp] = m;
        }
        __syncthreads();
        if (tid < gq) {
            float m = m_run[tid];
            for (int w = 0; w < 8; w++) m = fmaxf(m, red[tid][w]);
            corr[tid] = (m_run[tid] == -INFINITY) ? 0.0f : expf(m_run[tid] - m);
            m_run[tid] = m;
        }
        __syncthreads();
        for (int g = 0; g < gq; g++) {
            const float m = m_run[g];
            float l = 0.0f;
            for (int j = tid; j < cn; j += 256) { const float p = expf(s_s[g][j] - m); s_s[g][j] = p; l += p; }
            l = warp_sum(l);
            if (lane == 0) red[g][warp] = l;
        }
        __syncthreads();
        if (tid < gq) {
            float l = 0.0f;
            for (int w = 0; w < 8; w++) l += red[tid][w];
            l_run[tid] = l_run[tid] * corr[tid] + l;
        }
        // P.V of the tile: a thread owns one dimension pair of up to NH query heads; V rows are read as half2, coalesced over dp
#pragma unroll
        for (int i = 0; i < NH; i++) {
            const int g = g0 + i * SLOTS;
            if (g < gq) {
                float2 o = make_float2(acc[i].x * corr[g], acc[i].y * corr[g]);
                const float* pr = s_s[g];
#pragma unroll 4
                for (int j = 0; j < cn; j++) {
                    const float2 v = __half22float2(*reinterpret_cast<const __half2*>(vp + off_s[j] + 2 * dp));
                    const float p = pr[j];
                    o.x += p * v.x; o.y += p * v.y;
                }
                acc[i] = o;
            }
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < NH; i++) {
        const int g = g0 + i * SLOTS;
        if (g < gq) {
            const float inv = 1.0f / l_run[g];
            __nv_bfloat162 o = __floats2bfloat162_rn(acc[i].x * inv, acc[i].y * inv);
            *reinterpret_cast<__nv_bfloat162*>(a.out + (size_t)b * a.n_head * DH + (size_t)(hk * gq + g) * DH + 2 * dp) = o;
        }
    }
}

// every sequence's device position counter moves on by one
__global__ void advance_many_kernel(int32_t* const* pos_ptrs, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) pos_ptrs[i][0] += 1;
}

} // namespace blk
