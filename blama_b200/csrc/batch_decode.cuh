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
    pdl_launch_dependents(); pdl_wait();
    const int t = blockIdx.x;
    qkv_post_row(b.base, t, b.pos[t], b.k_pools[t][b.layer], b.v_pools[t][b.layer], b.page_table[t]);
}

// attention of every row over its own cache: grid (n_head_kv, n), 256 threads.  Tiles of BD_TK tokens (flash order, everything in
// f32; K and V are the f16 cache rows):
//   scores    8 lanes per token, DH / 8 dims each, reduced by three shuffles (32 tokens per pass; q broadcast from shared memory)
//   soft-max  tile maximum / sum per query head, running maximum across tiles
//   P.V       thread = (dimension pair, token group): 256 / (DH / 2) token groups walk the tile side by side, every thread carries all
//             query heads of the KV head; the groups are added through shared memory at the end
// (The first version gave a whole token to one thread in the score stage and the whole tile to one thread per (head, dimension pair) in
// P.V: 29.6 us per layer for 32 rows of ~100 tokens, 14 % of a batched step.)
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
    constexpr int TG = 256 / HP;                     // token groups of the P.V stage (4 for d_head 128, 8 for 64)
    constexpr int DL = DH / 8;                       // dims per lane of the score stage
    __shared__ __align__(16) float q_s[MAX_GQ * DH];
    __shared__ float s_s[MAX_GQ][BD_TK];
    __shared__ int off_s[BD_TK];                     // cache row offset (in halfs) of every token of the tile
    __shared__ float red[MAX_GQ][8];
    __shared__ float m_run[MAX_GQ], l_run[MAX_GQ], corr[MAX_GQ];
    __shared__ float2 pv_red[TG][MAX_GQ][HP];
    pdl_launch_dependents(); pdl_wait();
    const int hk = blockIdx.x, b = blockIdx.y, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int gq = a.n_head / a.n_head_kv, n_kv = a.pos[b] + 1;
    const int32_t* pt = a.page_table[b];
    const __half* kp = a.k_pools[b][a.layer] + (size_t)hk * DH;
    const __half* vp = a.v_pools[b][a.layer] + (size_t)hk * DH;
    for (int i = tid; i < gq * DH; i += 256) q_s[i] = __half2float(a.q[(size_t)b * a.n_head * DH + (size_t)hk * gq * DH + i]);
    if (tid < MAX_GQ) { m_run[tid] = -INFINITY; l_run[tid] = 0.0f; }
    const int dp = tid % HP, tg = tid / HP;          // P.V stage: this thread's dimension pair and token group
    const int ld = tid & 7;                          // score stage: this lane's DL dims of the token tid >> 3 of the pass
    float2 acc[MAX_GQ];
#pragma unroll
    for (int g = 0; g < MAX_GQ; g++) acc[g] = make_float2(0.0f, 0.0f);
    __syncthreads();
    for (int t0 = 0; t0 < n_kv; t0 += BD_TK) {
        const int cn = min(BD_TK, n_kv - t0);
        // scores of the tile
        for (int j0 = 0; j0 < cn; j0 += 32) {
            const int j = j0 + (tid >> 3);
            float sc[MAX_GQ];
#pragma unroll
            for (int g = 0; g < MAX_GQ; g++) sc[g] = 0.0f;
            if (j < cn) {
                const int t = t0 + j;
                const int off = (pt[t / KV_PAGE] * KV_PAGE + (t % KV_PAGE)) * a.kv_dim;
                if (ld == 0) off_s[j] = off;
                const uint4* kr = reinterpret_cast<const uint4*>(kp + off + ld * DL);
#pragma unroll
                for (int c = 0; c < DL / 8; c++) {
                    const uint4 kv = kr[c];
                    const __half2* kh = reinterpret_cast<const __half2*>(&kv);
                    float kf[8];
#pragma unroll
                    for (int i = 0; i < 4; i++) { const float2 f = __half22float2(kh[i]); kf[2 * i] = f.x; kf[2 * i + 1] = f.y; }
#pragma unroll
                    for (int g = 0; g < MAX_GQ; g++) if (g < gq) {
                        const float4 q0 = *reinterpret_cast<const float4*>(q_s + g * DH + ld * DL + c * 8);
                        const float4 q1 = *reinterpret_cast<const float4*>(q_s + g * DH + ld * DL + c * 8 + 4);
                        sc[g] += kf[0] * q0.x + kf[1] * q0.y + kf[2] * q0.z + kf[3] * q0.w + kf[4] * q1.x + kf[5] * q1.y + kf[6] * q1.z + kf[7] * q1.w;
                    }
                }
            }
#pragma unroll
            for (int g = 0; g < MAX_GQ; g++) if (g < gq) {
                float v = sc[g];
                v += __shfl_xor_sync(0xffffffffu, v, 1);
                v += __shfl_xor_sync(0xffffffffu, v, 2);
                v += __shfl_xor_sync(0xffffffffu, v, 4);
                if (ld == 0 && j < cn) s_s[g][j] = v * a.scale;
            }
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
        // P.V of the tile: token group tg takes tokens tg, tg + TG, ...; V rows are read as half2, coalesced over dp
#pragma unroll
        for (int g = 0; g < MAX_GQ; g++) if (g < gq) { acc[g].x *= corr[g]; acc[g].y *= corr[g]; }
#pragma unroll 4
        for (int j = tg; j < cn; j += TG) {
            const float2 v = __half22float2(*reinterpret_cast<const __half2*>(vp + off_s[j] + 2 * dp));
#pragma unroll
            for (int g = 0; g < MAX_GQ; g++) if (g < gq) { const float p = s_s[g][j]; acc[g].x += p * v.x; acc[g].y += p * v.y; }
        }
        __syncthreads();
    }
#pragma unroll
    for (int g = 0; g < MAX_GQ; g++) if (g < gq) pv_red[tg][g][dp] = acc[g];
    __syncthreads();
    for (int e = tid; e < gq * HP; e += 256) {
        const int g = e / HP, d2 = e % HP;
        float2 o = make_float2(0.0f, 0.0f);
#pragma unroll
        for (int k = 0; k < TG; k++) { const float2 t = pv_red[k][g][d2]; o.x += t.x; o.y += t.y; }
        const float inv = 1.0f / l_run[g];
        *reinterpret_cast<__nv_bfloat162*>(a.out + (size_t)b * a.n_head * DH + (size_t)(hk * gq + g) * DH + 2 * d2) = __floats2bfloat162_rn(o.x * inv, o.y * inv);
    }
}

// every sequence's device position counter moves on by one
__global__ void advance_many_kernel(int32_t* const* pos_ptrs, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) pos_ptrs[i][0] += 1;
}

} // namespace blk
