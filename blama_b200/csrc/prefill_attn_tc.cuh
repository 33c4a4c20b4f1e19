// prefill_attn_tc.cuh -- causal attention of a prefill chunk on the 5th-generation tensor cores (tcgen05 + TMEM + TMA).
//
// CTA = (tile of BQ = 128 / GQ tokens, KV head): the GQ query heads of the KV head stacked head-major give the 128 rows of one
// UMMA M = 128 tile.  Keys are walked in tiles of 128.  All four operands are K-major SWIZZLE_128B tiles fetched by TMA:
//   Q  [128 rows][dh = 128]   from the f16 query buffer (GQ boxes of BQ rows per 64-wide K block)
//   K  [128 keys][128]        from the paged f16 cache (one 64-row box per page and K block)
//   V^T [128 dims][128 keys]  from a per-layer transposed scratch (vt_transpose_kernel): P.V needs V with the KEYS contiguous
//   P  [128 rows][128 keys]   written by the soft-max warps (f16, the same swizzle the GEMM producers write)
// ONE pass over the keys (round 2; the first version walked them twice -- row maxima, then probabilities -- and was bound by its
// K / V tile loads: 96 KB per key tile through a two-stage ring, one CTA per SM): S = Q.K^T in TMEM, P = exp2(s - m_ref) -> f16
// with the row sum alongside, O += P.V^T in TMEM.  m_ref is a per-row REFERENCE maximum, raised only when a tile's maximum exceeds
// it by more than 8 (in the exp2 domain): then O (TMEM, tcgen05.ld / st) and the row sum are rescaled by exp2(m_old - m_new) before
// the tile's P.V is issued.  Probabilities therefore stay <= 2^8 (f16 holds 65504), O / l at the end is exact in the same sense as
// with the true maximum, and after the first tile a rescale is rare.  K and V tiles travel in separate two-stage rings (K of tile
// t + 2 is requested as soon as the scores of tile t are done, V of tile t + 1 when P.V of tile t - 1 is).
// Warp roles: warp 0 = TMA, warp 1 = MMA issue (one elected thread), warps 2-17 = soft-max / epilogue: four threads per row
// (= TMEM lane), 32 key columns each; the four exchange their tile maxima through shared memory (one named barrier per tile).
// Arithmetic as the mma.sync kernel (prefill_kernels.cuh): f16 operands, f32 accumulation, P rounded to f16 before P.V.
#pragma once
#include "prefill_gemm.cuh"

namespace blk {

constexpr int AT_SM_WARPS = 16;                            // soft-max warps: four per TMEM lane quarter
constexpr int AT_SM_THREADS = 32 * AT_SM_WARPS;
constexpr int AT_THREADS = 64 + AT_SM_THREADS;             // warp 0 TMA, warp 1 MMA, warps 2-17 soft-max (four threads per row)
constexpr int AT_TILE_BYTES = 128 * 64 * 2;                 // one [128][64] f16 K block = 16 KB
constexpr int AT_MX_BYTES = 2 * 4 * 128 * 2;              // tile maxima of the four threads of a row, f16, double-buffered
constexpr int AT_USED_BYTES = (2 + 2 * 2 + 2 * 2 + 2 * 2) * AT_TILE_BYTES + 256 /*barriers*/ + AT_MX_BYTES;      // Q, K x2, V x2, P x2
constexpr int AT_SMEM_BYTES = AT_USED_BYTES;               // the dynamic shared memory starts 1024-aligned (checked in the kernel)

__device__ __forceinline__ void tc_st_32x32b_x32(uint32_t taddr, const uint32_t* r) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
        "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
        "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};"
        ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]),
          "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]),
          "r"(r[16]), "r"(r[17]), "r"(r[18]), "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]),
          "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31])
        : "memory");
}
__device__ __forceinline__ void tc_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ float ex2_approx(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__host__ __device__ constexpr uint32_t umma_idesc_f16(int M, int N) {      // f16 x f16 -> f32, both K-major
    return (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

struct AttnTcArgs {
    const int32_t* page_table; const int32_t* pos0;
    __nv_bfloat16* out;          // [T][n_head * 128]
    int T, n_head, n_pages;
    float scale;
};

// V rows [key][128] of one KV head -> V^T [128][ctx_pad] (keys contiguous), zero beyond n_keys (P is 0 there, V must be finite)
__global__ void __launch_bounds__(256) vt_transpose_kernel(const __half* __restrict__ v_pool, const int32_t* __restrict__ page_table,
                                                           int kv_dim, int n_keys, int ctx_pad, __half* __restrict__ vt) {
    __shared__ __half tile[64][128 + 2];
    pdl_launch_dependents(); pdl_wait();
    const int k0 = blockIdx.x * 64, hk = blockIdx.y, tid = threadIdx.x;
    for (int i = tid; i < 64 * 16; i += 256) {
        const int r = i >> 4, c = i & 15, key = k0 + r;
        uint4 v = make_uint4(0, 0, 0, 0);
        if (key < n_keys) v = *reinterpret_cast<const uint4*>(v_pool + ((size_t)page_table[key / KV_PAGE] * KV_PAGE + (key % KV_PAGE)) * kv_dim + (size_t)hk * 128 + c * 8);
        const __half* h = reinterpret_cast<const __half*>(&v);
#pragma unroll
        for (int j = 0; j < 8; j++) tile[r][c * 8 + j] = h[j];
    }
    __syncthreads();
    for (int i = tid; i < 128 * 8; i += 256) {          // 8 keys (16 B) per store
        const int d = i >> 3, kc = i & 7;
        __align__(16) __half o[8];
#pragma unroll
        for (int j = 0; j < 8; j++) o[j] = tile[kc * 8 + j][d];
        *reinterpret_cast<uint4*>(vt + ((size_t)hk * 128 + d) * ctx_pad + k0 + kc * 8) = *reinterpret_cast<const uint4*>(o);
    }
}

// GQ = query heads per KV head; GQS = head slots of the 128-row tile (power of two >= GQ; rows of the unused slots are dead)
template <int GQ, int GQS = GQ>
__global__ void __launch_bounds__(AT_THREADS, 1) prefill_attn_tc_kernel(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_k,
                                                                        const __grid_constant__ CUtensorMap tmap_vt, const AttnTcArgs a) {
    constexpr int BQ = 128 / GQS;
    extern __shared__ __align__(1024) unsigned char at_smem_raw[];
    unsigned char* smem = at_smem_raw;      // 1024-aligned (checked below); no integer cast, so that the accesses stay LDS / STS
    unsigned char* sQ = smem;                                   // 2 K blocks
    unsigned char* sK = sQ + 2 * AT_TILE_BYTES;                 // 2 stages x 2 K blocks
    unsigned char* sV = sK + 4 * AT_TILE_BYTES;                 // 2 stages x 2 key blocks
    unsigned char* sP = sV + 4 * AT_TILE_BYTES;                 // 2 buffers x 2 key blocks
    uint64_t* bars = reinterpret_cast<uint64_t*>(sP + 4 * AT_TILE_BYTES);
    uint64_t* q_full = bars;            // 1
    uint64_t* k_full = bars + 1;        // [2]
    uint64_t* k_empty = bars + 3;       // [2]  tcgen05.commit after Q.K^T
    uint64_t* v_full = bars + 5;        // [2]
    uint64_t* v_empty = bars + 7;       // [2]  tcgen05.commit after P.V
    uint64_t* s_full = bars + 9;        // [2]
    uint64_t* s_empty = bars + 11;      // [2]  all soft-max threads
    uint64_t* p_full = bars + 13;       // [2]  all soft-max threads
    uint64_t* p_empty = bars + 15;      // [2]  tcgen05.commit after P.V
    uint64_t* o_full = bars + 17;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 18);
    __half* s_mx = reinterpret_cast<__half*>(reinterpret_cast<unsigned char*>(bars) + 256);      // [2][4 parts][128 rows]
    if (threadIdx.x == 0 && (smem_u32(at_smem_raw) & 1023u)) __trap();      // the tiles need 1024-byte alignment (SWIZZLE_128B atoms)

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int qt = (int)gridDim.x - 1 - (int)blockIdx.x, hk = blockIdx.y;      // late (long) query tiles first
    const int pos0 = a.pos0[0];
    const int q0 = qt * BQ;
    const int kv_end = min(pos0 + a.T, pos0 + q0 + BQ);         // keys this tile can see
    const int n_tiles = (kv_end + 127) / 128;

    if (threadIdx.x == 0) {
        mbar_init(q_full, 1);
        for (int s = 0; s < 2; s++) {
            mbar_init(k_full + s, 1); mbar_init(k_empty + s, 1); mbar_init(v_full + s, 1); mbar_init(v_empty + s, 1);
            mbar_init(s_full + s, 1); mbar_init(s_empty + s, AT_SM_THREADS); mbar_init(p_full + s, AT_SM_THREADS); mbar_init(p_empty + s, 1);
        }
        mbar_init(o_full, 1);
        mbar_fence_init();
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;
    const uint32_t tS[2] = {tmem, tmem + 128}, tO = tmem + 256;
    pdl_launch_dependents(); pdl_wait();      // the prologue touched no global memory

    if (warp == 0) {
        // ===================== TMA: K0 K1 V0 K2 V1 K3 ... (the order in which the MMA warp frees the slots) =====================
        if (lane == 0) {
            mbar_expect_tx(q_full, 2 * GQ * BQ * 128);
#pragma unroll
            for (int kb = 0; kb < 2; kb++)
                for (int g = 0; g < GQ; g++)
                    tma_load_2d(sQ + kb * AT_TILE_BYTES + g * BQ * 128, &tmap_q, (hk * GQ + g) * 128 + kb * 64, q0, q_full);
            auto load_k = [&](int t) {
                const int s = t & 1;
                mbar_wait(k_empty + s, ((t >> 1) & 1) ^ 1);
                mbar_expect_tx(k_full + s, 2 * AT_TILE_BYTES);
                const int pg0 = min(2 * t, a.n_pages - 1), pg1 = min(2 * t + 1, a.n_pages - 1);     // a tile = two 64-token pages
                const int r0 = a.page_table[pg0] * KV_PAGE, r1 = a.page_table[pg1] * KV_PAGE;
#pragma unroll
                for (int kb = 0; kb < 2; kb++) {
                    unsigned char* dk = sK + (s * 2 + kb) * AT_TILE_BYTES;
                    tma_load_2d(dk, &tmap_k, hk * 128 + kb * 64, r0, k_full + s);
                    tma_load_2d(dk + 64 * 128, &tmap_k, hk * 128 + kb * 64, r1, k_full + s);
                }
            };
            auto load_v = [&](int t) {
                const int s = t & 1;
                mbar_wait(v_empty + s, ((t >> 1) & 1) ^ 1);
                mbar_expect_tx(v_full + s, 2 * AT_TILE_BYTES);
#pragma unroll
                for (int kb = 0; kb < 2; kb++)          // key block kb of the tile: [128 dims][64 keys]
                    tma_load_2d(sV + (s * 2 + kb) * AT_TILE_BYTES, &tmap_vt, t * 128 + kb * 64, hk * 128, v_full + s);
            };
            load_k(0);
            if (n_tiles > 1) load_k(1);
            for (int t = 0; t < n_tiles; t++) {
                load_v(t);
                if (t + 2 < n_tiles) load_k(t + 2);
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issue =====================
        if (lane == 0) {
            constexpr uint32_t idesc = umma_idesc_f16(128, 128);
            mbar_wait(q_full, 0);
            tc_fence_after();
            // the scores of tile t + 1 are issued BEFORE P.V of tile t, so its soft-max overlaps that product
            auto issue_qk = [&](int t) {        // S[t & 1] = Q . K[t & 1]^T over dh = 2 K blocks x 4 steps of 16
                const int s = t & 1;
                mbar_wait(k_full + s, (t >> 1) & 1);
                mbar_wait(s_empty + s, ((t >> 1) & 1) ^ 1);
                tc_fence_after();
#pragma unroll
                for (int kb = 0; kb < 2; kb++) {
                    const uint64_t dq = umma_desc_sw128(sQ + kb * AT_TILE_BYTES), dk = umma_desc_sw128(sK + (s * 2 + kb) * AT_TILE_BYTES);
#pragma unroll
                    for (int k = 0; k < 4; k++) tc_mma_f16(tS[s], dq + (uint64_t)((k * 32) >> 4), dk + (uint64_t)((k * 32) >> 4), idesc, (kb | k) ? 1u : 0u);
                }
                tc_commit(k_empty + s);
                tc_commit(s_full + s);
            };
            issue_qk(0);
            for (int t = 0; t < n_tiles; t++) {
                if (t + 1 < n_tiles) issue_qk(t + 1);
                const int s = t & 1;
                mbar_wait(p_full + s, (t >> 1) & 1);                        // P of this tile is in shared memory, O has been rescaled if it had to be
                mbar_wait(v_full + s, (t >> 1) & 1);
                tc_fence_after();
#pragma unroll
                for (int kb = 0; kb < 2; kb++) {
                    const uint64_t dp = umma_desc_sw128(sP + (s * 2 + kb) * AT_TILE_BYTES), dv = umma_desc_sw128(sV + (s * 2 + kb) * AT_TILE_BYTES);
#pragma unroll
                    for (int k = 0; k < 4; k++) tc_mma_f16(tO, dp + (uint64_t)((k * 32) >> 4), dv + (uint64_t)((k * 32) >> 4), idesc, (t | kb | k) ? 1u : 0u);
                }
                tc_commit(v_empty + s);
                tc_commit(p_empty + s);
            }
            tc_commit(o_full);
        }
    } else {
        // ===================== soft-max / epilogue: four threads per row (= TMEM lane), 32 key columns each =====================
        const int row = 32 * (warp & 3) + lane;             // warp w may touch TMEM lanes 32 (w % 4) .. +31
        const int part = (warp - 2) >> 2;                   // key columns [32 part, 32 part + 32) of every tile
        const int g = row / BQ, tok = q0 + (row % BQ);
        const bool row_ok = tok < a.T && g < GQ;
        const int last_key = pos0 + tok;                    // causal: keys <= last_key
        const float sl2 = a.scale * 1.4426950408889634f;
        const uint32_t lane_off = (uint32_t)(32 * (warp & 3)) << 16;
        float m_ref = -INFINITY, l = 0.0f;                  // reference maximum (exp2 domain) and the row sum relative to it
        unsigned char* prow0 = sP + (part >> 1) * AT_TILE_BYTES + (row >> 3) * 1024 + (row & 7) * 128;
        for (int t = 0; t < n_tiles; t++) {
            const int b = t & 1;
            mbar_wait(s_full + b, (t >> 1) & 1);
            tc_fence_after();
            uint32_t v[32];
            tc_ld_32x32b_x32(tS[b] + (uint32_t)(part * 32) + lane_off, v);
            tc_wait_ld();
            tc_fence_before();
            mbar_arrive(s_empty + b);                       // the scores are in registers: S[b] is free for tile t + 2
            const int key0 = t * 128 + part * 32;
            const bool full = row_ok && key0 + 31 <= last_key;      // no key of this chunk is masked (every tile but the diagonal one)
            // -- this thread's maximum, the row's through shared memory --
            float lm;
            if (full) {
                float m0 = -INFINITY, m1 = -INFINITY, m2 = -INFINITY, m3 = -INFINITY;
#pragma unroll
                for (int j = 0; j < 32; j += 4) {
                    m0 = fmaxf(m0, __uint_as_float(v[j])); m1 = fmaxf(m1, __uint_as_float(v[j + 1]));
                    m2 = fmaxf(m2, __uint_as_float(v[j + 2])); m3 = fmaxf(m3, __uint_as_float(v[j + 3]));
                }
                lm = fmaxf(fmaxf(m0, m1), fmaxf(m2, m3));
            } else {
                lm = -INFINITY;
#pragma unroll
                for (int j = 0; j < 32; j++) if (row_ok && key0 + j <= last_key) lm = fmaxf(lm, __uint_as_float(v[j]));
            }
            __half* mx = s_mx + b * 512;
            mx[part * 128 + row] = __float2half_ru(lm * sl2);       // rounded UP: the reference may exceed the true maximum, never fall short of it
            asm volatile("bar.sync 1, %0;" ::"n"(AT_SM_THREADS) : "memory");
            const float mt = fmaxf(fmaxf(__half2float(mx[row]), __half2float(mx[128 + row])), fmaxf(__half2float(mx[256 + row]), __half2float(mx[384 + row])));
            // -- raise the reference when the tile exceeds it by more than 2^8; O and l follow (warp-uniform: tcgen05.ld / st are collective) --
            const bool raise = mt > m_ref + 8.0f;           // first tile: m_ref = -inf
            if (__any_sync(0xffffffffu, raise)) {
                const float f = raise ? ex2_approx(m_ref - mt) : 1.0f;      // exp2(-inf) = 0 on the first tile
                if (raise) { m_ref = mt; l *= f; }
                if (t > 0) {
                    mbar_wait(p_empty + ((t - 1) & 1), ((t - 1) >> 1) & 1);     // P.V of tile t - 1 has completed: O is quiescent
                    tc_fence_after();
                    uint32_t o[32];
                    tc_ld_32x32b_x32(tO + (uint32_t)(part * 32) + lane_off, o);
                    tc_wait_ld();
#pragma unroll
                    for (int j = 0; j < 32; j++) o[j] = __float_as_uint(__uint_as_float(o[j]) * f);
                    tc_st_32x32b_x32(tO + (uint32_t)(part * 32) + lane_off, o);
                    tc_wait_st();
                    tc_fence_before();
                }
            }
            // -- P = exp2(s - m_ref) -> f16, row sum alongside --
            uint32_t pk[16];
            if (full) {
                float l0 = 0.0f, l1 = 0.0f, l2 = 0.0f, l3 = 0.0f;
#pragma unroll
                for (int j = 0; j < 32; j += 4) {
                    const float p0 = ex2_approx(fmaf(__uint_as_float(v[j]), sl2, -m_ref)), p1 = ex2_approx(fmaf(__uint_as_float(v[j + 1]), sl2, -m_ref));
                    const float p2 = ex2_approx(fmaf(__uint_as_float(v[j + 2]), sl2, -m_ref)), p3 = ex2_approx(fmaf(__uint_as_float(v[j + 3]), sl2, -m_ref));
                    l0 += p0; l1 += p1; l2 += p2; l3 += p3;
                    __half2 h0 = __floats2half2_rn(p0, p1), h1 = __floats2half2_rn(p2, p3);
                    pk[j >> 1] = *reinterpret_cast<uint32_t*>(&h0);
                    pk[(j >> 1) + 1] = *reinterpret_cast<uint32_t*>(&h1);
                }
                l += (l0 + l1) + (l2 + l3);
            } else {
#pragma unroll
                for (int j = 0; j < 32; j += 2) {
                    const float p0 = (row_ok && key0 + j <= last_key) ? ex2_approx(fmaf(__uint_as_float(v[j]), sl2, -m_ref)) : 0.0f;
                    const float p1 = (row_ok && key0 + j + 1 <= last_key) ? ex2_approx(fmaf(__uint_as_float(v[j + 1]), sl2, -m_ref)) : 0.0f;
                    l += p0 + p1;
                    __half2 h = __floats2half2_rn(p0, p1);
                    pk[j >> 1] = *reinterpret_cast<uint32_t*>(&h);
                }
            }
            mbar_wait(p_empty + b, ((t >> 1) & 1) ^ 1);     // the P.V of tile t - 2 has read this P buffer
            unsigned char* prow = prow0 + b * 2 * AT_TILE_BYTES;
#pragma unroll
            for (int q = 0; q < 4; q++) {                   // this thread's 32 keys = four 16-byte chunks of its row in key block part / 2
                const int ch = (part & 1) * 4 + q;
                *reinterpret_cast<uint4*>(prow + ((ch ^ (row & 7)) << 4)) = make_uint4(pk[4 * q], pk[4 * q + 1], pk[4 * q + 2], pk[4 * q + 3]);
            }
            fence_proxy_async();
            mbar_arrive(p_full + b);
        }
        // ---- epilogue: O / l -> bf16, this thread's 32 of the 128 head dims ----
        mbar_wait(o_full, 0);                               // every MMA has completed: the P buffers are free for the row sums
        tc_fence_after();
        float* s_x = reinterpret_cast<float*>(sP);          // [4 parts][128 rows]
        s_x[part * 128 + row] = l;
        asm volatile("bar.sync 1, %0;" ::"n"(AT_SM_THREADS) : "memory");
        l = (s_x[row] + s_x[128 + row]) + (s_x[256 + row] + s_x[384 + row]);
        const float inv = l > 0.0f ? 1.0f / l : 0.0f;
        __nv_bfloat16* dst = a.out + (size_t)tok * ((size_t)a.n_head * 128) + (size_t)(hk * GQ + g) * 128 + part * 32;
        {
            uint32_t v[32];
            tc_ld_32x32b_x32(tO + (uint32_t)(part * 32) + lane_off, v);
            tc_wait_ld();
            if (row_ok) {
#pragma unroll
                for (int j = 0; j < 32; j += 8) {
                    uint4 o;
                    o.x = pack_bf16x2(__uint_as_float(v[j]) * inv, __uint_as_float(v[j + 1]) * inv);
                    o.y = pack_bf16x2(__uint_as_float(v[j + 2]) * inv, __uint_as_float(v[j + 3]) * inv);
                    o.z = pack_bf16x2(__uint_as_float(v[j + 4]) * inv, __uint_as_float(v[j + 5]) * inv);
                    o.w = pack_bf16x2(__uint_as_float(v[j + 6]) * inv, __uint_as_float(v[j + 7]) * inv);
                    *reinterpret_cast<uint4*>(dst + j) = o;
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512) : "memory");
}

} // namespace blk
