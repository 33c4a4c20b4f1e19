// qweights.cuh -- device-resident quantised weight matrices.
//
// A GGUF tensor arrives as an array of ggml blocks (block_q4_K = 144 B, block_q5_K = 176 B, block_q6_K = 210 B,
// block_q8_0 = 34 B; upstream ggml/src/ggml-common.h).  None of those sizes is a multiple of 16 B, so streaming them
// with 128-bit loads is impossible as-is.  At load time each matrix is re-tiled ON THE DEVICE into planes that keep
// exactly the same bits (same bytes per weight) but are 16 B aligned per row and per super-block:
//
//   Q4_K : qs  [N][K/2]      nibbles, ggml order (32 B per 64-element group)
//          hdr [N][K/256][16] {f16 d, f16 dmin, u8 scales[12]}
//   Q5_K : qs  [N][K/2], qh [N][K/8] (32 B / super-block), hdr as Q4_K
//   Q6_K : ql  [N][K/2], qh [N][K/4] (64 B / super-block), sc [N][K/16] int8, d [N][K/256] f16
//   Q8_0 : qs  [N][K] int8, d [N][K/32] f16
//   F32 / F16 : as is
//
// A "unit" is the slice of one row a single lane handles per step: 64 weights for the K-quants (32 B of nibbles),
// 32 weights for Q8_0 -- always two 128-bit loads of quantised data.
#pragma once
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace blk {

enum : int { QT_F32 = 0, QT_F16 = 1, QT_Q8_0 = 8, QT_Q4_K = 12, QT_Q5_K = 13, QT_Q6_K = 14 };

// activation formats the decode kernels consume (what ggml calls vec_dot_type of the weight type)
enum : int { ACT_F32 = 0, ACT_Q8_K = 1, ACT_Q8_0 = 2 };

__host__ __device__ inline int act_format_for(int qt) {
    return (qt == QT_Q4_K || qt == QT_Q5_K || qt == QT_Q6_K) ? ACT_Q8_K : (qt == QT_Q8_0 ? ACT_Q8_0 : ACT_F32);
}

struct QMat {
    int type = -1;
    int N = 0;            // rows (output features)
    int K = 0;            // row length (input features)
    const uint8_t* p0 = nullptr;   // qs / ql / raw
    const uint8_t* p1 = nullptr;   // hdr (Q4_K,Q5_K) | qh (Q6_K) | d (Q8_0)
    const uint8_t* p2 = nullptr;   // qh (Q5_K) | sc (Q6_K)
    const uint8_t* p3 = nullptr;   // d (Q6_K)
    size_t bytes = 0;              // total device bytes (== ggml bytes)
};

__host__ __device__ inline int qmat_unit_elems(int type) { return type == QT_Q8_0 ? 32 : (type == QT_F32 || type == QT_F16 ? 8 : 64); }

// ---------------------------------------------------------------------------------------------------------------
// K-quant 6-bit scale / min pairs of 64-element group j (sub-blocks 2j, 2j+1) from the 12 packed bytes
// (hdr.y = bytes 0-3, hdr.z = 4-7, hdr.w = 8-11).  Result: two bytes packed in the low 16 bits.
// ---------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void k4_scale_min_pair(const uint4& hdr, int j, uint32_t& sc2, uint32_t& mn2) {
    const int sh = (j & 1) * 16;
    const uint32_t a = hdr.y >> sh, b = hdr.z >> sh, c = hdr.w >> sh;
    if (j < 2) {
        sc2 = a & 0x3F3Fu;
        mn2 = b & 0x3F3Fu;
    } else {
        sc2 = (c & 0x0F0Fu) | ((a >> 2) & 0x3030u);
        mn2 = ((c >> 4) & 0x0F0Fu) | ((b >> 2) & 0x3030u);
    }
}

__device__ __forceinline__ float2 hdr_d_dmin(const uint4& hdr) {
    const __half2 h = *reinterpret_cast<const __half2*>(&hdr.x);
    return __half22float2(h);
}

// ---------------------------------------------------------------------------------------------------------------
// element-wise dequantisation on the split layout (embedding gather, prefill GEMM operand producer, tests).
// Same float expression order as ggml-quants.c dequantize_row_* so values are bit-identical to the reference's.
// ---------------------------------------------------------------------------------------------------------------
// 64 consecutive weights of row `row`, unit u (elements 64u .. 64u+63), K-quants; 32 for Q8_0 unit.
__device__ __forceinline__ void dequant_unit_q4k(const QMat& W, int64_t row, int u, float* out /*64*/) {
    const int nsb = W.K >> 8;
    const uint4 hdr = *reinterpret_cast<const uint4*>(W.p1 + ((size_t)row * nsb + (u >> 2)) * 16);
    const uint8_t* q = W.p0 + (size_t)row * (W.K >> 1) + (size_t)u * 32;
    uint32_t sc2, mn2; k4_scale_min_pair(hdr, u & 3, sc2, mn2);
    const float2 dm = hdr_d_dmin(hdr);
    const float d1 = dm.x * (float)(sc2 & 0xFF), m1 = dm.y * (float)(mn2 & 0xFF);
    const float d2 = dm.x * (float)(sc2 >> 8), m2 = dm.y * (float)(mn2 >> 8);
#pragma unroll
    for (int l = 0; l < 32; l++) {
        const uint8_t b = q[l];
        out[l] = __fsub_rn(__fmul_rn(d1, (float)(b & 0xF)), m1);          // no FMA contraction: bit-identical to ggml
        out[32 + l] = __fsub_rn(__fmul_rn(d2, (float)(b >> 4)), m2);
    }
}

__device__ __forceinline__ void dequant_unit_q5k(const QMat& W, int64_t row, int u, float* out) {
    const int nsb = W.K >> 8;
    const int s = u >> 2, j = u & 3;
    const uint4 hdr = *reinterpret_cast<const uint4*>(W.p1 + ((size_t)row * nsb + s) * 16);
    const uint8_t* q = W.p0 + (size_t)row * (W.K >> 1) + (size_t)u * 32;
    const uint8_t* qh = W.p2 + ((size_t)row * nsb + s) * 32;
    uint32_t sc2, mn2; k4_scale_min_pair(hdr, j, sc2, mn2);
    const float2 dm = hdr_d_dmin(hdr);
    const float d1 = dm.x * (float)(sc2 & 0xFF), m1 = dm.y * (float)(mn2 & 0xFF);
    const float d2 = dm.x * (float)(sc2 >> 8), m2 = dm.y * (float)(mn2 >> 8);
#pragma unroll
    for (int l = 0; l < 32; l++) {
        const uint8_t b = q[l], h = qh[l];
        out[l] = __fsub_rn(__fmul_rn(d1, (float)((b & 0xF) + (((h >> (2 * j)) & 1) ? 16 : 0))), m1);
        out[32 + l] = __fsub_rn(__fmul_rn(d2, (float)((b >> 4) + (((h >> (2 * j + 1)) & 1) ? 16 : 0))), m2);
    }
}

// Q6_K unit u: super-block s = u/4, half hh = (u%4)/2, t = u%2 -> l in [16t, 16t+16); produces the 4 x 16 weights at
// elements 256s + 128hh + 32k + l  (k = 0..3) -- out[k*16 + (l-16t)]
__device__ __forceinline__ void dequant_unit_q6k(const QMat& W, int64_t row, int u, float* out) {
    const int nsb = W.K >> 8;
    const int s = u >> 2, hh = (u >> 1) & 1, t = u & 1;
    const uint8_t* ql = W.p0 + (size_t)row * (W.K >> 1) + (size_t)s * 128 + hh * 64 + t * 16;
    const uint8_t* qh = W.p1 + (size_t)row * (W.K >> 2) + (size_t)s * 64 + hh * 32 + t * 16;
    const int8_t* sc = reinterpret_cast<const int8_t*>(W.p2) + (size_t)row * (W.K >> 4) + s * 16 + hh * 8 + t;
    const float d = __half2float(reinterpret_cast<const __half*>(W.p3)[(size_t)row * nsb + s]);
#pragma unroll
    for (int l = 0; l < 16; l++) {
        const uint8_t a = ql[l], b = ql[l + 32], h = qh[l];
        const int q1 = (int)((a & 0xF) | (((h >> 0) & 3) << 4)) - 32;
        const int q2 = (int)((b & 0xF) | (((h >> 2) & 3) << 4)) - 32;
        const int q3 = (int)((a >> 4) | (((h >> 4) & 3) << 4)) - 32;
        const int q4 = (int)((b >> 4) | (((h >> 6) & 3) << 4)) - 32;
        out[l] = d * (float)sc[0] * (float)q1;
        out[16 + l] = d * (float)sc[2] * (float)q2;
        out[32 + l] = d * (float)sc[4] * (float)q3;
        out[48 + l] = d * (float)sc[6] * (float)q4;
    }
}
// element index of out[i] of a Q6_K unit within the row
__device__ __forceinline__ int q6k_unit_elem(int u, int i) {
    const int s = u >> 2, hh = (u >> 1) & 1, t = u & 1;
    return 256 * s + 128 * hh + 32 * (i >> 4) + 16 * t + (i & 15);
}

__device__ __forceinline__ void dequant_unit_q80(const QMat& W, int64_t row, int u, float* out /*32*/) {
    const int8_t* q = reinterpret_cast<const int8_t*>(W.p0) + (size_t)row * W.K + (size_t)u * 32;
    const float d = __half2float(reinterpret_cast<const __half*>(W.p1)[(size_t)row * (W.K >> 5) + u]);
#pragma unroll
    for (int l = 0; l < 32; l++) out[l] = (float)q[l] * d;
}

// dequantise one full row into f32 (any type) -- one CTA per row
__device__ inline void dequant_row_cta(const QMat& W, int64_t row, float* out) {
    const int tid = threadIdx.x, nt = blockDim.x;
    if (W.type == QT_F32) {
        const float* src = reinterpret_cast<const float*>(W.p0) + (size_t)row * W.K;
        for (int i = tid; i < W.K; i += nt) out[i] = src[i];
    } else if (W.type == QT_F16) {
        const __half* src = reinterpret_cast<const __half*>(W.p0) + (size_t)row * W.K;
        for (int i = tid; i < W.K; i += nt) out[i] = __half2float(src[i]);
    } else if (W.type == QT_Q8_0) {
        for (int u = tid; u < (W.K >> 5); u += nt) { float v[32]; dequant_unit_q80(W, row, u, v); for (int l = 0; l < 32; l++) out[u * 32 + l] = v[l]; }
    } else if (W.type == QT_Q4_K) {
        for (int u = tid; u < (W.K >> 6); u += nt) { float v[64]; dequant_unit_q4k(W, row, u, v); for (int l = 0; l < 64; l++) out[u * 64 + l] = v[l]; }
    } else if (W.type == QT_Q5_K) {
        for (int u = tid; u < (W.K >> 6); u += nt) { float v[64]; dequant_unit_q5k(W, row, u, v); for (int l = 0; l < 64; l++) out[u * 64 + l] = v[l]; }
    } else if (W.type == QT_Q6_K) {
        for (int u = tid; u < (W.K >> 6); u += nt) { float v[64]; dequant_unit_q6k(W, row, u, v); for (int l = 0; l < 64; l++) out[q6k_unit_elem(u, l)] = v[l]; }
    }
}

} // namespace blk
