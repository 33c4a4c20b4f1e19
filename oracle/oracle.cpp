// oracle.cpp -- CPU oracle for the blama hot path.  TEST INFRASTRUCTURE ONLY (see oracle.h).
//
// Restates, in plain C++, what blama computes through llama.cpp b5187 on its CPU backend
// (reference inference/code/llama/Model.cpp:27-31 selects it with gpu=false):
//   * GGUF v3 parsing                      [upstream ggml/src/gguf.cpp]
//   * block formats + dequantisation       [upstream ggml/src/ggml-common.h, ggml-quants.c]
//   * Q8_K / Q8_0 activation quantisation  [upstream ggml-quants.c quantize_row_q8_K_ref / q8_0_ref]
//   * integer dot products                 [upstream ggml-cpu-quants.c ggml_vec_dot_*_q8_K, generic branch]
//   * rms_norm / rope / soft_max / silu    [upstream ggml-cpu ops]
//   * llama / qwen2 layer graphs           [upstream src/llama-model.cpp llm_build_llama / llm_build_qwen2]
//   * sampler chain                        [upstream src/llama-sampling.cpp], as configured by
//                                          reference inference/code/llama/Sampler.cpp:15-97
//   * blama's Session control flow         reference inference/code/llama/Session.cpp:65-107,169-282
//   * blama's LogitComparer                reference inference/code/llama/LogitComparer.cpp:8-128
// llama.cpp itself is NOT present in this container (no network); its algorithms are restated from
// the published source at that tag ("upstream-recall" in SURVEY.md).  PARITY of the transformer forward is
// therefore UNPINNED by reference fixtures; LogitComparer and dequantisation ARE pinned (tests/), and the float path of the
// forward is pinned to an independent implementation reading the same GGUF (Hugging Face transformers, tests/test_oracle_hf_pin.py).
#include "oracle.h"

#include <algorithm>
#include <cassert>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <functional>
#include <map>
#include <memory>
#include <random>
#include <stdexcept>
#include <string>
#include <unordered_map>
#include <vector>

#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>
#include <atomic>
#include <condition_variable>
#include <mutex>
#include <thread>

namespace {

// ------------------------------------------------------------------------------------------------
// persistent thread pool (ggml-cpu uses its own threadpool; llama.cpp's default is 4 threads)
// ------------------------------------------------------------------------------------------------
class ThreadPool {
public:
    explicit ThreadPool(int n) : n_(std::max(1, n)) {
        for (int i = 1; i < n_; i++) workers_.emplace_back([this, i] { loop(i); });
    }
    ~ThreadPool() {
        { std::lock_guard<std::mutex> lk(mu_); stop_ = true; gen_++; }
        cv_.notify_all();
        for (auto& t : workers_) t.join();
    }
    int size() const { return n_; }
    // fn(begin, end) over [0, total) split into contiguous chunks, dynamic chunk claiming
    void run(int64_t total, int64_t grain, const std::function<void(int64_t, int64_t)>& fn) {
        if (total <= 0) return;
        if (n_ == 1 || total <= grain) { fn(0, total); return; }
        {
            std::lock_guard<std::mutex> lk(mu_);
            fn_ = &fn; total_ = total; grain_ = std::max<int64_t>(1, grain); next_.store(0); pending_ = n_ - 1; gen_++;
        }
        cv_.notify_all();
        work();
        std::unique_lock<std::mutex> lk(mu_);
        done_cv_.wait(lk, [this] { return pending_ == 0; });
        fn_ = nullptr;
    }
private:
    void work() {
        for (;;) {
            int64_t b = next_.fetch_add(grain_);
            if (b >= total_) break;
            (*fn_)(b, std::min(total_, b + grain_));
        }
    }
    void loop(int) {
        uint64_t seen = 0;
        for (;;) {
            {
                std::unique_lock<std::mutex> lk(mu_);
                cv_.wait(lk, [&] { return gen_ != seen; });
                seen = gen_;
                if (stop_) return;
            }
            work();
            { std::lock_guard<std::mutex> lk(mu_); if (--pending_ == 0) done_cv_.notify_one(); }
        }
    }
    int n_; std::vector<std::thread> workers_;
    std::mutex mu_; std::condition_variable cv_, done_cv_;
    const std::function<void(int64_t, int64_t)>* fn_ = nullptr;
    int64_t total_ = 0, grain_ = 1; std::atomic<int64_t> next_{0}; int pending_ = 0; uint64_t gen_ = 0; bool stop_ = false;
};

// ------------------------------------------------------------------------------------------------
// scalar helpers
// ------------------------------------------------------------------------------------------------
inline float h2f(uint16_t h) { _Float16 v; memcpy(&v, &h, 2); return (float)v; }
inline uint16_t f2h(float f) { _Float16 v = (_Float16)f; uint16_t h; memcpy(&h, &v, 2); return h; }
inline float bf16_round(float f) {
    uint32_t u; memcpy(&u, &f, 4);
    if ((u & 0x7fffffffu) > 0x7f800000u) return f;       // NaN
    u += 0x7fffu + ((u >> 16) & 1u);
    u &= 0xffff0000u;
    memcpy(&f, &u, 4); return f;
}
// ggml-quants.c nearest_int: round-half-even through the 1.5*2^23 magic constant
inline int nearest_int(float fval) {
    float val = fval + 12582912.f;
    int i; memcpy(&i, &val, sizeof(int));
    return (i & 0x007fffff) - 0x00400000;
}

enum { T_F32 = 0, T_F16 = 1, T_Q8_0 = 8, T_Q4_K = 12, T_Q5_K = 13, T_Q6_K = 14 };
constexpr int QK_K = 256;

struct TypeInfo { int blck; int bytes; };
inline bool type_info(int t, TypeInfo& ti) {
    switch (t) {
        case T_F32: ti = {1, 4}; return true;
        case T_F16: ti = {1, 2}; return true;
        case T_Q8_0: ti = {32, 34}; return true;
        case T_Q4_K: ti = {256, 144}; return true;
        case T_Q5_K: ti = {256, 176}; return true;
        case T_Q6_K: ti = {256, 210}; return true;
    }
    return false;
}
inline size_t row_bytes(int type, int64_t k) { TypeInfo ti{}; type_info(type, ti); return (size_t)(k / ti.blck) * ti.bytes; }

// K-quant 6-bit scale/min unpack (ggml-quants.c get_scale_min_k4)
inline void scale_min_k4(int j, const uint8_t* q, uint8_t& sc, uint8_t& mn) {
    if (j < 4) { sc = q[j] & 63; mn = q[j + 4] & 63; }
    else {
        sc = (q[j + 4] & 0xF) | ((q[j - 4] >> 6) << 4);
        mn = (q[j + 4] >> 4) | ((q[j] >> 6) << 4);
    }
}

// ------------------------------------------------------------------------------------------------
// dequantisation (ggml-quants.c dequantize_row_*)
// ------------------------------------------------------------------------------------------------
void dequant_row(int type, const uint8_t* src, int64_t n, float* y) {
    switch (type) {
    case T_F32: memcpy(y, src, n * 4); return;
    case T_F16: { const uint16_t* h = (const uint16_t*)src; for (int64_t i = 0; i < n; i++) y[i] = h2f(h[i]); return; }
    case T_Q8_0:
        for (int64_t b = 0; b < n / 32; b++) {
            const uint8_t* blk = src + b * 34;
            uint16_t dh; memcpy(&dh, blk, 2);
            const float d = h2f(dh);
            const int8_t* qs = (const int8_t*)(blk + 2);
            for (int j = 0; j < 32; j++) y[b * 32 + j] = qs[j] * d;
        }
        return;
    case T_Q4_K:
        for (int64_t b = 0; b < n / QK_K; b++) {
            const uint8_t* blk = src + b * 144;
            uint16_t dh, mh; memcpy(&dh, blk, 2); memcpy(&mh, blk + 2, 2);
            const float d = h2f(dh), dmin = h2f(mh);
            const uint8_t* scales = blk + 4; const uint8_t* q = blk + 16;
            float* o = y + b * QK_K;
            int is = 0;
            for (int j = 0; j < QK_K; j += 64) {
                uint8_t sc, m;
                scale_min_k4(is + 0, scales, sc, m); const float d1 = d * sc, m1 = dmin * m;
                scale_min_k4(is + 1, scales, sc, m); const float d2 = d * sc, m2 = dmin * m;
                for (int l = 0; l < 32; l++) *o++ = d1 * (q[l] & 0xF) - m1;
                for (int l = 0; l < 32; l++) *o++ = d2 * (q[l] >> 4) - m2;
                q += 32; is += 2;
            }
        }
        return;
    case T_Q5_K:
        for (int64_t b = 0; b < n / QK_K; b++) {
            const uint8_t* blk = src + b * 176;
            uint16_t dh, mh; memcpy(&dh, blk, 2); memcpy(&mh, blk + 2, 2);
            const float d = h2f(dh), dmin = h2f(mh);
            const uint8_t* scales = blk + 4; const uint8_t* qh = blk + 16; const uint8_t* ql = blk + 48;
            float* o = y + b * QK_K;
            int is = 0; uint8_t u1 = 1, u2 = 2;
            for (int j = 0; j < QK_K; j += 64) {
                uint8_t sc, m;
                scale_min_k4(is + 0, scales, sc, m); const float d1 = d * sc, m1 = dmin * m;
                scale_min_k4(is + 1, scales, sc, m); const float d2 = d * sc, m2 = dmin * m;
                for (int l = 0; l < 32; l++) *o++ = d1 * ((ql[l] & 0xF) + (qh[l] & u1 ? 16 : 0)) - m1;
                for (int l = 0; l < 32; l++) *o++ = d2 * ((ql[l] >> 4) + (qh[l] & u2 ? 16 : 0)) - m2;
                ql += 32; is += 2; u1 <<= 2; u2 <<= 2;
            }
        }
        return;
    case T_Q6_K:
        for (int64_t b = 0; b < n / QK_K; b++) {
            const uint8_t* blk = src + b * 210;
            const uint8_t* ql = blk; const uint8_t* qh = blk + 128; const int8_t* sc = (const int8_t*)(blk + 192);
            uint16_t dh; memcpy(&dh, blk + 208, 2);
            const float d = h2f(dh);
            float* o = y + b * QK_K;
            for (int n2 = 0; n2 < QK_K; n2 += 128) {
                for (int l = 0; l < 32; l++) {
                    const int is = l / 16;
                    const int8_t q1 = (int8_t)((ql[l + 0] & 0xF) | (((qh[l] >> 0) & 3) << 4)) - 32;
                    const int8_t q2 = (int8_t)((ql[l + 32] & 0xF) | (((qh[l] >> 2) & 3) << 4)) - 32;
                    const int8_t q3 = (int8_t)((ql[l + 0] >> 4) | (((qh[l] >> 4) & 3) << 4)) - 32;
                    const int8_t q4 = (int8_t)((ql[l + 32] >> 4) | (((qh[l] >> 6) & 3) << 4)) - 32;
                    o[l + 0] = d * sc[is + 0] * q1;
                    o[l + 32] = d * sc[is + 2] * q2;
                    o[l + 64] = d * sc[is + 4] * q3;
                    o[l + 96] = d * sc[is + 6] * q4;
                }
                o += 128; ql += 64; qh += 32; sc += 8;
            }
        }
        return;
    }
}

// ------------------------------------------------------------------------------------------------
// activation quantisation (ggml-quants.c quantize_row_q8_K_ref / quantize_row_q8_0_ref)
// ------------------------------------------------------------------------------------------------
struct Q8K { std::vector<int8_t> qs; std::vector<float> d; std::vector<int16_t> bsums; };
struct Q80 { std::vector<int8_t> qs; std::vector<float> d; };   // d already rounded through fp16

void quantize_q8_K(const float* x, int64_t k, int8_t* qs, float* dout, int16_t* bsums) {
    const int64_t nb = k / QK_K;
    for (int64_t i = 0; i < nb; i++) {
        float max = 0, amax = 0;
        for (int j = 0; j < QK_K; j++) { float ax = fabsf(x[j]); if (ax > amax) { amax = ax; max = x[j]; } }
        if (!amax) {
            dout[i] = 0; memset(qs, 0, QK_K); memset(bsums, 0, 16 * sizeof(int16_t));
            x += QK_K; qs += QK_K; bsums += 16; continue;
        }
        const float iscale = -127.f / max;
        for (int j = 0; j < QK_K; j++) { int v = nearest_int(iscale * x[j]); qs[j] = (int8_t)std::min(127, v); }
        for (int j = 0; j < 16; j++) { int s = 0; for (int ii = 0; ii < 16; ii++) s += qs[j * 16 + ii]; bsums[j] = (int16_t)s; }
        dout[i] = 1 / iscale;
        x += QK_K; qs += QK_K; bsums += 16;
    }
}

void quantize_q8_0(const float* x, int64_t k, int8_t* qs, float* dout) {
    const int64_t nb = k / 32;
    for (int64_t i = 0; i < nb; i++) {
        float amax = 0;
        for (int j = 0; j < 32; j++) amax = std::max(amax, fabsf(x[i * 32 + j]));
        const float d = amax / 127;
        const float id = d ? 1.0f / d : 0.0f;
        dout[i] = h2f(f2h(d));
        for (int j = 0; j < 32; j++) qs[i * 32 + j] = (int8_t)roundf(x[i * 32 + j] * id);
    }
}

// ORC_MODE_GGML_ALT: identical integer arithmetic, but the eight fp32 lane sums (K-quants) / the block products (Q8_0) are
// added in the opposite order.
// Exists only to measure how far two faithful implementations of the reference arithmetic drift apart
// (tests/test_oracle.py::test_summation_order_noise_floor).
static bool g_alt_order = false;
inline float lane_sum8(const float* sums, float sumf) {
    if (g_alt_order) { for (int l = 7; l >= 0; l--) sumf += sums[l]; }
    else { for (int l = 0; l < 8; l++) sumf += sums[l]; }
    return sumf;
}

// ------------------------------------------------------------------------------------------------
// integer dot products (ggml-cpu-quants.c, generic branch: 8 fp32 partial lanes per row)
// ------------------------------------------------------------------------------------------------
float vec_dot_q4_K_q8_K(int64_t k, const uint8_t* w, const int8_t* aq, const float* ad, const int16_t* absum) {
    const int64_t nb = k / QK_K;
    float sums[8] = {0}; float sumf = 0;
    int8_t a[QK_K];
    for (int64_t i = 0; i < nb; i++) {
        const uint8_t* blk = w + i * 144;
        uint16_t dh, mh; memcpy(&dh, blk, 2); memcpy(&mh, blk + 2, 2);
        const uint8_t* scales = blk + 4; const uint8_t* q4 = blk + 16; const int8_t* q8 = aq + i * QK_K;
        for (int j = 0; j < 4; j++) for (int l = 0; l < 32; l++) { a[j * 64 + l] = q4[j * 32 + l] & 0xF; a[j * 64 + 32 + l] = q4[j * 32 + l] >> 4; }
        uint8_t sc[8], mn[8];
        for (int j = 0; j < 8; j++) scale_min_k4(j, scales, sc[j], mn[j]);
        int sumi = 0;
        for (int j = 0; j < 16; j++) sumi += absum[i * 16 + j] * mn[j / 2];
        int32_t aux32[8] = {0};
        for (int j = 0; j < 8; j++) {
            const int32_t s = sc[j];
            for (int r = 0; r < 4; r++) for (int l = 0; l < 8; l++) aux32[l] += s * (int32_t)(q8[j * 32 + r * 8 + l] * a[j * 32 + r * 8 + l]);
        }
        const float d = h2f(dh) * ad[i];
        for (int l = 0; l < 8; l++) sums[l] += d * aux32[l];
        const float dmin = h2f(mh) * ad[i];
        sumf -= dmin * sumi;
    }
    return lane_sum8(sums, sumf);
}

float vec_dot_q5_K_q8_K(int64_t k, const uint8_t* w, const int8_t* aq, const float* ad, const int16_t* absum) {
    const int64_t nb = k / QK_K;
    float sums[8] = {0}; float sumf = 0;
    int8_t a[QK_K];
    for (int64_t i = 0; i < nb; i++) {
        const uint8_t* blk = w + i * 176;
        uint16_t dh, mh; memcpy(&dh, blk, 2); memcpy(&mh, blk + 2, 2);
        const uint8_t* scales = blk + 4; const uint8_t* hm = blk + 16; const uint8_t* q4 = blk + 48; const int8_t* q8 = aq + i * QK_K;
        uint8_t m = 1;
        for (int j = 0; j < 4; j++) {
            for (int l = 0; l < 32; l++) a[j * 64 + l] = (int8_t)((q4[j * 32 + l] & 0xF) + (hm[l] & m ? 16 : 0));
            m <<= 1;
            for (int l = 0; l < 32; l++) a[j * 64 + 32 + l] = (int8_t)((q4[j * 32 + l] >> 4) + (hm[l] & m ? 16 : 0));
            m <<= 1;
        }
        uint8_t sc[8], mn[8];
        for (int j = 0; j < 8; j++) scale_min_k4(j, scales, sc[j], mn[j]);
        int sumi = 0;
        for (int j = 0; j < 16; j++) sumi += absum[i * 16 + j] * mn[j / 2];
        int32_t aux32[8] = {0};
        for (int j = 0; j < 8; j++) {
            const int32_t s = sc[j];
            for (int r = 0; r < 4; r++) for (int l = 0; l < 8; l++) aux32[l] += s * (int32_t)(q8[j * 32 + r * 8 + l] * a[j * 32 + r * 8 + l]);
        }
        const float d = h2f(dh) * ad[i];
        for (int l = 0; l < 8; l++) sums[l] += d * aux32[l];
        const float dmin = h2f(mh) * ad[i];
        sumf -= dmin * sumi;
    }
    return lane_sum8(sums, sumf);
}

float vec_dot_q6_K_q8_K(int64_t k, const uint8_t* w, const int8_t* aq, const float* ad) {
    const int64_t nb = k / QK_K;
    float sums[8] = {0};
    int8_t a[QK_K];
    for (int64_t i = 0; i < nb; i++) {
        const uint8_t* blk = w + i * 210;
        const uint8_t* q4 = blk; const uint8_t* qh = blk + 128; const int8_t* sc = (const int8_t*)(blk + 192);
        uint16_t dh; memcpy(&dh, blk + 208, 2);
        const int8_t* q8 = aq + i * QK_K;
        for (int j = 0; j < 2; j++) {
            int8_t* o = a + j * 128;
            for (int l = 0; l < 32; l++) {
                o[l + 0] = (int8_t)((q4[l + 0] & 0xF) | (((qh[l] >> 0) & 3) << 4)) - 32;
                o[l + 32] = (int8_t)((q4[l + 32] & 0xF) | (((qh[l] >> 2) & 3) << 4)) - 32;
                o[l + 64] = (int8_t)((q4[l + 0] >> 4) | (((qh[l] >> 4) & 3) << 4)) - 32;
                o[l + 96] = (int8_t)((q4[l + 32] >> 4) | (((qh[l] >> 6) & 3) << 4)) - 32;
            }
            q4 += 64; qh += 32;
        }
        int32_t aux32[8] = {0};
        for (int j = 0; j < 16; j++) {
            const int32_t s = sc[j];
            for (int r = 0; r < 2; r++) for (int l = 0; l < 8; l++) aux32[l] += s * (int32_t)(q8[j * 16 + r * 8 + l] * a[j * 16 + r * 8 + l]);
        }
        const float d = h2f(dh) * ad[i];
        for (int l = 0; l < 8; l++) sums[l] += d * aux32[l];
    }
    float sumf = 0;
    return lane_sum8(sums, sumf);
}

float vec_dot_q8_0_q8_0(int64_t k, const uint8_t* w, const int8_t* aq, const float* ad) {
    const int64_t nb = k / 32;
    float sumf = 0;
    for (int64_t ii = 0; ii < nb; ii++) {
        const int64_t i = g_alt_order ? nb - 1 - ii : ii;      // noise-floor probe: the same block products added back to front
        const uint8_t* blk = w + i * 34;
        uint16_t dh; memcpy(&dh, blk, 2);
        const int8_t* q = (const int8_t*)(blk + 2);
        int sumi = 0;
        for (int j = 0; j < 32; j++) sumi += q[j] * aq[i * 32 + j];
        sumf += sumi * (h2f(dh) * ad[i]);
    }
    return sumf;
}

// ------------------------------------------------------------------------------------------------
// GGUF v3 reader
// ------------------------------------------------------------------------------------------------
struct Tensor { int type = -1; int nd = 0; int64_t ne[4] = {1, 1, 1, 1}; const uint8_t* data = nullptr; };

struct Reader {
    const uint8_t* p; const uint8_t* end;
    template <class T> T get() { T v; if (p + sizeof(T) > end) throw std::runtime_error("gguf: truncated"); memcpy(&v, p, sizeof(T)); p += sizeof(T); return v; }
    std::string str() { uint64_t n = get<uint64_t>(); if (p + n > end) throw std::runtime_error("gguf: truncated string"); std::string s((const char*)p, n); p += n; return s; }
};

} // namespace

struct orc_model {
    int fd = -1; const uint8_t* base = nullptr; size_t size = 0;
    std::string arch;
    std::map<std::string, double> num;      // numeric metadata
    std::map<std::string, std::string> strs;
    std::map<std::string, Tensor> tensors;
    int n_vocab = 0, n_embd = 0, n_layer = 0, n_head = 0, n_head_kv = 0, d_head = 0, n_ff = 0, n_ctx_train = 0, n_rot = 0;
    float rms_eps = 1e-5f, rope_theta = 10000.f;
    bool neox = false;
    int bos = -1, eos = -1, eot = -1, eom = -1;
    std::vector<int> eog_by_text;           // tokens whose text llama.cpp's vocabulary loader treats as end-of-generation
    const Tensor* get(const std::string& n) const { auto it = tensors.find(n); return it == tensors.end() ? nullptr : &it->second; }
    ~orc_model() { if (base) munmap((void*)base, size); if (fd >= 0) close(fd); }
};

namespace {

void skip_value(Reader& r, uint32_t t, orc_model* m, const std::string& key) {
    static const int sz[] = {1, 1, 2, 2, 4, 4, 4, 1, 0, 0, 8, 8, 8};
    switch (t) {
        case 0: m->num[key] = r.get<uint8_t>(); break;
        case 1: m->num[key] = r.get<int8_t>(); break;
        case 2: m->num[key] = r.get<uint16_t>(); break;
        case 3: m->num[key] = r.get<int16_t>(); break;
        case 4: m->num[key] = r.get<uint32_t>(); break;
        case 5: m->num[key] = r.get<int32_t>(); break;
        case 6: m->num[key] = r.get<float>(); break;
        case 7: m->num[key] = r.get<uint8_t>(); break;
        case 8: m->strs[key] = r.str(); break;
        case 10: m->num[key] = (double)r.get<uint64_t>(); break;
        case 11: m->num[key] = (double)r.get<int64_t>(); break;
        case 12: m->num[key] = r.get<double>(); break;
        case 9: {
            uint32_t et = r.get<uint32_t>(); uint64_t n = r.get<uint64_t>();
            m->num[key + ".count"] = (double)n;
            if (et == 8) {
                // llama-vocab.cpp (special_eog_ids): these token texts end a generation whatever the metadata ids say
                static const char* const eog_texts[] = {"<|eot_id|>", "<|im_end|>", "<|end|>", "<end_of_turn>", "<|endoftext|>", "<|eom_id|>", "<EOT>", "_<EOT>"};
                const bool toks = key == "tokenizer.ggml.tokens";
                for (uint64_t i = 0; i < n; i++) {
                    const std::string s = r.str();
                    if (toks) for (const char* t : eog_texts) if (s == t) m->eog_by_text.push_back((int)i);
                }
            }
            else if (et < 13 && sz[et]) { r.p += n * sz[et]; }
            else throw std::runtime_error("gguf: nested arrays unsupported");
            break;
        }
        default: throw std::runtime_error("gguf: bad value type");
    }
}

} // namespace

extern "C" orc_model* orc_model_load(const char* path) {
    auto* m = new orc_model();
    try {
        m->fd = open(path, O_RDONLY);
        if (m->fd < 0) throw std::runtime_error("cannot open file");
        struct stat st; fstat(m->fd, &st); m->size = st.st_size;
        m->base = (const uint8_t*)mmap(nullptr, m->size, PROT_READ, MAP_PRIVATE, m->fd, 0);
        if (m->base == MAP_FAILED) { m->base = nullptr; throw std::runtime_error("mmap failed"); }
        Reader r{m->base, m->base + m->size};
        if (r.get<uint32_t>() != 0x46554747u) throw std::runtime_error("bad magic");
        uint32_t ver = r.get<uint32_t>(); if (ver < 2 || ver > 3) throw std::runtime_error("unsupported gguf version");
        uint64_t n_t = r.get<uint64_t>(), n_kv = r.get<uint64_t>();
        for (uint64_t i = 0; i < n_kv; i++) { std::string k = r.str(); uint32_t t = r.get<uint32_t>(); skip_value(r, t, m, k); }
        struct Info { std::string name; Tensor t; uint64_t off; };
        std::vector<Info> infos(n_t);
        for (auto& in : infos) {
            in.name = r.str(); in.t.nd = r.get<uint32_t>();
            for (int d = 0; d < in.t.nd; d++) in.t.ne[d] = (int64_t)r.get<uint64_t>();
            in.t.type = (int)r.get<uint32_t>(); in.off = r.get<uint64_t>();
        }
        size_t align = m->num.count("general.alignment") ? (size_t)m->num["general.alignment"] : 32;
        size_t data0 = ((size_t)(r.p - m->base) + align - 1) / align * align;
        for (auto& in : infos) { in.t.data = m->base + data0 + in.off; m->tensors[in.name] = in.t; }
        m->arch = m->strs["general.architecture"];
        auto hp = [&](const std::string& k, double def = -1) { auto it = m->num.find(m->arch + "." + k); if (it == m->num.end()) { if (def < 0) throw std::runtime_error("missing key " + k); return def; } return it->second; };
        m->n_embd = (int)hp("embedding_length"); m->n_layer = (int)hp("block_count"); m->n_ff = (int)hp("feed_forward_length");
        m->n_head = (int)hp("attention.head_count"); m->n_head_kv = (int)hp("attention.head_count_kv", m->n_head);
        m->n_ctx_train = (int)hp("context_length"); m->rms_eps = (float)hp("attention.layer_norm_rms_epsilon", 1e-5);
        m->rope_theta = (float)hp("rope.freq_base", 10000.0);
        m->d_head = m->n_embd / m->n_head;
        m->n_rot = (int)hp("rope.dimension_count", m->d_head);
        m->neox = (m->arch == "qwen2");
        if (m->arch != "llama" && m->arch != "qwen2") throw std::runtime_error("unsupported architecture " + m->arch);
        const Tensor* te = m->get("token_embd.weight"); if (!te) throw std::runtime_error("no token_embd");
        m->n_vocab = (int)te->ne[1];
        auto tk = [&](const char* k) { auto it = m->num.find(std::string("tokenizer.ggml.") + k); return it == m->num.end() ? -1 : (int)it->second; };
        m->bos = tk("bos_token_id"); m->eos = tk("eos_token_id"); m->eot = tk("eot_token_id"); m->eom = tk("eom_token_id");
    } catch (const std::exception& e) {
        fprintf(stderr, "orc_model_load(%s): %s\n", path, e.what());
        delete m; return nullptr;
    }
    return m;
}
extern "C" void orc_model_free(orc_model* m) { delete m; }
extern "C" int32_t orc_n_vocab(const orc_model* m) { return m->n_vocab; }
extern "C" int32_t orc_n_ctx_train(const orc_model* m) { return m->n_ctx_train; }
extern "C" int32_t orc_n_embd(const orc_model* m) { return m->n_embd; }
extern "C" int32_t orc_n_layer(const orc_model* m) { return m->n_layer; }
extern "C" int32_t orc_token_bos(const orc_model* m) { return m->bos; }
extern "C" int32_t orc_is_eog(const orc_model* m, int32_t t) {
    if (t < 0) return 0;
    if (t == m->eos || t == m->eot || t == m->eom) return 1;
    for (int e : m->eog_by_text) if (e == t) return 1;
    return 0;
}
extern "C" int64_t orc_weight_bytes_per_token(const orc_model* m) {
    int64_t tot = 0;
    for (auto& kv : m->tensors) {
        if (kv.first == "token_embd.weight" && m->get("output.weight")) continue;
        if (kv.first == "rope_freqs.weight") continue;
        const Tensor& t = kv.second;
        tot += (int64_t)row_bytes(t.type, t.ne[0]) * t.ne[1] * t.ne[2];
    }
    return tot;
}

// ------------------------------------------------------------------------------------------------
// context + forward
// ------------------------------------------------------------------------------------------------
struct orc_ctx {
    orc_model* m; int n_ctx; int mode; int n_threads;
    std::unique_ptr<ThreadPool> pool;
    int n_past = 0;
    void par(int64_t total, int64_t grain, const std::function<void(int64_t, int64_t)>& fn) const {
        if (pool) pool->run(total, grain, fn); else fn(0, total);
    }
    std::vector<std::vector<uint16_t>> kc, vc;     // per layer [n_ctx][n_head_kv*d_head] f16
    // Self-Extend: position of every cell (empty = its index) and the rotation not yet applied to its K row; rope_off = position of
    // the next token - n_past (llama_batch_allocr gives a batch without positions pos_max + 1 ...)
    std::vector<int> cell_pos, cell_shift; int rope_off = 0;
    std::vector<float> logits; int n_logit_rows = 0; int last_n = 0;
    std::vector<float> hidden;
};

namespace {

// Y[t][r] = W[r,:] . X[t,:]
void matmul(const orc_ctx* c, const Tensor& W, const float* X, int n_tok, float* Y) {
    const int64_t K = W.ne[0], N = W.ne[1];
    const size_t rb = row_bytes(W.type, K);
    const int mode = (c->mode == 3) ? ORC_MODE_GGML : c->mode;
    g_alt_order = (c->mode == 3);
    const bool kq = (W.type == T_Q4_K || W.type == T_Q5_K || W.type == T_Q6_K);
    if (mode == ORC_MODE_GGML && (kq || W.type == T_Q8_0)) {
        const int64_t nblk = kq ? K / QK_K : K / 32;
        std::vector<int8_t> aq((size_t)n_tok * K); std::vector<float> ad((size_t)n_tok * nblk); std::vector<int16_t> bs(kq ? (size_t)n_tok * (K / 16) : 0);
        c->par(n_tok, 1, [&](int64_t t0, int64_t t1) { for (int64_t t = t0; t < t1; t++) {
            if (kq) quantize_q8_K(X + (size_t)t * K, K, aq.data() + (size_t)t * K, ad.data() + (size_t)t * nblk, bs.data() + (size_t)t * (K / 16));
            else quantize_q8_0(X + (size_t)t * K, K, aq.data() + (size_t)t * K, ad.data() + (size_t)t * nblk);
        } });
        c->par(N, 32, [&](int64_t r0, int64_t r1) { for (int64_t r = r0; r < r1; r++) {
            const uint8_t* w = W.data + r * rb;
            for (int t = 0; t < n_tok; t++) {
                const int8_t* q = aq.data() + (size_t)t * K; const float* d = ad.data() + (size_t)t * nblk; const int16_t* b = kq ? bs.data() + (size_t)t * (K / 16) : nullptr;
                float v;
                switch (W.type) {
                    case T_Q4_K: v = vec_dot_q4_K_q8_K(K, w, q, d, b); break;
                    case T_Q5_K: v = vec_dot_q5_K_q8_K(K, w, q, d, b); break;
                    case T_Q6_K: v = vec_dot_q6_K_q8_K(K, w, q, d); break;
                    default: v = vec_dot_q8_0_q8_0(K, w, q, d); break;
                }
                Y[(size_t)t * N + r] = v;
            }
        } });
        return;
    }
    // float paths: F32 (ideal), BF16 (tensor-core operand rounding), and GGML mode on F16/F32 weights
    std::vector<float> Xr;
    const float* Xs = X;
    if (mode == ORC_MODE_BF16 || (mode == ORC_MODE_GGML && W.type == T_F16)) {
        Xr.resize((size_t)n_tok * K);
        for (size_t i = 0; i < Xr.size(); i++) Xr[i] = (mode == ORC_MODE_BF16) ? bf16_round(X[i]) : h2f(f2h(X[i]));
        Xs = Xr.data();
    }
    c->par(N, 32, [&](int64_t r0, int64_t r1) {
        std::vector<float> wrow(K);
        for (int64_t r = r0; r < r1; r++) {
            dequant_row(W.type, W.data + r * rb, K, wrow.data());
            if (mode == ORC_MODE_BF16) for (int64_t i = 0; i < K; i++) wrow[i] = bf16_round(wrow[i]);
            for (int t = 0; t < n_tok; t++) {
                const float* x = Xs + (size_t)t * K;
                double acc = 0;
                for (int64_t i = 0; i < K; i++) acc += (double)wrow[i] * (double)x[i];
                Y[(size_t)t * N + r] = (float)acc;
            }
        }
    });
}

// ggml_compute_forward_rms_norm_f32 followed by ggml_mul with the norm weight
void rms_norm_mul(const float* x, const float* w, int n, float eps, float* y) {
    double sum = 0.0;
    for (int i = 0; i < n; i++) sum += (double)(x[i] * x[i]);
    const float mean = (float)(sum / n);
    const float scale = 1.0f / sqrtf(mean + eps);
    for (int i = 0; i < n; i++) y[i] = (x[i] * scale) * w[i];
}

// ggml rope (ext_factor = 0, freq_scale = 1, attn_factor = 1): theta_i = pos * theta_scale^i / freq_factor_i
void rope(float* v, int n_heads, int d_head, int n_rot, int pos, float theta_base, const float* ff, bool neox) {
    const float theta_scale = powf(theta_base, -2.0f / n_rot);
    std::vector<float> cs(n_rot);
    float theta = (float)pos;
    for (int i0 = 0; i0 < n_rot; i0 += 2) {
        const float f = ff ? ff[i0 / 2] : 1.0f;
        const float th = theta / f;
        cs[i0] = cosf(th); cs[i0 + 1] = sinf(th);
        theta *= theta_scale;
    }
    for (int h = 0; h < n_heads; h++) {
        float* x = v + (size_t)h * d_head;
        for (int i0 = 0; i0 < n_rot; i0 += 2) {
            const float c = cs[i0], s = cs[i0 + 1];
            if (!neox) { const float x0 = x[i0], x1 = x[i0 + 1]; x[i0] = x0 * c - x1 * s; x[i0 + 1] = x0 * s + x1 * c; }
            else { const int ic = i0 / 2; const float x0 = x[ic], x1 = x[ic + n_rot / 2]; x[ic] = x0 * c - x1 * s; x[ic + n_rot / 2] = x0 * s + x1 * c; }
        }
    }
}

// llama.cpp's K-shift: before the next decode every cell whose position changed gets its K row rotated by the accumulated delta
void orc_apply_pos_shift(orc_ctx* c) {
    if (c->cell_shift.empty()) return;
    const orc_model* m = c->m;
    const int dh = m->d_head, nkv = m->n_head_kv, dkv = nkv * dh;
    const Tensor* rf = m->get("rope_freqs.weight");
    const float* ffac = rf ? (const float*)rf->data : nullptr;
    std::vector<float> row(dkv);
    for (int t = 0; t < c->n_past; t++) {
        const int dl = c->cell_shift[t];
        if (!dl) continue;
        for (int l = 0; l < m->n_layer; l++) {
            uint16_t* k = &c->kc[l][(size_t)t * dkv];
            for (int i = 0; i < dkv; i++) row[i] = h2f(k[i]);
            rope(row.data(), nkv, dh, m->n_rot, dl, m->rope_theta, ffac, m->neox);
            for (int i = 0; i < dkv; i++) k[i] = f2h(row[i]);
        }
        c->cell_shift[t] = 0;
    }
}

int forward(orc_ctx* c, const int32_t* tokens, int n, bool all_logits) {
    orc_model* m = c->m;
    if (n <= 0 || c->n_past + n > c->n_ctx) return 1;
    for (int i = 0; i < n; i++) if (tokens[i] < 0 || tokens[i] >= m->n_vocab) return 2;
    const int d = m->n_embd, dh = m->d_head, nh = m->n_head, nkv = m->n_head_kv, dq = nh * dh, dkv = nkv * dh, ff = m->n_ff;
    const int pos0 = c->n_past;                      // cell index of the first new token
    const int rpos0 = c->n_past + c->rope_off;       // its rotary position
    orc_apply_pos_shift(c);
    if (!c->cell_pos.empty()) for (int t = 0; t < n; t++) { c->cell_pos[pos0 + t] = rpos0 + t; c->cell_shift[pos0 + t] = 0; }
    const Tensor* te = m->get("token_embd.weight");
    std::vector<float> X((size_t)n * d), H((size_t)n * d), Q((size_t)n * dq), Kc((size_t)n * dkv), Vc((size_t)n * dkv), A((size_t)n * dq),
        X2((size_t)n * d), G((size_t)n * ff), U((size_t)n * ff), T((size_t)n * std::max(d, dq));
    for (int t = 0; t < n; t++) dequant_row(te->type, te->data + (size_t)tokens[t] * row_bytes(te->type, d), d, X.data() + (size_t)t * d);
    const Tensor* rf = m->get("rope_freqs.weight");
    const float* ffac = rf ? (const float*)rf->data : nullptr;
    const float kq_scale = 1.0f / sqrtf((float)dh);
    for (int l = 0; l < m->n_layer; l++) {
        const std::string p = "blk." + std::to_string(l) + ".";
        auto W = [&](const char* s) { const Tensor* t = m->get(p + s); if (!t) throw std::runtime_error("missing tensor " + p + s); return t; };
        const float* an = (const float*)W("attn_norm.weight")->data;
        for (int t = 0; t < n; t++) rms_norm_mul(X.data() + (size_t)t * d, an, d, m->rms_eps, H.data() + (size_t)t * d);
        matmul(c, *W("attn_q.weight"), H.data(), n, Q.data());
        matmul(c, *W("attn_k.weight"), H.data(), n, Kc.data());
        matmul(c, *W("attn_v.weight"), H.data(), n, Vc.data());
        if (const Tensor* b = m->get(p + "attn_q.bias")) for (int t = 0; t < n; t++) for (int i = 0; i < dq; i++) Q[(size_t)t * dq + i] += ((const float*)b->data)[i];
        if (const Tensor* b = m->get(p + "attn_k.bias")) for (int t = 0; t < n; t++) for (int i = 0; i < dkv; i++) Kc[(size_t)t * dkv + i] += ((const float*)b->data)[i];
        if (const Tensor* b = m->get(p + "attn_v.bias")) for (int t = 0; t < n; t++) for (int i = 0; i < dkv; i++) Vc[(size_t)t * dkv + i] += ((const float*)b->data)[i];
        for (int t = 0; t < n; t++) {
            rope(Q.data() + (size_t)t * dq, nh, dh, m->n_rot, rpos0 + t, m->rope_theta, ffac, m->neox);
            rope(Kc.data() + (size_t)t * dkv, nkv, dh, m->n_rot, rpos0 + t, m->rope_theta, ffac, m->neox);
        }
        // KV cache store: f32 -> f16 (ggml_cpy)
        for (int t = 0; t < n; t++) for (int i = 0; i < dkv; i++) {
            c->kc[l][(size_t)(pos0 + t) * dkv + i] = f2h(Kc[(size_t)t * dkv + i]);
            c->vc[l][(size_t)(pos0 + t) * dkv + i] = f2h(Vc[(size_t)t * dkv + i]);
        }
        // attention (llama-graph.cpp build_attn_mha, flash_attn off): K.q with q rounded to f16 (F16 vec_dot type),
        // f32 scores, soft_max_ext, probabilities rounded to f16 for the V product
        const int gq = nh / nkv;
        c->par((int64_t)n * nh, 1, [&](int64_t i0, int64_t i1) { for (int64_t it = i0; it < i1; it++) {
            const int t = (int)(it / nh), h = (int)(it % nh);
            const int n_kv = pos0 + t + 1; const int hk = h / gq;
            std::vector<float> s(n_kv), qh(dh);
            for (int i = 0; i < dh; i++) qh[i] = h2f(f2h(Q[(size_t)t * dq + (size_t)h * dh + i]));
            float mx = -INFINITY;
            for (int j = 0; j < n_kv; j++) {
                const uint16_t* kr = &c->kc[l][(size_t)j * dkv + (size_t)hk * dh];
                float acc = 0;
                for (int i = 0; i < dh; i++) acc += h2f(kr[i]) * qh[i];
                s[j] = acc * kq_scale; mx = std::max(mx, s[j]);
            }
            double sum = 0;
            for (int j = 0; j < n_kv; j++) { s[j] = expf(s[j] - mx); sum += (double)s[j]; }
            const float inv = (float)(1.0 / sum);
            for (int j = 0; j < n_kv; j++) s[j] = h2f(f2h(s[j] * inv));
            float* o = A.data() + (size_t)t * dq + (size_t)h * dh;
            for (int i = 0; i < dh; i++) o[i] = 0;
            for (int j = 0; j < n_kv; j++) {
                const uint16_t* vr = &c->vc[l][(size_t)j * dkv + (size_t)hk * dh]; const float pj = s[j];
                for (int i = 0; i < dh; i++) o[i] += h2f(vr[i]) * pj;
            }
        } });
        matmul(c, *W("attn_output.weight"), A.data(), n, T.data());
        for (size_t i = 0; i < (size_t)n * d; i++) X2[i] = T[i] + X[i];
        const float* fn = (const float*)W("ffn_norm.weight")->data;
        for (int t = 0; t < n; t++) rms_norm_mul(X2.data() + (size_t)t * d, fn, d, m->rms_eps, H.data() + (size_t)t * d);
        matmul(c, *W("ffn_gate.weight"), H.data(), n, G.data());
        matmul(c, *W("ffn_up.weight"), H.data(), n, U.data());
        for (size_t i = 0; i < (size_t)n * ff; i++) { const float g = G[i]; G[i] = (g / (1.0f + expf(-g))) * U[i]; }
        matmul(c, *W("ffn_down.weight"), G.data(), n, T.data());
        for (size_t i = 0; i < (size_t)n * d; i++) X[i] = T[i] + X2[i];
    }
    const float* on = (const float*)m->get("output_norm.weight")->data;
    const Tensor* wo = m->get("output.weight"); if (!wo) wo = te;
    const int t0 = all_logits ? 0 : n - 1, nt = n - t0;
    c->hidden.resize((size_t)nt * d);
    for (int t = 0; t < nt; t++) rms_norm_mul(X.data() + (size_t)(t0 + t) * d, on, d, m->rms_eps, c->hidden.data() + (size_t)t * d);
    c->logits.resize((size_t)nt * m->n_vocab);
    matmul(c, *wo, c->hidden.data(), nt, c->logits.data());
    c->n_logit_rows = nt; c->last_n = n;
    c->n_past += n;
    return 0;
}

} // namespace

extern "C" orc_ctx* orc_ctx_create(orc_model* m, int32_t n_ctx, int32_t mode, int32_t n_threads) {
    if (!m) return nullptr;
    auto* c = new orc_ctx{m, n_ctx > 0 ? n_ctx : m->n_ctx_train, mode, n_threads > 0 ? n_threads : 1};
    if (c->n_threads > 1) c->pool = std::make_unique<ThreadPool>(c->n_threads);
    const size_t dkv = (size_t)m->n_head_kv * m->d_head;
    c->kc.assign(m->n_layer, std::vector<uint16_t>((size_t)c->n_ctx * dkv));
    c->vc.assign(m->n_layer, std::vector<uint16_t>((size_t)c->n_ctx * dkv));
    return c;
}
extern "C" void orc_ctx_free(orc_ctx* c) { delete c; }
extern "C" void orc_kv_clear(orc_ctx* c) { c->n_past = 0; c->cell_pos.clear(); c->cell_shift.clear(); c->rope_off = 0; }

// Self-Extend (reference Session.cpp:348-368): llama_kv_self_seq_add / seq_div on the POSITIONS of the cells (llama-kv-cache.cpp:
// pos += delta / pos /= d, the change accumulated in cell.delta and applied to K by the next decode's K-shift)
namespace {
void orc_positions(orc_ctx* c) {
    if (!c->cell_pos.empty()) return;
    c->cell_pos.resize(c->n_ctx); c->cell_shift.assign(c->n_ctx, 0);
    for (int i = 0; i < c->n_ctx; i++) c->cell_pos[i] = i;
}
void orc_refresh_off(orc_ctx* c) {
    int mx = -1;
    for (int i = 0; i < c->n_past; i++) mx = std::max(mx, c->cell_pos[i]);
    c->rope_off = mx + 1 - c->n_past;
}
} // namespace
extern "C" int32_t orc_kv_seq_add(orc_ctx* c, int32_t p0, int32_t p1, int32_t delta) {
    if (p0 < 0 || p1 < p0) return 1;
    orc_positions(c);
    for (int i = 0; i < c->n_past; i++) if (c->cell_pos[i] >= p0 && c->cell_pos[i] < p1) { c->cell_pos[i] += delta; c->cell_shift[i] += delta; }
    orc_refresh_off(c);
    return 0;
}
extern "C" int32_t orc_kv_seq_div(orc_ctx* c, int32_t p0, int32_t p1, int32_t d) {
    if (p0 < 0 || p1 < p0 || d <= 0) return 1;
    orc_positions(c);
    for (int i = 0; i < c->n_past; i++) if (c->cell_pos[i] >= p0 && c->cell_pos[i] < p1) { const int old = c->cell_pos[i]; c->cell_pos[i] /= d; c->cell_shift[i] += c->cell_pos[i] - old; }
    orc_refresh_off(c);
    return 0;
}
extern "C" int32_t orc_next_pos(const orc_ctx* c) { return c->n_past + c->rope_off; }

// Context shift (reference Session.cpp:341-342): llama_kv_self_seq_rm(ctx, 0, p0, p1) drops the cells of positions [p0, p1);
// llama_kv_self_seq_add(ctx, 0, p1, n_past, -(p1 - p0)) moves the positions of the cells behind them down, which llama.cpp applies
// before the next decode as a K-shift: ggml_rope_ext_inplace on the F16 K cache with the position DELTA of every cell
// (llama-context.cpp build_rope_shift; ggml-cpu rope on F16: f16 -> f32, rotate, f32 -> f16).  V rows only move.
extern "C" int32_t orc_kv_shift(orc_ctx* c, int32_t p0, int32_t p1) {
    if (p0 < 0 || p1 <= p0 || p1 > c->n_past) return 1;
    const orc_model* m = c->m;
    const int d = p1 - p0, dh = m->d_head, nkv = m->n_head_kv, dkv = nkv * dh;
    const Tensor* rf = m->get("rope_freqs.weight");
    const float* ffac = rf ? (const float*)rf->data : nullptr;
    std::vector<float> row(dkv);
    for (int l = 0; l < m->n_layer; l++) {
        for (int t = p1; t < c->n_past; t++) {
            const uint16_t* src = &c->kc[l][(size_t)t * dkv];
            for (int i = 0; i < dkv; i++) row[i] = h2f(src[i]);
            rope(row.data(), nkv, dh, m->n_rot, -d, m->rope_theta, ffac, m->neox);
            uint16_t* dst = &c->kc[l][(size_t)(t - d) * dkv];
            for (int i = 0; i < dkv; i++) dst[i] = f2h(row[i]);
            memcpy(&c->vc[l][(size_t)(t - d) * dkv], &c->vc[l][(size_t)t * dkv], (size_t)dkv * 2);
        }
    }
    c->n_past -= d;
    return 0;
}
extern "C" int32_t orc_n_past(const orc_ctx* c) { return c->n_past; }
extern "C" int32_t orc_decode(orc_ctx* c, const int32_t* tokens, int32_t n, int32_t all_logits) {
    try { return forward(c, tokens, n, all_logits != 0); }
    catch (const std::exception& e) { fprintf(stderr, "orc_decode: %s\n", e.what()); return 3; }
}
extern "C" const float* orc_get_logits(orc_ctx* c, int32_t i) {
    if (c->n_logit_rows == 0) return nullptr;
    int row = (c->n_logit_rows == 1) ? 0 : (i < 0 ? c->n_logit_rows + i : i);
    if (row < 0 || row >= c->n_logit_rows) return nullptr;
    return c->logits.data() + (size_t)row * c->m->n_vocab;
}
extern "C" const float* orc_get_hidden(orc_ctx* c, int32_t i) {
    if (c->n_logit_rows == 0) return nullptr;
    int row = (c->n_logit_rows == 1) ? 0 : (i < 0 ? c->n_logit_rows + i : i);
    return c->hidden.data() + (size_t)row * c->m->n_embd;
}

// ------------------------------------------------------------------------------------------------
// top-k / gather  (Session.cpp:246-282)
// ------------------------------------------------------------------------------------------------
extern "C" void orc_topk(const float* logits, int32_t n_vocab, int32_t k, orc_token_data* out) {
    std::vector<orc_token_data> v(n_vocab);
    for (int i = 0; i < n_vocab; i++) v[i] = {i, logits[i]};
    k = std::min(k, n_vocab);
    // the reference sorts the whole vocabulary with comparator a.logit > b.logit; ties are unspecified there,
    // resolved here by lower id first
    std::partial_sort(v.begin(), v.begin() + k, v.end(), [](const orc_token_data& a, const orc_token_data& b) {
        return a.logit > b.logit || (a.logit == b.logit && a.token < b.token); });
    for (int i = 0; i < k; i++) out[i] = v[i];
}

extern "C" int32_t orc_gather_sorted(const float* logits, int32_t n_vocab, const int32_t* ids, int32_t n_ids, orc_token_data* out) {
    // fillLogits(pred) walks the vocabulary in id order and keeps ids present in the claimed set (each once)
    std::vector<int32_t> u(ids, ids + n_ids);
    std::sort(u.begin(), u.end()); u.erase(std::unique(u.begin(), u.end()), u.end());
    int n = 0;
    for (int32_t id : u) if (id >= 0 && id < n_vocab) out[n++] = {id, logits[id]};
    std::stable_sort(out, out + n, [](const orc_token_data& a, const orc_token_data& b) { return a.logit > b.logit; });
    return n;
}

// ------------------------------------------------------------------------------------------------
// sampler chain (llama-sampling.cpp; configured as in reference Sampler.cpp:15-97)
// ------------------------------------------------------------------------------------------------
struct orc_sampler {
    uint32_t seed; float temp, top_p, min_p; int top_k; size_t min_keep;
    std::mt19937 rng;
    struct TD { int id; float logit; float p; };
    std::vector<TD> cur;
};

namespace {
using TD = orc_sampler::TD;

void softmax_impl(std::vector<TD>& c, size_t size, bool& sorted) {
    if (!sorted) { std::sort(c.begin(), c.begin() + size, [](const TD& a, const TD& b) { return a.logit > b.logit; }); sorted = true; }
    const float max_l = c[0].logit; float cum = 0.0f;
    for (size_t i = 0; i < size; i++) { float p = expf(c[i].logit - max_l); c[i].p = p; cum += p; }
    for (size_t i = 0; i < size; i++) c[i].p /= cum;
}

int run_chain(orc_sampler* s, size_t size, bool sorted) {
    auto& c = s->cur;
    // logit_bias (empty) and penalties (repeat 1.0 / freq 0 / present 0) are no-ops with blama's defaults
    // top-k
    {
        int k = s->top_k; if (k <= 0) k = (int)size; k = std::min(k, (int)size);
        if (!sorted) {
            std::partial_sort(c.begin(), c.begin() + k, c.begin() + size, [](const TD& a, const TD& b) { return a.logit > b.logit; });
            sorted = true;
        }
        size = k;
    }
    // typical(p = 1.0): no-op
    // top-p
    if (s->top_p < 1.0f) {
        softmax_impl(c, size, sorted);
        float cum = 0.0f; size_t last = size;
        for (size_t i = 0; i < size; i++) { cum += c[i].p; if (cum >= s->top_p && i + 1 >= s->min_keep) { last = i + 1; break; } }
        size = last;
    }
    // min-p (sorted branch)
    if (s->min_p > 0.0f && size > 0) {
        const float min_logit = c[0].logit + logf(s->min_p);
        size_t i = 1;
        for (; i < size; i++) if (c[i].logit < min_logit && i >= s->min_keep) break;
        size = i;
    }
    // temp_ext with delta 0 == temp
    if (s->temp <= 0.0f) {
        size_t best = 0; for (size_t i = 1; i < size; i++) if (c[i].logit > c[best].logit) best = i;
        for (size_t i = 0; i < size; i++) if (i != best) c[i].logit = -INFINITY;
    } else {
        for (size_t i = 0; i < size; i++) c[i].logit /= s->temp;
    }
    // dist
    softmax_impl(c, size, sorted);
    std::vector<double> probs(size);
    for (size_t i = 0; i < size; i++) probs[i] = c[i].p;
    std::discrete_distribution<int> dist(probs.begin(), probs.end());
    return c[dist(s->rng)].id;
}
} // namespace

extern "C" orc_sampler* orc_sampler_create_ex(uint32_t seed, float temp, float top_p, int32_t top_k, float min_p, int32_t min_keep) {
    auto* s = new orc_sampler{seed, temp, top_p, min_p, top_k, (size_t)min_keep, std::mt19937(seed), {}};
    return s;
}
extern "C" orc_sampler* orc_sampler_create(uint32_t seed, float temp, float top_p) { return orc_sampler_create_ex(seed, temp, top_p, 40, 0.05f, 0); }
extern "C" void orc_sampler_free(orc_sampler* s) { delete s; }
extern "C" void orc_sampler_reset(orc_sampler* s) { s->rng.seed(s->seed); }
extern "C" int32_t orc_sampler_sample(orc_sampler* s, const float* logits, int32_t n_vocab) {
    s->cur.resize(n_vocab);
    for (int i = 0; i < n_vocab; i++) s->cur[i] = {i, logits[i], 0.0f};
    return run_chain(s, n_vocab, false);
}
extern "C" int32_t orc_sampler_sample_candidates(orc_sampler* s, const orc_token_data* cand, int32_t n) {
    s->cur.resize(n);
    for (int i = 0; i < n; i++) s->cur[i] = {cand[i].token, cand[i].logit, 0.0f};
    return run_chain(s, n, true);
}

// ------------------------------------------------------------------------------------------------
// Session control flow (Session.cpp)
// ------------------------------------------------------------------------------------------------
extern "C" int32_t orc_session_complete(orc_ctx* c, const int32_t* prompt, int32_t n_prompt, int32_t max_tokens,
                                        uint32_t seed, float temp, float top_p, int32_t* out_tokens, orc_token_data* out_top10) {
    orc_kv_clear(c);                                   // Session ctor, Session.cpp:53
    int32_t bos = c->m->bos;
    if (n_prompt == 0) { prompt = &bos; n_prompt = 1; } // Session.cpp:76-79
    if (orc_decode(c, prompt, n_prompt, 0)) return -1;  // setInitialPrompt -> doDecode
    orc_sampler* s = orc_sampler_create(seed, temp, top_p);
    int n_out = 0;
    for (int i = 0; i < max_tokens; i++) {
        // getToken: sample from the pending logits (Session.cpp:179) ...
        int32_t tok = orc_sampler_sample(s, orc_get_logits(c, -1), c->m->n_vocab);
        if (orc_is_eog(c->m, tok)) break;               // :181-184, :206-208
        // ... then getLogitsFromCtx(10) first decodes the sampled token (flushPendingState, :252) and sorts
        if (orc_decode(c, &tok, 1, 0)) { orc_sampler_free(s); return -1; }
        out_tokens[n_out] = tok;
        orc_topk(orc_get_logits(c, -1), c->m->n_vocab, 10, out_top10 + (size_t)n_out * 10);
        n_out++;
    }
    orc_sampler_free(s);
    return n_out;
}

extern "C" int32_t orc_session_fill_ctx(orc_ctx* c, const int32_t* prompt, int32_t n_prompt, const int32_t* resp, int32_t n_resp,
                                        const int32_t* claimed, const int32_t* n_claimed, orc_token_data* out, int32_t* out_n) {
    orc_kv_clear(c);
    int32_t bos = c->m->bos;
    if (n_prompt == 0) { prompt = &bos; n_prompt = 1; }
    if (orc_decode(c, prompt, n_prompt, 0)) return -1;
    for (int i = 0; i < n_resp; i++) {                 // Session.cpp:235-241: one single-token decode per response token
        if (orc_decode(c, resp + i, 1, 0)) return -1;
        out_n[i] = orc_gather_sorted(orc_get_logits(c, -1), c->m->n_vocab, claimed + (size_t)i * 10, n_claimed[i], out + (size_t)i * 10);
    }
    return 0;
}

// ------------------------------------------------------------------------------------------------
// LogitComparer (LogitComparer.cpp).  The reference accumulates float sums while iterating std::unordered_map,
// so the iteration order of libstdc++'s hash table is part of the arithmetic; the same container, bucket hint and
// insertion sequence are used here so results agree bit for bit with oracle/_ref.
// ------------------------------------------------------------------------------------------------
namespace {
using ProbMap = std::unordered_map<int32_t, float>;

ProbMap lc_softmax(const orc_token_data* d, int n) {
    ProbMap r((size_t)n);
    const float mx = d[0].logit;           // element 0 is taken as the max (LogitComparer.cpp:12)
    float sum = 0.0f;
    for (int i = 0; i < n; i++) { const float e = std::exp(d[i].logit - mx); r[d[i].token] = e; sum += e; }
    for (auto& kv : r) kv.second /= sum;
    return r;
}
float lc_kl(const ProbMap& P, const ProbMap& Q) {
    float kl = 0.0f;
    for (const auto& [tok, p] : P) {
        if (p > 0.0f) { auto it = Q.find(tok); if (it != Q.end() && it->second > 0.0f) kl += p * std::log(p / it->second); }
    }
    return kl;
}
float lc_jsd(const ProbMap& p1, const ProbMap& p2) {
    ProbMap avg;
    for (const auto& [tok, p] : p1) { auto it = p2.find(tok); if (it != p2.end()) avg[tok] = (p + it->second) / 2.0f; }
    return (lc_kl(p1, avg) + lc_kl(p2, avg)) / 2.0f;
}
float lc_sumsq(const orc_token_data* d, size_t n) { float s = 0.0f; for (size_t i = 0; i < n; i++) s += d[i].logit * d[i].logit; return s; }
} // namespace

extern "C" orc_metrics orc_lc_compare(const orc_token_data* a, int32_t na, const orc_token_data* b, int32_t nb) {
    orc_metrics m;
    m.top1Match = a[0].token == b[0].token ? 1.0f : 0.0f;
    const size_t mn = (size_t)std::min(na, nb);
    const float d1 = lc_sumsq(a, mn), d2 = lc_sumsq(b, mn);
    m.distance = std::fabs(d1 - d2) / std::max(d1, d2);
    m.jsd = lc_jsd(lc_softmax(a, na), lc_softmax(b, nb));
    return m;
}
extern "C" float orc_lc_similarity(const orc_token_data* a, int32_t na, const orc_token_data* b, int32_t nb) {
    ProbMap l2;
    for (int i = 0; i < nb; i++) l2[b[i].token] = b[i].logit;
    float wsum = 0.0f, wtot = 0.0f;
    for (int i = 0; i < na; i++) {
        const float w = std::abs(a[i].logit); float sim = 0.0f;
        auto it = l2.find(a[i].token);
        if (it != l2.end()) sim = 1 - (std::abs(a[i].logit - it->second) / std::abs(std::max(a[i].logit, it->second)));
        wsum += w * sim; wtot += w;
    }
    return wtot > 0.0f ? wsum / wtot : 0.0f;
}
extern "C" float orc_lc_score(const orc_metrics* m, int32_t n) {
    // MetricsAggregator::pushAndVerify re-sums the whole history in double on every push; the last push's value is
    // what Server::verify returns (Server.cpp:151-157)
    double total = 0.0;
    for (int i = 0; i < n; i++) total += 0.5 * (1.0f - m[i].distance) + 0.5 * (1.0f - m[i].jsd);
    return n ? float(total / n) : 0.0f;
}

// ------------------------------------------------------------------------------------------------
// unit-level entry points
// ------------------------------------------------------------------------------------------------
extern "C" int32_t orc_dequantize(int32_t type, const void* blocks, int64_t n, float* out) {
    TypeInfo ti{}; if (!type_info(type, ti) || n % ti.blck) return 1;
    dequant_row(type, (const uint8_t*)blocks, n, out); return 0;
}
extern "C" int32_t orc_matvec(int32_t type, const void* w, int64_t rows, int64_t k, const float* x, float* y, int32_t mode) {
    TypeInfo ti{}; if (!type_info(type, ti) || k % ti.blck) return 1;
    orc_ctx c{nullptr, 0, mode, 1};
    Tensor W; W.type = type; W.nd = 2; W.ne[0] = k; W.ne[1] = rows; W.data = (const uint8_t*)w;
    matmul(&c, W, x, 1, y);
    return 0;
}
extern "C" int32_t orc_quantize_q8_K(const float* x, int64_t k, int8_t* qs, float* d, int16_t* bsums) { if (k % QK_K) return 1; quantize_q8_K(x, k, qs, d, bsums); return 0; }
extern "C" int32_t orc_quantize_q8_0(const float* x, int64_t k, int8_t* qs, float* d) { if (k % 32) return 1; quantize_q8_0(x, k, qs, d); return 0; }
