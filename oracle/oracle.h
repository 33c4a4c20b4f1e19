/*
 * oracle.h -- C API of the CPU oracle.  TEST INFRASTRUCTURE ONLY.
 *
 * The oracle is a CPU restatement of the arithmetic blama reaches through
 * llama.cpp b5187 (un-vendored; reference inference/code/CMakeLists.txt:35) plus
 * blama's own control flow (Session.cpp) and verdict (LogitComparer.cpp).
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
 * reference legs may load this library; the product (libblama_b200.so) never
 * links or calls it.
 *
 * Parity pin status (see DESIGN.md "Oracle"):
 *   - LogitComparer / MetricsAggregator : PINNED against the reference's own
 *     LogitComparer.cpp compiled into oracle/_ref (golden vectors in tests/golden).
 *   - block dequantisation              : PINNED against gguf-py's numpy dequantisers.
 *   - transformer forward               : PARITY UNPINNED by reference tests (the reference
 *     only ever tests GPT-2-117M fixtures that are not available); restates ggml-cpu.
 *     Its float path (F32 mode) is pinned to an INDEPENDENT implementation instead: Hugging Face transformers
 *     reading the same GGUF through its own loader (tests/golden/hf_forward_golden.npz, tools/gen_hf_forward_golden.py).
 */
#pragma once
#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct orc_model orc_model;
typedef struct orc_ctx orc_ctx;

/* arithmetic of W.x :  GGML = activations quantised to Q8_K / Q8_0 + integer dot (ggml-cpu);
 *                      F32  = dequantised weights, f32 activations (ideal);
 *                      BF16 = weights and activations rounded to bf16, f32 accumulate
 *                             (what the tcgen05 prefill GEMM computes). */
enum { ORC_MODE_GGML = 0, ORC_MODE_F32 = 1, ORC_MODE_BF16 = 2,
       ORC_MODE_GGML_ALT = 3 /* GGML arithmetic, fp32 lane sums added in reverse order: noise-floor probe */ };

typedef struct { int32_t token; float logit; } orc_token_data;

orc_model* orc_model_load(const char* gguf_path);
void       orc_model_free(orc_model*);
int32_t    orc_n_vocab(const orc_model*);
int32_t    orc_n_ctx_train(const orc_model*);
int32_t    orc_n_embd(const orc_model*);
int32_t    orc_n_layer(const orc_model*);
int32_t    orc_token_bos(const orc_model*);
int32_t    orc_is_eog(const orc_model*, int32_t tok);
int64_t    orc_weight_bytes_per_token(const orc_model*);

orc_ctx* orc_ctx_create(orc_model*, int32_t n_ctx, int32_t mode, int32_t n_threads);
void     orc_ctx_free(orc_ctx*);
void     orc_kv_clear(orc_ctx*);
/* llama_kv_self_seq_rm(p0, p1) + llama_kv_self_seq_add(p1, n_past, -(p1 - p0)) + the K-shift llama.cpp applies before the next decode
 * (reference Session.cpp:341-342) */
int32_t  orc_kv_shift(orc_ctx*, int32_t p0, int32_t p1);
int32_t  orc_n_past(const orc_ctx*);
/* llama_kv_self_seq_add / llama_kv_self_seq_div on cell positions (Self-Extend, reference Session.cpp:359-361) */
int32_t  orc_kv_seq_add(orc_ctx*, int32_t p0, int32_t p1, int32_t delta);
int32_t  orc_kv_seq_div(orc_ctx*, int32_t p0, int32_t p1, int32_t d);
int32_t  orc_next_pos(const orc_ctx*);
/* decode n tokens at positions n_past..; all_logits!=0 keeps logits of every position. returns 0 on success */
int32_t  orc_decode(orc_ctx*, const int32_t* tokens, int32_t n, int32_t all_logits);
/* logits of the i-th token of the last decode (-1 = last); pointer valid until the next decode */
const float* orc_get_logits(orc_ctx*, int32_t i);
/* hidden state (f32, n_embd) after the final norm for token i of the last decode -- debugging aid */
const float* orc_get_hidden(orc_ctx*, int32_t i);

/* Session.cpp:246-261 : top-k of a logits row, sorted by logit descending (ties: lower id first) */
void orc_topk(const float* logits, int32_t n_vocab, int32_t k, orc_token_data* out);
/* Session.cpp:263-282 : the row's logits at the claimed ids (in vocab order, duplicates once), sorted descending.
 * returns the number written */
int32_t orc_gather_sorted(const float* logits, int32_t n_vocab, const int32_t* ids, int32_t n_ids, orc_token_data* out);

/* llama.cpp sampler chain as configured by blama Sampler.cpp:15-97 with default params:
 * top-k 40 -> typical(1.0) -> top-p -> min-p 0.05 -> temp -> dist(seed) over the FULL vocab */
typedef struct orc_sampler orc_sampler;
orc_sampler* orc_sampler_create(uint32_t seed, float temp, float top_p);
orc_sampler* orc_sampler_create_ex(uint32_t seed, float temp, float top_p, int32_t top_k, float min_p, int32_t min_keep);
void         orc_sampler_free(orc_sampler*);
void         orc_sampler_reset(orc_sampler*);
int32_t      orc_sampler_sample(orc_sampler*, const float* logits, int32_t n_vocab);
/* same chain over an already-selected candidate list (ids+logits sorted descending): used to check the
 * product's top-40 shortcut against the full-vocab chain */
int32_t      orc_sampler_sample_candidates(orc_sampler*, const orc_token_data* cand, int32_t n);

/* blama Session restated (Session.cpp:65-107, 169-213, 231-282). top10 buffers are [n][10]. */
int32_t orc_session_complete(orc_ctx*, const int32_t* prompt, int32_t n_prompt, int32_t max_tokens,
                             uint32_t seed, float temp, float top_p,
                             int32_t* out_tokens, orc_token_data* out_top10 /* [max_tokens][10] */);
/* fillCtx: claimed[i][0..n_claimed[i]) are the prover's ids for response token i. verifier output: out[i][..] sorted
 * desc by verifier logit, out_n[i] entries. */
int32_t orc_session_fill_ctx(orc_ctx*, const int32_t* prompt, int32_t n_prompt,
                             const int32_t* resp_tokens, int32_t n_resp,
                             const int32_t* claimed /* [n_resp][10] */, const int32_t* n_claimed,
                             orc_token_data* out /* [n_resp][10] */, int32_t* out_n);

/* LogitComparer.cpp:39-55 / 57-80 / 117-128 restated */
typedef struct { float top1Match, distance, jsd; } orc_metrics;
orc_metrics orc_lc_compare(const orc_token_data* a, int32_t na, const orc_token_data* b, int32_t nb);
float       orc_lc_similarity(const orc_token_data* a, int32_t na, const orc_token_data* b, int32_t nb);
/* running score after pushing metrics[0..n) one at a time (Server.cpp:151-156) */
float       orc_lc_score(const orc_metrics* m, int32_t n);

/* unit-level entry points used to pin the restatement */
/* dequantise n elements of ggml type `type` */
int32_t orc_dequantize(int32_t type, const void* blocks, int64_t n, float* out);
/* y[r] = W[r,:] . x for r < rows, in the given arithmetic mode */
int32_t orc_matvec(int32_t type, const void* w, int64_t rows, int64_t k, const float* x, float* y, int32_t mode);
/* ggml quantize_row_q8_K_ref / q8_0_ref : qs (int8[k]), d (float[k/256] or [k/32] (fp16-rounded)) */
int32_t orc_quantize_q8_K(const float* x, int64_t k, int8_t* qs, float* d, int16_t* bsums);
int32_t orc_quantize_q8_0(const float* x, int64_t k, int8_t* qs, float* d);

#ifdef __cplusplus
}
#endif
