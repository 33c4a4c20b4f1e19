// host_capi.cpp -- C entry points over the bl::llama host classes so that tests (ctypes) and foreign callers can drive
// the reference-shaped API: Model / Instance / Session / LogitComparer / MetricsAggregator / Sampler.
// Every function returns 0 on success, 1 when the C++ layer threw; blh_last_error() then holds e.what() -- the same
// strings the reference's tests pin (inference/test/t-integration.cpp:137-217).
#include "llama/Init.hpp"
#include "llama/Instance.hpp"
#include "llama/LogitComparer.hpp"
#include "llama/Model.hpp"
#include "llama/Sampler.hpp"
#include "llama/Session.hpp"

#include <blama_b200.h>

#include <cstring>
#include <string>

using namespace bl::llama;

namespace {
thread_local std::string g_err;
template <class F> int guard(F&& f) {
    try { f(); return 0; }
    catch (const std::exception& e) { g_err = e.what(); return 1; }
    catch (...) { g_err = "unknown error"; return 1; }
}
struct InstanceBox {
    std::unique_ptr<Instance> inst;
    Session* session = nullptr;
};
TokenDataVector toVec(const blk_token_data* p, int32_t n) {
    TokenDataVector v(static_cast<size_t>(n));
    for (int32_t i = 0; i < n; ++i) v[size_t(i)] = {p[i].token, p[i].logit};
    return v;
}
} // namespace

extern "C" {

BLK_API const char* blh_last_error(void) { return g_err.c_str(); }
BLK_API void blh_init(void) { initLibrary(); }

BLK_API int blh_model_create(const char* path, int device, int gpu, int prefix_bos, void** out) {
    return guard([&] {
        Model::Params p; p.gpu = gpu != 0; p.device = device; p.prefixInputsWithBos = prefix_bos != 0;
        *out = new Model(path, p);
    });
}
BLK_API void blh_model_free(void* m) { delete static_cast<Model*>(m); }
BLK_API int blh_model_train_ctx(void* m) { return int(static_cast<Model*>(m)->trainCtxLength()); }
BLK_API int blh_model_tokenize(void* m, const char* text, int add_special, int32_t* out, int cap) {
    auto v = static_cast<Model*>(m)->vocab().tokenize(text, add_special != 0, true);
    for (int i = 0; i < int(v.size()) && i < cap; ++i) out[i] = v[size_t(i)];
    return int(v.size());
}
BLK_API int blh_model_token_to_string(void* m, int32_t tok, char* buf, int cap) {
    auto s = static_cast<Model*>(m)->vocab().tokenToString(tok);
    const int n = int(s.size());
    memcpy(buf, s.data(), size_t(n < cap ? n : cap));
    return n;
}

BLK_API int blh_instance_create(void* model, uint32_t ctx_size, uint32_t batch_size, void** out) {
    return guard([&] {
        auto box = std::make_unique<InstanceBox>();
        Instance::InitParams ip; ip.ctxSize = ctx_size; if (batch_size) ip.batchSize = batch_size;
        box->inst = std::make_unique<Instance>(*static_cast<Model*>(model), ip);
        *out = box.release();
    });
}
BLK_API void blh_instance_free(void* i) { delete static_cast<InstanceBox*>(i); }
// the C-ABI context behind an Instance (bench.py times it with blk_timer_* / counts its launches)
BLK_API void* blh_instance_ctx(void* i) { return static_cast<InstanceBox*>(i)->inst->lctx(); }
BLK_API int blh_instance_warmup(void* i) { return guard([&] { static_cast<InstanceBox*>(i)->inst->warmup(); }); }

BLK_API int blh_session_start(void* i, uint32_t seed, float temp, float top_p, int sequential_verify) {
    return guard([&] {
        auto* box = static_cast<InstanceBox*>(i);
        Session::InitParams sp; sp.seed = seed; sp.temperature = temp; sp.topP = top_p; sp.sequentialVerify = sequential_verify != 0;
        box->session = &box->inst->startSession(sp);
    });
}
BLK_API void blh_session_stop(void* i) { auto* box = static_cast<InstanceBox*>(i); box->inst->stopSession(); box->session = nullptr; }
BLK_API int blh_session_set_initial_prompt(void* i, const int32_t* toks, int n) {
    return guard([&] { static_cast<InstanceBox*>(i)->session->setInitialPrompt({toks, size_t(n)}); });
}
// out_top10: [max_tokens][10]; out_n_logits[i] = number of valid logits for token i; *out_n = tokens produced
BLK_API int blh_session_complete(void* i, const int32_t* prompt, int n_prompt, int max_tokens, int32_t* out_tokens,
                                 blk_token_data* out_top10, int32_t* out_n_logits, int32_t* out_n) {
    return guard([&] {
        auto preds = static_cast<InstanceBox*>(i)->session->complete({.prompt = {prompt, size_t(n_prompt)}, .suffix = {}, .maxTokens = max_tokens});
        *out_n = int32_t(preds.size());
        for (size_t t = 0; t < preds.size(); ++t) {
            out_tokens[t] = preds[t].token;
            out_n_logits[t] = int32_t(preds[t].logits.size());
            for (size_t j = 0; j < preds[t].logits.size() && j < 10; ++j) out_top10[t * 10 + j] = {preds[t].logits[j].token, preds[t].logits[j].logit};
        }
    });
}
BLK_API int blh_session_stream(void* i, int max_tokens, int32_t* out_tokens, int32_t* out_n) {
    return guard([&] {
        auto gen = static_cast<InstanceBox*>(i)->session->completeStream({.prompt = {}, .suffix = {}, .maxTokens = max_tokens});
        int n = 0;
        while (gen.status() == Session::StreamGenerator::Status::InProgress) {
            auto p = gen.complete();
            if (!p) break;
            out_tokens[n++] = p.token;
        }
        *out_n = n;
    });
}
BLK_API int blh_session_fill_ctx(void* i, const int32_t* toks, int n, const blk_token_data* claimed /*[n][10]*/, const int32_t* n_claimed,
                                 blk_token_data* out /*[n][10]*/, int32_t* out_n) {
    return guard([&] {
        std::vector<TokenPrediction> in(static_cast<size_t>(n));
        for (int t = 0; t < n; ++t) { in[size_t(t)].token = toks[t]; in[size_t(t)].logits = toVec(claimed + size_t(t) * 10, n_claimed[t]); }
        auto res = static_cast<InstanceBox*>(i)->session->fillCtx(in);
        for (size_t t = 0; t < res.size(); ++t) {
            out_n[t] = int32_t(res[t].logits.size());
            for (size_t j = 0; j < res[t].logits.size() && j < 10; ++j) out[t * 10 + j] = {res[t].logits[j].token, res[t].logits[j].logit};
        }
    });
}
// Server::verify in one call (reference Server.cpp:127-161 after tokenisation): fillCtx + LogitComparer over every position
BLK_API int blh_session_verify(void* i, const int32_t* toks, int n, const blk_token_data* claimed /*[n][10]*/, const int32_t* n_claimed, float* score) {
    return guard([&] {
        std::vector<TokenPrediction> orig(static_cast<size_t>(n));
        for (int t = 0; t < n; ++t) { orig[size_t(t)].token = toks[t]; orig[size_t(t)].logits = toVec(claimed + size_t(t) * 10, n_claimed[t]); }
        auto mine = static_cast<InstanceBox*>(i)->session->fillCtx(orig);
        // The reference pushes the metrics one by one and re-sums its whole history on every push (Server.cpp:153-156, O(n^2));
        // only the value of the LAST push is returned, and that is the in-order double sum over all metrics: one push of the
        // whole span computes exactly that float.
        std::vector<TokenPredictionView> pairs(orig.size());
        for (size_t t = 0; t < orig.size(); ++t) pairs[t] = {&orig[t].logits, &mine[t].logits};
        const std::vector<ComparisonMetrics> ms = compareAll(pairs);
        MetricsAggregator agg;
        *score = ms.empty() ? 0.0f : agg.pushAndVerify(ms);
    });
}
BLK_API int blh_session_get_state(void* i) { return guard([&] { (void)static_cast<InstanceBox*>(i)->session->getState(); }); }
BLK_API int blh_session_set_state(void* i) { return guard([&] { (void)static_cast<InstanceBox*>(i)->session->setState({}); }); }

// ---- verdict -------------------------------------------------------------------------------------------------------
BLK_API void blh_lc_compare(const blk_token_data* a, int32_t na, const blk_token_data* b, int32_t nb, float* out3) {
    auto m = LogitComparer::compare(toVec(a, na), toVec(b, nb));
    out3[0] = m.top1Match; out3[1] = m.distance; out3[2] = m.jsd;
}
BLK_API float blh_lc_similarity(const blk_token_data* a, int32_t na, const blk_token_data* b, int32_t nb) {
    return LogitComparer::logitSimilarity(toVec(a, na), toVec(b, nb));
}
// pushes metrics one at a time (as Server::verify does) and returns the final score
BLK_API float blh_lc_score(const float* metrics3, int32_t n) {
    MetricsAggregator agg; float score = 0;
    for (int32_t i = 0; i < n; ++i) { ComparisonMetrics m{metrics3[i * 3], metrics3[i * 3 + 1], metrics3[i * 3 + 2]}; score = agg.pushAndVerify({&m, 1}); }
    return score;
}

// ---- sampler (host only; no device needed) -------------------------------------------------------------------------
// draws n_draws tokens from the same sorted candidate list with one sampler instance
BLK_API int blh_sampler_draw(uint32_t seed, float temp, float top_p, int32_t top_k, float min_p, int32_t min_keep,
                             const blk_token_data* cand, int32_t n_cand, int32_t sorted, int32_t n_draws, int32_t* out) {
    return guard([&] {
        Sampler::Params sp; sp.rngSeed = seed; sp.temp = temp; sp.topP = top_p; sp.topK = top_k; sp.minP = min_p; sp.minKeep = min_keep;
        Model* none = nullptr;
        Sampler s(*none, sp);
        auto v = toVec(cand, n_cand);
        for (int32_t i = 0; i < n_draws; ++i) out[i] = s.sample(v, sorted != 0);
    });
}

} // extern "C"
