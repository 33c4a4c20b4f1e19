// ref_lc_wrapper.cpp -- C wrapper around the REFERENCE's own LogitComparer (compiled from
// /root/reference/inference/code/llama/LogitComparer.cpp where it lies; never copied into this repo).
// Built only where /root/reference exists (oracle/Makefile target _ref); used to pin the oracle's restatement and
// to generate tests/golden/logit_comparer_golden.json (tools/gen_logit_comparer_golden.py).
#include <llama/LogitComparer.hpp>
#include <cstdint>

using namespace bl::llama;

extern "C" {
struct ref_td { int32_t token; float logit; };
struct ref_metrics { float top1Match, distance, jsd; };

static TokenDataVector to_vec(const ref_td* a, int32_t n) {
    TokenDataVector v; v.reserve(n);
    for (int32_t i = 0; i < n; i++) v.push_back({a[i].token, a[i].logit});
    return v;
}
ref_metrics ref_lc_compare(const ref_td* a, int32_t na, const ref_td* b, int32_t nb) {
    auto m = LogitComparer::compare(to_vec(a, na), to_vec(b, nb));
    return {m.top1Match, m.distance, m.jsd};
}
float ref_lc_similarity(const ref_td* a, int32_t na, const ref_td* b, int32_t nb) {
    return LogitComparer::logitSimilarity(to_vec(a, na), to_vec(b, nb));
}
// score after pushing metrics one at a time, as Server::verify does (Server.cpp:151-156)
float ref_lc_score(const ref_metrics* m, int32_t n) {
    MetricsAggregator agg; float score = 0;
    for (int32_t i = 0; i < n; i++) { ComparisonMetrics cm{m[i].top1Match, m[i].distance, m[i].jsd}; score = agg.pushAndVerify({&cm, 1}); }
    return score;
}
}
