// LogitComparer.cpp -- the verdict half of /verify_completion (reference inference/code/llama/LogitComparer.cpp).
//
// Bit-exactness note: the reference adds up fp32 terms while ITERATING std::unordered_map (softmax normalisation,
// KL terms), so libstdc++'s bucket order is part of its arithmetic.  To return the very same floats this
// implementation keeps the same container type, the same bucket-count hint (list length) and the same insertion
// sequence; everything else is organised differently (one table type, lookups instead of count()+at()).
// Checked bit-for-bit against the reference's own object code in tests/test_logit_comparer.py.
#include "LogitComparer.hpp"

#include <algorithm>
#include <thread>

#include <algorithm>
#include <cmath>
#include <unordered_map>

namespace bl::llama {
namespace {

using Table = std::unordered_map<Token, float>;

// softmax over a (descending) logit list; the FIRST entry is used as the stabilising maximum (reference :12)
Table probabilities(const TokenDataVector& list) {
    Table table(list.size());
    const float pivot = list[0].logit;
    float norm = 0.0f;
    for (const TokenData& td : list) {
        const float e = std::exp(td.logit - pivot);
        table[td.token] = e;       // duplicate ids overwrite, yet still count towards norm (reference behaviour)
        norm += e;
    }
    for (auto& entry : table) entry.second /= norm;
    return table;
}

float klTerms(const Table& p, const Table& q) {
    float acc = 0.0f;
    for (const auto& entry : p) {
        if (!(entry.second > 0.0f)) continue;
        const auto other = q.find(entry.first);
        if (other == q.end() || !(other->second > 0.0f)) continue;
        acc += entry.second * std::log(entry.second / other->second);
    }
    return acc;
}

float jensenShannon(const Table& p1, const Table& p2) {
    Table mid;                        // only ids present on both sides (reference :83-88)
    for (const auto& entry : p1) {
        const auto other = p2.find(entry.first);
        if (other != p2.end()) mid[entry.first] = (entry.second + other->second) / 2.0f;
    }
    return (klTerms(p1, mid) + klTerms(p2, mid)) / 2.0f;
}

float sumOfSquares(const TokenData* data, size_t n) {
    float s = 0.0f;
    for (size_t i = 0; i < n; ++i) s += data[i].logit * data[i].logit;
    return s;
}

} // namespace

ComparisonMetrics LogitComparer::compare(const TokenDataVector& data1, const TokenDataVector& data2) {
    ComparisonMetrics out{};
    out.top1Match = (data1[0].token == data2[0].token) ? 1.0f : 0.0f;
    const size_t common = std::min(data1.size(), data2.size());
    const float e1 = sumOfSquares(data1.data(), common);
    const float e2 = sumOfSquares(data2.data(), common);
    out.distance = std::fabs(e1 - e2) / std::max(e1, e2);
    out.jsd = jensenShannon(probabilities(data1), probabilities(data2));
    return out;
}

float LogitComparer::logitSimilarity(const TokenDataVector& data1, const TokenDataVector& data2) {
    Table theirs;
    for (const TokenData& td : data2) theirs[td.token] = td.logit;
    float weighted = 0.0f, weights = 0.0f;
    for (const TokenData& td : data1) {
        const float w = std::abs(td.logit);
        float sim = 0.0f;
        const auto hit = theirs.find(td.token);
        if (hit != theirs.end()) sim = 1 - (std::abs(td.logit - hit->second) / std::abs(std::max(td.logit, hit->second)));
        weighted += w * sim;
        weights += w;
    }
    return weights > 0.0f ? weighted / weights : 0.0f;
}

ComparisonMetrics compareChecked(const TokenDataVector& data1, const TokenDataVector& data2) {
    // LogitComparer::compare reads element 0 of both lists (reference :41): an empty list -- a response position without claimed
    // logits, or one whose claimed ids are all outside the vocabulary -- is undefined behaviour there.  Such a position cannot be
    // verified, so it scores as a complete mismatch instead of reading out of bounds.
    if (data1.empty() || data2.empty()) return ComparisonMetrics{0.0f, 1.0f, 1.0f};
    return LogitComparer::compare(data1, data2);
}

std::vector<ComparisonMetrics> compareAll(std::span<const TokenPredictionView> pairs) {
    std::vector<ComparisonMetrics> out(pairs.size());
    const size_t n = pairs.size();
    const unsigned hw = std::max(1u, std::thread::hardware_concurrency());
    const size_t n_thr = n >= 256 ? std::min<size_t>({size_t(hw), size_t(8), n / 128}) : 1;
    auto work = [&](size_t lo, size_t hi) { for (size_t i = lo; i < hi; ++i) out[i] = compareChecked(*pairs[i].a, *pairs[i].b); };
    if (n_thr <= 1) { work(0, n); return out; }
    std::vector<std::thread> thr;
    const size_t per = (n + n_thr - 1) / n_thr;
    for (size_t t = 1; t < n_thr; ++t) thr.emplace_back(work, std::min(n, t * per), std::min(n, (t + 1) * per));
    work(0, std::min(n, per));
    for (auto& t : thr) t.join();
    return out;
}

float MetricsAggregator::pushAndVerify(std::span<const ComparisonMetrics> m) {
    m_history.insert(m_history.end(), m.begin(), m.end());
    // the reference re-sums its whole history in double on every push (O(n^2) per request, LogitComparer.cpp:117-128);
    // kept as is: the summation order defines the returned float
    double total = 0.0;
    for (const ComparisonMetrics& h : m_history) total += 0.5 * (1.0f - h.distance) + 0.5 * (1.0f - h.jsd);
    return float(total / m_history.size());
}

} // namespace bl::llama
