// LogitComparer.hpp -- prover-vs-verifier metric (mirror of reference inference/code/llama/LogitComparer.hpp:12-34).
#pragma once
#include "Token.hpp"

#include <span>
#include <vector>

namespace bl::llama {

struct ComparisonMetrics {
    float top1Match;   // 1 if both lists start with the same token
    float distance;    // |sum a^2 - sum b^2| / max(...)  over the common prefix length
    float jsd;         // Jensen-Shannon divergence of the two softmaxed lists over shared ids
};

class LogitComparer {
public:
    static ComparisonMetrics compare(const TokenDataVector& data1, const TokenDataVector& data2);
    static float logitSimilarity(const TokenDataVector& data1, const TokenDataVector& data2);
};

// LogitComparer::compare for every position of a verified response (prover's list vs verifier's list).  The positions are
// independent, so long responses are compared on several host threads; each metric is the same float either way.
// compare() for lists that come from the wire: an empty side gives {top1Match 0, distance 1, jsd 1} instead of the reference's
// out-of-bounds read
ComparisonMetrics compareChecked(const TokenDataVector& data1, const TokenDataVector& data2);
struct TokenPredictionView { const TokenDataVector* a; const TokenDataVector* b; };
std::vector<ComparisonMetrics> compareAll(std::span<const TokenPredictionView> pairs);

struct MetricsAggregator {
    // appends m and returns the mean of 0.5(1-distance) + 0.5(1-jsd) over everything pushed so far
    float pushAndVerify(std::span<const ComparisonMetrics> m);
private:
    std::vector<ComparisonMetrics> m_history;
};

} // namespace bl::llama
