// Session.hpp -- per-request state machine (mirror of reference inference/code/llama/Session.hpp:29-137).
//
// Same public surface and error strings as the reference; underneath, every llama.cpp call is replaced by the C ABI:
//   llama_decode(1 token) + llama_get_logits_ith + host sort   ->  blk_decode_topk (one graph launch, 512 B back)
//   fillCtx's N single-token decodes + V-long host gather     ->  blk_verify_prefill (one causal prefill)
#pragma once
#include "Sampler.hpp"
#include "Token.hpp"

#include <memory>
#include <span>
#include <string>
#include <vector>

struct blk_ctx;

namespace bl::llama {

class Instance;

struct TokenPrediction {
    Token token;
    TokenDataVector logits;      // top-10 {id, logit} of the distribution AFTER `token` was decoded
    operator bool() const { return token != Token_Invalid; }
};

class Session {
public:
    struct InitParams {
        uint32_t gaFactor = 1;          // group-attention (Self-Extend) factor; only 1 is supported by this build
        uint32_t gaWidth = 512;
        bool infiniteContext = true;    // context shifting is not supported by this build: a full context throws
        uint32_t seed = 0;
        std::string grammar;
        float temperature = 0.80f;
        float topP = 0.95f;
        // extension: false = fillCtx runs one causal prefill (fast, fp tolerance); true = N single-token decodes,
        // bit-identical to complete() (what the reference does, t-integration.cpp:219-248)
        bool sequentialVerify = false;
    };

    Session(Instance& instance, blk_ctx* ctx, InitParams params);
    Session(const Session&) = delete;
    Session& operator=(const Session&) = delete;
    ~Session();

    void setInitialPrompt(std::span<const Token> prompt);
    bool setState(std::span<uint8_t> state);

    struct CompleteParams {
        std::span<const Token> prompt;
        std::span<const Token> suffix;
        int32_t maxTokens = 0;
    };
    std::vector<TokenPrediction> complete(CompleteParams params);

    class StreamGenerator {
    public:
        enum class Status { InProgress, Completed, Aborted };
        StreamGenerator(Session& session, CompleteParams params)
            : m_session(session), m_params(params), m_genTokens(0), m_status(Status::InProgress) {}
        TokenPrediction complete();
        void abort();
        Status status() const { return m_status; }
    private:
        Session& m_session;
        CompleteParams m_params;
        int32_t m_genTokens;
        Status m_status;
    };
    StreamGenerator completeStream(CompleteParams params);

    std::vector<TokenPrediction> fillCtx(std::span<TokenPrediction> tokens);
    std::vector<uint8_t> getState();
    void resetSampler(const Sampler::Params& params);

private:
    friend class StreamGenerator;
    enum class Source { InitialPrompt, InteractivePrompt, Generated };
    enum class Phase { Initial, Generating, Streaming };

    void pushPrompt(std::span<const Token> prompt, std::span<const Token> postfix = {});
    TokenPrediction getToken();
    void doDecode(std::span<const Token> tokens, Source src);
    void flushPendingState();
    TokenDataVector getLogitsFromCtx(int32_t topK);
    TokenDataVector getLogitsFromCtx(TokenDataVector tokens);
    void requireStarted(bool allowStreaming) const;
    void refreshCandidates();

    Instance& m_instance;
    blk_ctx* m_ctx;
    std::unique_ptr<Sampler> m_sampler;
    InitParams m_params;

    Phase m_phase = Phase::Initial;
    Token m_currToken = Token_Invalid;      // sampled but not yet decoded
    unsigned m_maxTokens = 0;
    unsigned m_numKeep = 0;
    uint32_t m_numPast = 0;
    TokenDataVector m_candidates;           // device top-k of the current logits, descending
};

} // namespace bl::llama
