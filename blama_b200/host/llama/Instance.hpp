// Instance.hpp -- owns one device context: paged KV cache + workspaces + CUDA graphs
// (mirror of reference inference/code/llama/Instance.hpp:18-52).
#pragma once
#include "Session.hpp"

#include <memory>
#include <optional>

struct blk_ctx;

namespace bl::llama {

class Model;

class Instance {
public:
    struct InitParams {
        uint32_t ctxSize = 0;       // 0 = the model's training context
        uint32_t batchSize = 2048;  // logical prefill batch
        uint32_t ubatchSize = 512;  // kept for source compatibility; the prefill GEMM picks its own tiles
        bool flashAttn = false;     // kept for source compatibility; attention is always fused here
    };

    explicit Instance(Model& model, InitParams params);
    ~Instance();

    void warmup();                                  // one throw-away decode so first-request latency is flat
    Session& startSession(const Session::InitParams params);   // only one live session per instance
    void stopSession() noexcept;
    Model& model() const noexcept { return m_model; }
    blk_ctx* lctx() noexcept { return m_ctx.get(); }

private:
    Model& m_model;
    std::unique_ptr<blk_ctx, void (*)(blk_ctx*)> m_ctx;
    std::optional<Session> m_session;
};

} // namespace bl::llama
