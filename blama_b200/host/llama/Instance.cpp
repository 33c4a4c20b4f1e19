// Instance.cpp -- reference inference/code/llama/Instance.cpp:34-132 over the C ABI.
#include "Instance.hpp"
#include "Errors.hpp"
#include "Init.hpp"
#include "Model.hpp"

#include <blama_b200.h>

#include <vector>

namespace bl::llama {

Instance::Instance(Model& model, InitParams params)
    : m_model(model)
    , m_ctx(blk_ctx_create(model.lmodel(), int32_t(params.ctxSize), int32_t(params.batchSize)), blk_ctx_free) {
    if (!m_ctx) Raise{} << "Failed to create llama context";
    const auto ctxLen = uint32_t(blk_ctx_n_ctx(m_ctx.get()));
    if (ctxLen > model.trainCtxLength())
        logLine(LogLevel::Warning, "Instance requested context length " + std::to_string(ctxLen) +
                                       " is greater than the model's training context length " + std::to_string(model.trainCtxLength()));
}

Instance::~Instance() = default;

void Instance::warmup() {
    logLine(LogLevel::Info, "Running warmup");
    std::vector<Token> tmp;
    const Token bos = blk_model_token_bos(m_model.lmodel());
    const Token eos = blk_model_token_eos(m_model.lmodel());
    if (bos >= 0) tmp.push_back(bos);
    if (eos >= 0) tmp.push_back(eos);
    if (tmp.empty()) tmp.push_back(0);
    (void)blk_decode(m_ctx.get(), tmp.data(), int32_t(tmp.size()));
    (void)blk_kv_clear(m_ctx.get());
    (void)blk_sync(m_ctx.get());
}

Session& Instance::startSession(const Session::InitParams params) {
    if (m_session.has_value()) Raise{} << "Session is already started. Stop it to start a new one.";
    m_session.emplace(*this, m_ctx.get(), params);
    return *m_session;
}

void Instance::stopSession() noexcept { m_session.reset(); }

} // namespace bl::llama
