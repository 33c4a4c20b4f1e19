// Model.cpp -- reference inference/code/llama/Model.cpp:50-87 over the C ABI (blk_model_load instead of
// llama_model_load_from_file).
#include "Model.hpp"
#include "Errors.hpp"

#include <blama_b200.h>

namespace bl::llama {
namespace {
int32_t progressTrampoline(float p, void* user) {
    auto* cb = static_cast<ModelLoadProgressCb*>(user);
    if (*cb) (*cb)(p);
    return 1;
}
blk_model* load(const std::string& gguf, const Model::Params& params, ModelLoadProgressCb& pcb) {
    if (params.vocabOnly) {      // metadata + vocabulary: no device involved (reference test "vocab only", t-integration.cpp:25-43)
        blk_model* v = blk_model_load_vocab(gguf.c_str());
        if (!v) Raise{} << "Failed to load model " << gguf << ": " << blk_last_error();
        return v;
    }
    if (!params.gpu) Raise{} << "blama_b200 has no CPU backend: Model::Params::gpu must be true";
    blk_model* m = blk_model_load(gguf.c_str(), params.device, pcb ? progressTrampoline : nullptr, &pcb);
    // the reference does not check for a null model (Model.cpp:50-53) and crashes later; fail here with the cause
    if (!m) Raise{} << "Failed to load model " << gguf << ": " << blk_last_error();
    return m;
}
} // namespace

Model::Model(const std::string& gguf, Params params, ModelLoadProgressCb pcb)
    : m_params(params), m_handle(load(gguf, params, pcb), blk_model_free) {}

Model::~Model() = default;

uint32_t Model::trainCtxLength() const noexcept { return uint32_t(blk_model_n_ctx_train(m_handle.get())); }   // 0 for vocabOnly
bool Model::shouldAddBosToken() const noexcept { return blk_model_add_bos(m_handle.get()) != 0; }

std::string Model::getChatTemplateId() const {
    char buf[2048];
    const int32_t len = blk_model_meta_str(m_handle.get(), "tokenizer.chat_template", buf, sizeof(buf));
    if (len < 0) return "chatml";
    return std::string(buf, size_t(len) < sizeof(buf) ? size_t(len) : sizeof(buf));
}

} // namespace bl::llama
