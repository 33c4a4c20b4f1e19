// Model.hpp -- owns the device-resident weights (mirror of reference inference/code/llama/Model.hpp:26-58).
#pragma once
#include "Vocab.hpp"

#include <functional>
#include <memory>
#include <string>

struct blk_model;

namespace bl::llama {

using ModelLoadProgressCb = std::function<void(float)>;

class Model {
public:
    struct Params {
        bool gpu = true;                  // the reference falls back to the CPU backend when false; this build has no
                                          // CPU backend: gpu = false is rejected
        bool vocabOnly = false;           // do not load the weights (tokenizer use only; needs no GPU)
        bool prefixInputsWithBos = false; // add bos token to interactive inputs (#13)
        int device = 0;                   // extension: which GPU holds this replica (one replica per GPU)
        bool operator==(const Params& other) const noexcept = default;
    };

    Model(const std::string& gguf, Params params, ModelLoadProgressCb pcb = {});
    ~Model();
    Model(const Model&) = delete;
    Model& operator=(const Model&) = delete;

    const Params& params() const noexcept { return m_params; }
    uint32_t trainCtxLength() const noexcept;
    bool shouldAddBosToken() const noexcept;
    bool hasEncoder() const noexcept { return false; }
    bool prefixInputsWithBos() const noexcept { return m_params.prefixInputsWithBos; }
    std::string getChatTemplateId() const;   // "chatml" when the GGUF carries no template

    blk_model* lmodel() noexcept { return m_handle.get(); }
    const blk_model* lmodel() const noexcept { return m_handle.get(); }
    const Vocab& vocab() const noexcept { return m_vocab; }

private:
    const Params m_params;
    std::unique_ptr<blk_model, void (*)(blk_model*)> m_handle;
    Vocab m_vocab{*this};
};

} // namespace bl::llama
