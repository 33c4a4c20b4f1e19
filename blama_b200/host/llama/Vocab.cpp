// Vocab.cpp -- reference inference/code/llama/Vocab.cpp:13-72 over the C ABI.
#include "Vocab.hpp"
#include "Errors.hpp"
#include "Model.hpp"
#include "Tokenizer.hpp"

#include <blama_b200.h>

namespace bl::llama {
namespace {
// a string-valued entry point called with the grow-and-retry protocol of the C ABI
template <class F> std::string fetch(F&& call) {
    std::string out(64, '\0');
    int32_t len = call(out.data(), int32_t(out.size()));
    if (len > int32_t(out.size())) {
        out.resize(size_t(len));
        len = call(out.data(), int32_t(out.size()));
    }
    out.resize(size_t(len < 0 ? 0 : len));
    return out;
}
} // namespace

Vocab::Vocab(const Model& model) : m_model(model) {}
Vocab::~Vocab() = default;

const blk_model* Vocab::lvocab() const noexcept { return m_model.lmodel(); }
Token Vocab::decoderStartToken() const noexcept { return blk_model_token_bos(m_model.lmodel()); }
bool Vocab::isEog(Token token) const noexcept { return blk_model_is_eog(m_model.lmodel(), token) != 0; }
int32_t Vocab::nTokens() const noexcept { return blk_model_n_vocab(m_model.lmodel()); }

const BpeTokenizer& Vocab::tokenizer() const {
    std::call_once(m_once, [&] {
        const blk_model* m = m_model.lmodel();
        const std::string kind = fetch([&](char* b, int32_t c) { return blk_model_meta_str(m, "tokenizer.ggml.model", b, c); });
        if (!kind.empty() && kind != "gpt2") Raise{} << "unsupported tokenizer model '" << kind << "' (only byte-level BPE, tokenizer.ggml.model = gpt2)";
        BpeTokenizer::Config cfg;
        const int32_t n = blk_model_n_vocab(m);
        cfg.tokens.resize(size_t(n));
        cfg.types.resize(size_t(n));
        for (Token t = 0; t < n; ++t) {
            cfg.tokens[size_t(t)] = fetch([&](char* b, int32_t c) { return blk_model_token_text(m, t, b, c); });
            cfg.types[size_t(t)] = blk_model_token_type(m, t);
        }
        const int32_t nm = blk_model_n_merges(m);
        cfg.merges.resize(size_t(nm));
        for (int32_t r = 0; r < nm; ++r) cfg.merges[size_t(r)] = fetch([&](char* b, int32_t c) { return blk_model_merge_text(m, r, b, c); });
        cfg.pre = fetch([&](char* b, int32_t c) { return blk_model_meta_str(m, "tokenizer.ggml.pre", b, c); });
        cfg.bos = blk_model_token_bos(m);
        cfg.eos = blk_model_token_eos(m);
        cfg.addBos = blk_model_add_bos(m) != 0;
        cfg.addEos = blk_model_add_eos(m) != 0;
        m_tokenizer = std::make_unique<BpeTokenizer>(std::move(cfg));
    });
    return *m_tokenizer;
}

std::vector<Token> Vocab::tokenize(std::string_view text, bool addSpecial, bool parseSpecial) const {
    return tokenizer().tokenize(text, addSpecial, parseSpecial);
}

std::string Vocab::tokenToString(Token token, bool special) const { return tokenizer().tokenToPiece(token, special); }

} // namespace bl::llama
