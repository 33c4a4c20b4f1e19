// Vocab.cpp -- reference inference/code/llama/Vocab.cpp:13-72 over the C ABI.
#include "Vocab.hpp"
#include "Model.hpp"

#include <blama_b200.h>

namespace bl::llama {

Vocab::Vocab(const Model& model) : m_model(model) {}
Vocab::~Vocab() = default;

const blk_model* Vocab::lvocab() const noexcept { return m_model.lmodel(); }
Token Vocab::decoderStartToken() const noexcept { return blk_model_token_bos(m_model.lmodel()); }
bool Vocab::isEog(Token token) const noexcept { return blk_model_is_eog(m_model.lmodel(), token) != 0; }
int32_t Vocab::nTokens() const noexcept { return blk_model_n_vocab(m_model.lmodel()); }

std::string Vocab::tokenToString(Token token, bool /*special*/) const {
    std::string out(32, '\0');
    int32_t len = blk_model_token_text(m_model.lmodel(), token, out.data(), int32_t(out.size()));
    if (len > int32_t(out.size())) {
        out.resize(size_t(len));
        len = blk_model_token_text(m_model.lmodel(), token, out.data(), int32_t(out.size()));
    }
    out.resize(size_t(len < 0 ? 0 : len));
    return out;
}

void Vocab::buildIndex() const {
    if (m_indexed) return;
    const int32_t n = nTokens();
    m_byText.reserve(size_t(n));
    for (Token t = 0; t < n; ++t) {
        std::string s = tokenToString(t);
        if (s.empty()) continue;
        m_longest = std::max(m_longest, s.size());
        m_byText.emplace(std::move(s), t);      // first id wins for duplicate texts
    }
    m_indexed = true;
}

std::vector<Token> Vocab::tokenize(std::string_view text, bool addSpecial, bool /*parseSpecial*/) const {
    buildIndex();
    std::vector<Token> out;
    if (addSpecial && m_model.shouldAddBosToken()) out.push_back(blk_model_token_bos(m_model.lmodel()));
    size_t pos = 0;
    std::string probe;
    while (pos < text.size()) {
        size_t len = std::min(m_longest, text.size() - pos);
        Token found = Token_Invalid;
        for (; len > 0; --len) {
            probe.assign(text.substr(pos, len));
            const auto it = m_byText.find(probe);
            if (it != m_byText.end()) { found = it->second; break; }
        }
        if (found == Token_Invalid) { ++pos; continue; }   // bytes with no vocabulary entry are skipped
        out.push_back(found);
        pos += len;
    }
    return out;
}

} // namespace bl::llama
