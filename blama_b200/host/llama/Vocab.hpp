// Vocab.hpp -- token <-> text helpers (mirror of reference inference/code/llama/Vocab.hpp:16-34).
#pragma once
#include "Token.hpp"

#include <string>
#include <string_view>
#include <unordered_map>
#include <vector>

struct blk_model;

namespace bl::llama {

class Model;

class Vocab {
public:
    explicit Vocab(const Model& model);
    ~Vocab();

    // Text -> ids.  First pass: greedy longest match over the vocabulary's token texts, which is
    // exact for the synthetic vocabularies ("<t123><t7>...") and for special tokens; the BPE merge pass of
    // llama-vocab.cpp is the next item.  addSpecial prepends BOS when the model asks for it (llama_tokenize semantics).
    std::vector<Token> tokenize(std::string_view text, bool addSpecial, bool parseSpecial) const;

    Token decoderStartToken() const noexcept;   // no encoder models here: BOS
    bool isEog(Token token) const noexcept;
    int32_t nTokens() const noexcept;
    std::string tokenToString(Token token, bool special = true) const;

    const blk_model* lvocab() const noexcept;   // the C handle the vocabulary lives in

private:
    void buildIndex() const;
    const Model& m_model;
    mutable bool m_indexed = false;
    mutable std::unordered_map<std::string, Token> m_byText;
    mutable size_t m_longest = 0;
};

} // namespace bl::llama
