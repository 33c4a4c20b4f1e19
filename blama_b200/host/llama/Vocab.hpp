// Vocab.hpp -- token <-> text helpers (mirror of reference inference/code/llama/Vocab.hpp:16-34).
#pragma once
#include "Token.hpp"

#include <memory>
#include <mutex>
#include <string>
#include <string_view>
#include <vector>

struct blk_model;

namespace bl::llama {

class Model;
class BpeTokenizer;

class Vocab {
public:
    explicit Vocab(const Model& model);
    ~Vocab();

    // Text -> ids (llama_tokenize, reference Vocab.cpp:37-51): byte-level BPE with the pre-tokenizer the GGUF names
    // (Tokenizer.hpp).  addSpecial prepends BOS / appends EOS when the model's metadata asks for it; parseSpecial turns the
    // texts of control tokens ("<|eot_id|>") into those tokens instead of spelling them out.
    std::vector<Token> tokenize(std::string_view text, bool addSpecial, bool parseSpecial) const;

    Token decoderStartToken() const noexcept;   // no encoder models here: BOS
    bool isEog(Token token) const noexcept;
    int32_t nTokens() const noexcept;
    // llama_token_to_piece (reference Vocab.cpp:53-72): the bytes a token stands for; control tokens only when `special`
    std::string tokenToString(Token token, bool special = true) const;

    const blk_model* lvocab() const noexcept;   // the C handle the vocabulary lives in

private:
    const BpeTokenizer& tokenizer() const;      // built from the model's vocabulary on first use
    const Model& m_model;
    mutable std::once_flag m_once;
    mutable std::unique_ptr<BpeTokenizer> m_tokenizer;
};

} // namespace bl::llama
