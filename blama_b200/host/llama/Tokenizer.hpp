// Tokenizer.hpp -- byte-level BPE tokenizer / detokenizer behind Vocab::tokenize and Vocab::tokenToString
// (reference inference/code/llama/Vocab.cpp:37-72 -> llama_tokenize / llama_token_to_piece; the algorithm lives in
// llama.cpp b5187 src/llama-vocab.cpp + src/unicode.cpp, un-vendored, restated here from its published behaviour):
//
//   1. special-token partition: with parseSpecial every CONTROL / USER_DEFINED / UNKNOWN token text found in the input becomes
//      that token (longest texts first); without it only USER_DEFINED ones are matched, the rest is ordinary text;
//   2. pre-tokenizer: the text between special tokens is split by the pattern the file names in tokenizer.ggml.pre
//        llama-bpe / llama3 / llama-v3   (?i:'s|'t|'re|'ve|'m|'ll|'d)|[^\r\n\p{L}\p{N}]?\p{L}+|\p{N}{1,3}| ?[^\s\p{L}\p{N}]+[\r\n]*|\s*[\r\n]+|\s+(?!\S)|\s+
//        qwen2                           the same with \p{N} one digit at a time
//        gpt-2 / default                 's|'t|'re|'ve|'m|'ll|'d| ?\p{L}+| ?\p{N}+| ?[^\s\p{L}\p{N}]+|\s+(?!\S)|\s+
//      written out as a hand-rolled matcher (no regex engine), with the leftmost-first / greedy / backtracking meaning of the patterns;
//   3. every word is mapped byte -> printable code point (GPT-2's bytes_to_unicode) and merged bottom-up: always the adjacent pair
//      with the lowest merge rank, leftmost first; llama-bpe files skip the merges for a word that is a vocabulary entry as a whole
//      (ignore_merges);
//   4. pieces missing from the vocabulary fall back to their single bytes.
// Detokenisation maps the code points of a NORMAL token back to bytes; CONTROL / USER_DEFINED tokens print their text verbatim
// (CONTROL ones only when `special` is set).
#pragma once
#include "Token.hpp"

#include <string>
#include <string_view>
#include <unordered_map>
#include <vector>

namespace bl::llama {

class BpeTokenizer {
public:
    enum class Pre { Gpt2, Llama3, Qwen2 };
    struct Config {
        std::vector<std::string> tokens;     // tokenizer.ggml.tokens
        std::vector<int32_t> types;          // tokenizer.ggml.token_type (1 normal, 2 unknown, 3 control, 4 user defined, 5 unused, 6 byte)
        std::vector<std::string> merges;     // tokenizer.ggml.merges, rank order
        std::string pre;                     // tokenizer.ggml.pre
        Token bos = Token_Invalid, eos = Token_Invalid;
        bool addBos = false, addEos = false;
    };
    explicit BpeTokenizer(Config config);

    std::vector<Token> tokenize(std::string_view text, bool addSpecial, bool parseSpecial) const;
    std::string tokenToPiece(Token token, bool special) const;

    // the pre-tokenizer alone (byte offsets of the words); exposed for the tests
    std::vector<std::string_view> split(std::string_view text) const;

private:
    void bpeWord(const std::string& word, std::vector<Token>& out) const;

    Config m_cfg;
    Pre m_pre;
    bool m_ignoreMerges;
    std::unordered_map<std::string, Token> m_byText;
    std::unordered_map<std::string, int32_t> m_rank;      // "left right" -> rank
    std::vector<Token> m_specials;                         // longest text first
};

} // namespace bl::llama
