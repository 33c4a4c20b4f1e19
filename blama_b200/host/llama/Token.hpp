// Token.hpp -- boundary POD types of the inference API (mirror of reference inference/code/llama/Token.hpp:9-17).
#pragma once
#include <cstdint>
#include <vector>

namespace bl::llama {

using Token = std::int32_t;
inline constexpr Token Token_Invalid = -1;

// {token id, raw logit}: 8 bytes, layout-compatible with blk_token_data of the C ABI
struct TokenData {
    Token token;
    float logit;
};
using TokenDataVector = std::vector<TokenData>;

} // namespace bl::llama
