// Token.hpp -- boundary value types of the inference API.  Same names and meaning as the reference's
// inference/code/llama/Token.hpp:9-17 (callers of bl::llama keep compiling); here they are additionally pinned to the
// C ABI: a TokenData array is handed to blk_topk_last / blk_decode_topk as a blk_token_data array without conversion.
#pragma once
#include <cstddef>
#include <cstdint>
#include <type_traits>
#include <vector>

#include "blama_b200.h"

namespace bl::llama {

using Token = std::int32_t;                            // index into the model's vocabulary
inline constexpr Token Token_Invalid = Token(-1);      // "no token"

struct TokenData {                                     // one entry of a top-k / claimed-logits list
    Token token;                                       //   vocabulary id
    float logit;                                       //   raw (pre-softmax) logit of that id
};
using TokenDataVector = std::vector<TokenData>;

static_assert(std::is_trivially_copyable_v<TokenData> && std::is_standard_layout_v<TokenData>);
static_assert(sizeof(TokenData) == sizeof(blk_token_data) && alignof(TokenData) == alignof(blk_token_data));
static_assert(offsetof(TokenData, token) == offsetof(blk_token_data, token) && offsetof(TokenData, logit) == offsetof(blk_token_data, logit));

} // namespace bl::llama
