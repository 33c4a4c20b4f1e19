// Wire.hpp -- the JSON bodies of POST /complete and POST /verify_completion
// (reference server/code/http/HttpServerMain.cpp:37-51 toJson, :53-70 toCompleteResponse, :85-94 toCompleteParams,
//  :255-275 getCompleteResponse, :277-288 getVerifyResponse).
//
//   request            {"prompt": str, "max_tokens"?: uint, "seed"?: uint, "suffix"?: str, "temp"?: float, "top_p"?: float}
//   /complete answer   {"text": concat(str), "tokenData": [{"id": uint, "logits": [{"id": uint, "logit": float} x <=10], "str": str}]}
//   /verify body       {"request": <request>, "response": </complete answer>}
//   /verify answer     {"result": float}
// Keys come out in byte order (nlohmann's std::map), floats as the shortest decimal of the widened double (Json.hpp).
#pragma once
#include "Server.hpp"

#include <string>
#include <string_view>

namespace bl::llama::server::wire {

// throws bl::json::ParseError / TypeError with nlohmann-like messages on malformed bodies
Server::CompleteRequestParams parseCompleteParams(std::string_view body);
struct VerifyBody { Server::CompleteRequestParams request; Server::CompleteReponse response; };
VerifyBody parseVerifyBody(std::string_view body);

std::string completeResponseJson(const Server::CompleteReponse& response);
std::string verifyResponseJson(float score);

} // namespace bl::llama::server::wire
