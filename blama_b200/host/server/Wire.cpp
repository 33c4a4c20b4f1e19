// Wire.cpp -- see Wire.hpp.  Field-by-field the reference's converters (HttpServerMain.cpp:37-94), on bl::json.
#include "Wire.hpp"
#include "Json.hpp"

namespace bl::llama::server::wire {
namespace {

// opt_get (HttpServerMain.cpp:72-83): present -> converted, absent -> default kept
template <class T> void optNum(const json::Value& dict, std::string_view key, T& value) {
    if (const json::Value* v = dict.find(key)) value = v->num<T>();
}
void optStr(const json::Value& dict, std::string_view key, std::string& value) {
    if (const json::Value* v = dict.find(key)) value = v->str();
}

Server::CompleteRequestParams paramsOf(const json::Value& j) {
    Server::CompleteRequestParams p;
    p.prompt = j["prompt"].str();                 // required: a missing / non-string prompt throws (:87)
    optNum(j, "max_tokens", p.maxTokens);
    optNum(j, "seed", p.seed);
    optStr(j, "suffix", p.suffix);
    optNum(j, "temp", p.temperature);
    optNum(j, "top_p", p.topP);
    return p;
}

Server::CompleteReponse responseOf(const json::Value& j) {
    Server::CompleteReponse gen;
    const json::Value& tokens = j["tokenData"];
    if (tokens.isNull()) return gen;              // iterating a null json visits nothing (:57)
    gen.reserve(tokens.size());
    for (const json::Value& jt : tokens.array()) {
        auto& g = gen.emplace_back();
        g.tokenStr = jt["str"].str();
        g.tokenId = uint32_t(jt["id"].num<int>());
        const json::Value& jl = jt["logits"];
        if (jl.isNull()) continue;
        g.logits.reserve(jl.size());
        for (const json::Value& l : jl.array()) {
            auto& out = g.logits.emplace_back();
            out.tokenId = uint32_t(l["id"].num<int>());
            out.logit = l["logit"].num<float>();
        }
    }
    return gen;
}

} // namespace

Server::CompleteRequestParams parseCompleteParams(std::string_view body) { return paramsOf(json::parse(body)); }

VerifyBody parseVerifyBody(std::string_view body) {
    const json::Value j = json::parse(body);
    return {paramsOf(j["request"]), responseOf(j["response"])};
}

std::string completeResponseJson(const Server::CompleteReponse& response) {
    std::string text;
    json::Array tokens;
    tokens.reserve(response.size());
    for (const auto& g : response) {
        text += g.tokenStr;
        json::Object jt;
        jt["str"] = g.tokenStr;
        jt["id"] = g.tokenId;
        json::Array jl;
        jl.reserve(g.logits.size());
        for (const auto& l : g.logits) {
            json::Object o;
            o["id"] = l.tokenId;
            o["logit"] = l.logit;
            jl.emplace_back(std::move(o));
        }
        jt["logits"] = std::move(jl);
        tokens.emplace_back(std::move(jt));
    }
    json::Object out;
    out["text"] = std::move(text);
    out["tokenData"] = std::move(tokens);
    return json::Value(std::move(out)).dump();
}

std::string verifyResponseJson(float score) {
    json::Object out;
    out["result"] = score;
    return json::Value(std::move(out)).dump();
}

} // namespace bl::llama::server::wire
