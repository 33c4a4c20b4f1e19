"""Generates tests/golden/json_wire_golden.json with nlohmann::json 3.x itself (the library the reference's HTTP layer uses,
server/code/http/HttpServerMain.cpp:23; a copy ships inside cudnn_frontend's headers in this image): for a set of floats the text
`nlohmann::json(float).dump()` produces, and for a /complete answer built exactly like the reference's toJson + getCompleteResponse
(HttpServerMain.cpp:37-51, 255-261) the dumped body.  Run in the build container only; the vectors are committed."""
import json, os, struct, subprocess, sys, tempfile
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
INC = None
for base in sys.path:
    p = os.path.join(base, "include", "cudnn_frontend", "thirdparty")
    if os.path.exists(os.path.join(p, "nlohmann", "json.hpp")):
        INC = p
assert INC, "nlohmann/json.hpp not found"

SRC = r'''
#include <nlohmann/json.hpp>
#include <cstdio>
#include <cstring>
#include <iostream>
#include <string>
int main() {
    // stdin: one hex float bit pattern per line until "--", then token records "id nlogits (id bits)*  str-hex"
    std::string line;
    nlohmann::json out;
    out["floats"] = nlohmann::json::array();
    while (std::getline(std::cin, line) && line != "--") {
        unsigned u = std::stoul(line, nullptr, 16); float f; memcpy(&f, &u, 4);
        nlohmann::json j = f;
        out["floats"].push_back({{"bits", line}, {"text", j.dump()}});
    }
    nlohmann::json tokens = nlohmann::json::array();
    std::string text;
    while (std::getline(std::cin, line)) {
        std::istringstream ss(line);
        unsigned id; int n; ss >> id >> n;
        auto& jt = tokens.emplace_back();
        auto& jl = jt["logits"] = nlohmann::json::array();
        for (int i = 0; i < n; i++) { unsigned lid; std::string bits; ss >> lid >> bits; unsigned u = std::stoul(bits, nullptr, 16); float f; memcpy(&f, &u, 4); auto& l = jl.emplace_back(); l["id"] = lid; l["logit"] = f; }
        std::string hex; ss >> hex; std::string s;
        for (size_t i = 0; i + 1 < hex.size(); i += 2) s += char(std::stoul(hex.substr(i, 2), nullptr, 16));
        if (hex == "-") s.clear();
        jt["str"] = s; jt["id"] = id; text += s;
    }
    nlohmann::json body; body["text"] = text; body["tokenData"] = tokens;
    out["complete_body"] = body.dump();
    nlohmann::json v({{"result", 0.9973522424697876f}});
    out["verify_body"] = v.dump();
    std::cout << out.dump() << std::endl;
}
'''

def main():
    rng = np.random.default_rng(20251018)
    floats = [0.0, -0.0, 1.0, -1.0, 17.5, 13.0, 0.1, 0.2, 1e-4, 1e-5, 9.999999e-5, 123456.789, 1e15, 1e16, 3.4028235e38, 1.17549435e-38, 1e-45,
              0.9973522424697876, 16777216.0, 16777217.0, 0.333333343, 2.5e-7, 1234567.0, 99999.99, 100000.0, 1e10, 5e14, 9.99e14, 1.5e15]
    floats += list(rng.normal(0, 3, 3000).astype(np.float32))
    floats += list((rng.normal(0, 1, 500) * 10.0 ** rng.integers(-12, 20, 500)).astype(np.float32))
    bits = [struct.pack("<f", np.float32(f)).hex() for f in floats]
    bits = [struct.unpack("<I", bytes.fromhex(b))[0] for b in bits]
    lines = ["%08x" % b for b in bits] + ["--"]
    toks = []
    strs = [" Bush", "\"quote\"\\", "tab\there", "nl\n", "été", "\U0001F600", "\x01ctl", "", "plain"]
    for i, s in enumerate(strs):
        n = int(rng.integers(0, 11))
        rec = [str(1000 + i), str(n)]
        for j in range(n):
            rec += [str(int(rng.integers(0, 128256))), "%08x" % struct.unpack("<I", struct.pack("<f", np.float32(rng.normal(0, 4))))[0]]
        rec.append(s.encode("utf-8").hex() or "-")
        toks.append(" ".join(rec))
    with tempfile.TemporaryDirectory() as td:
        src = os.path.join(td, "g.cpp"); exe = os.path.join(td, "g")
        open(src, "w").write("#include <sstream>\n" + SRC)
        subprocess.check_call(["g++", "-O1", "-std=c++17", "-I", INC, src, "-o", exe])
        out = subprocess.run([exe], input="\n".join(lines + toks) + "\n", capture_output=True, text=True, check=True).stdout
    data = json.loads(out)
    data["token_records"] = toks
    os.makedirs(os.path.join(ROOT, "tests", "golden"), exist_ok=True)
    with open(os.path.join(ROOT, "tests", "golden", "json_wire_golden.json"), "w") as f:
        json.dump(data, f)
    print("floats:", len(data["floats"]), "body bytes:", len(data["complete_body"]))

if __name__ == "__main__":
    main()
