// HttpServerMain.cpp -- the server executable (reference server/code/http/HttpServerMain.cpp:372-448): environment configuration,
// model load, Server + HTTP front end.
//
//   BLAMA_MODEL  path of the .gguf file (must end in .gguf, exist and be a regular file; reference :412-435).  The reference falls
//                back to its GPT-2 test fixture; this build has no bundled model, so the variable is required.
//   BLAMA_HOST   IPv4 address to bind, default 0.0.0.0 (:379, :383-394)
//   BLAMA_PORT   port, default 7331; trailing characters and values above 65535 are rejected with the reference's texts (:396-410)
//   BLAMA_GPUS   extension: number of GPU replicas (one model copy, Instance and worker thread per GPU), default 1
//   BLAMA_CTX    extension: context length per replica, default 0 = the model's training context (reference Instance.hpp:22)
#include "../../llama/Init.hpp"
#include "../../llama/Model.hpp"
#include "../Http.hpp"
#include "../Server.hpp"

#include <chrono>
#include <csignal>
#include <cstdlib>
#include <cstring>
#include <filesystem>
#include <iostream>
#include <limits>
#include <stdexcept>
#include <thread>

namespace fs = std::filesystem;
using bl::llama::Model;

namespace {
volatile std::sig_atomic_t g_stop = 0;
void onSignal(int) { g_stop = 1; }

unsigned long envNumber(const char* name, const char* text) {
    size_t idx = 0;
    const unsigned long value = std::stoul(text, &idx, 10);
    if (idx != std::strlen(text)) throw std::invalid_argument(std::string("Extra characters after ") + name + " number");
    return value;
}
} // namespace

int main() {
    try {
        bl::llama::initLibrary();

        std::string host = "0.0.0.0";
        uint16_t port = 7331;
        if (const char* h = std::getenv("BLAMA_HOST")) host = h;
        if (const char* p = std::getenv("BLAMA_PORT")) {
            const unsigned long value = envNumber("BLAMA_PORT", p);
            if (value > std::numeric_limits<uint16_t>::max()) throw std::out_of_range("Value exceeds uint16_t max");
            port = static_cast<uint16_t>(value);
        }
        const char* modelEnv = std::getenv("BLAMA_MODEL");
        if (!modelEnv || std::string(modelEnv).empty()) throw std::runtime_error("Environment variable not set or empty: BLAMA_MODEL");
        const std::string modelPath(modelEnv);
        if (!modelPath.ends_with(".gguf")) throw std::runtime_error("BLAMA_MODEL does not end with .gguf: " + modelPath);
        if (!fs::exists(modelPath)) throw std::runtime_error("BLAMA_MODEL does not exist: " + modelPath);
        if (!fs::is_regular_file(modelPath)) throw std::runtime_error("BLAMA_MODEL is not a regular file: " + modelPath);
        unsigned long gpus = 1, ctx = 0;
        if (const char* g = std::getenv("BLAMA_GPUS")) gpus = envNumber("BLAMA_GPUS", g);
        if (const char* c = std::getenv("BLAMA_CTX")) ctx = envNumber("BLAMA_CTX", c);
        if (gpus == 0 || gpus > 64) throw std::out_of_range("BLAMA_GPUS must be between 1 and 64");

        std::cerr << "Loading model " << modelPath << " on " << gpus << " GPU(s)\n";
        std::vector<std::shared_ptr<Model>> replicas;
        for (unsigned long g = 0; g < gpus; ++g) {
            Model::Params mp;
            mp.device = int(g);
            replicas.push_back(std::make_shared<Model>(modelPath, mp, [g](float progress) {
                const int pct = int(progress * 100);
                if (pct % 10 == 0) std::cerr << "\rLoading model (GPU " << g << "): " << pct << "%" << std::flush;      // reference :353-370 prints the progress
            }));
            std::cerr << "\n";
        }
        bl::llama::Instance::InitParams ip;
        ip.ctxSize = uint32_t(ctx);
        bl::llama::server::Server server(std::move(replicas), ip);
        server.setErrorHandler([](const std::string& e) { std::cerr << "request failed: " << e << "\n"; });
        bl::llama::server::HttpFrontEnd http(server, host, port, 4);
        std::cerr << "Listening on port " << http.port() << "\n";

        std::signal(SIGINT, onSignal);
        std::signal(SIGTERM, onSignal);
        while (!g_stop) std::this_thread::sleep_for(std::chrono::milliseconds(100));
        std::cerr << "shutting down after " << http.requestsServed() << " requests\n";
        return 0;
    } catch (const std::exception& e) {
        std::cerr << "blama-server: " << e.what() << "\n";
        return 1;
    }
}
