#include "Init.hpp"
#include "Errors.hpp"

#include <blama_b200.h>

#include <cstdio>
#include <mutex>

namespace bl::llama {
namespace {
std::mutex g_sinkMutex;
std::function<void(LogLevel, const std::string&)> g_sink;

void engineLog(int level, const char* text, void*) {
    logLine(static_cast<LogLevel>(level < 0 ? 0 : (level > 3 ? 3 : level)), text ? text : "");
}
} // namespace

void setLogSink(std::function<void(LogLevel, const std::string&)> sink) {
    std::lock_guard<std::mutex> lock(g_sinkMutex);
    g_sink = std::move(sink);
}

void logLine(LogLevel level, const std::string& text) {
    std::function<void(LogLevel, const std::string&)> sink;
    {
        std::lock_guard<std::mutex> lock(g_sinkMutex);
        sink = g_sink;
    }
    if (sink) sink(level, text);
    else if (level >= LogLevel::Warning) std::fprintf(stderr, "[bl:llama] %s\n", text.c_str());
}

void initLibrary() {
    blk_set_log_callback(engineLog, nullptr);
    // no device is not fatal here: the reference's llama_backend_init cannot fail either; Model construction reports it
    (void)blk_init();
}

void throwIfFailed(int status, const char* what) {
    if (status != BLK_OK) Raise{} << what << ": " << blk_last_error();
}

} // namespace bl::llama
