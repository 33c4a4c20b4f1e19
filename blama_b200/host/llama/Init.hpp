// Init.hpp -- library initialisation (reference inference/code/llama/Init.hpp, Init.cpp:34-38).
#pragma once
#include <functional>
#include <string>

namespace bl::llama {

// must be called once before anything else (reference HttpServerMain.cpp:376)
void initLibrary();

enum class LogLevel { Debug = 0, Info = 1, Warning = 2, Error = 3 };
// route engine + host log lines to a sink (reference routes ggml logs to jalog scope "bl:llama", Init.cpp:11-30)
void setLogSink(std::function<void(LogLevel, const std::string&)> sink);
void logLine(LogLevel level, const std::string& text);

} // namespace bl::llama
