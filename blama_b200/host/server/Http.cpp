// Http.cpp -- see Http.hpp.
#include "Http.hpp"
#include "Json.hpp"
#include "Wire.hpp"

#include <arpa/inet.h>
#include <netinet/in.h>
#include <netinet/tcp.h>
#include <poll.h>
#include <sys/socket.h>
#include <unistd.h>

#include <atomic>
#include <cerrno>
#include <cmath>
#include <condition_variable>
#include <cstring>
#include <deque>
#include <future>
#include <mutex>
#include <stdexcept>
#include <thread>
#include <vector>

namespace bl::llama::server {
namespace {

constexpr size_t kMaxHeaderBytes = 64 * 1024;
constexpr size_t kMaxBodyBytes = 256u * 1024 * 1024;     // a 2048-token verify body is ~1.5 MB

struct Request {
    std::string method, target, version;     // version: "HTTP/1.1"
    bool wantsClose = false;                 // "Connection: close", or HTTP/1.0 without keep-alive
    std::string body;
};

bool iequals(std::string_view a, std::string_view b) {
    if (a.size() != b.size()) return false;
    for (size_t i = 0; i < a.size(); ++i) if (std::tolower(static_cast<unsigned char>(a[i])) != std::tolower(static_cast<unsigned char>(b[i]))) return false;
    return true;
}
std::string_view trim(std::string_view s) {
    while (!s.empty() && (s.front() == ' ' || s.front() == '\t')) s.remove_prefix(1);
    while (!s.empty() && (s.back() == ' ' || s.back() == '\t' || s.back() == '\r')) s.remove_suffix(1);
    return s;
}

// reads one request from fd; false when the peer sent something that is not HTTP (the connection is just dropped)
bool readRequest(int fd, Request& req) {
    std::string buf;
    size_t headerEnd = std::string::npos;
    char tmp[16384];
    while (headerEnd == std::string::npos) {
        const ssize_t n = ::recv(fd, tmp, sizeof(tmp), 0);
        if (n <= 0) return false;
        buf.append(tmp, size_t(n));
        headerEnd = buf.find("\r\n\r\n");
        if (headerEnd == std::string::npos && buf.size() > kMaxHeaderBytes) return false;
    }
    const std::string_view head(buf.data(), headerEnd);
    const size_t lineEnd = head.find("\r\n");
    const std::string_view start = head.substr(0, lineEnd);
    const size_t sp1 = start.find(' '), sp2 = start.rfind(' ');
    if (sp1 == std::string_view::npos || sp2 == sp1) return false;
    req.method = std::string(start.substr(0, sp1));
    req.target = std::string(start.substr(sp1 + 1, sp2 - sp1 - 1));
    req.version = std::string(start.substr(sp2 + 1));
    if (req.version.rfind("HTTP/", 0) != 0) return false;
    size_t contentLength = 0;
    bool keepAliveHeader = false, closeHeader = false, expectContinue = false, chunked = false;
    size_t pos = lineEnd == std::string_view::npos ? head.size() : lineEnd + 2;
    while (pos < head.size()) {
        size_t e = head.find("\r\n", pos);
        if (e == std::string_view::npos) e = head.size();
        const std::string_view line = head.substr(pos, e - pos);
        pos = e + 2;
        const size_t colon = line.find(':');
        if (colon == std::string_view::npos) continue;
        const std::string_view name = trim(line.substr(0, colon)), value = trim(line.substr(colon + 1));
        if (iequals(name, "Content-Length")) contentLength = size_t(std::strtoull(std::string(value).c_str(), nullptr, 10));
        else if (iequals(name, "Connection")) { if (iequals(value, "close")) closeHeader = true; else if (iequals(value, "keep-alive")) keepAliveHeader = true; }
        else if (iequals(name, "Expect") && iequals(value, "100-continue")) expectContinue = true;
        else if (iequals(name, "Transfer-Encoding") && iequals(value, "chunked")) chunked = true;
    }
    req.wantsClose = closeHeader || (req.version == "HTTP/1.0" && !keepAliveHeader);
    if (contentLength > kMaxBodyBytes) return false;
    if (expectContinue) { const char c100[] = "HTTP/1.1 100 Continue\r\n\r\n"; (void)::send(fd, c100, sizeof(c100) - 1, MSG_NOSIGNAL); }
    req.body.assign(buf, headerEnd + 4, std::string::npos);
    if (chunked) {
        // de-chunk: read until the terminating 0-size chunk
        std::string raw = std::move(req.body), out;
        size_t i = 0;
        for (;;) {
            size_t eol;
            while ((eol = raw.find("\r\n", i)) == std::string::npos) {
                const ssize_t n = ::recv(fd, tmp, sizeof(tmp), 0);
                if (n <= 0) return false;
                raw.append(tmp, size_t(n));
            }
            const size_t len = size_t(std::strtoull(raw.substr(i, eol - i).c_str(), nullptr, 16));
            i = eol + 2;
            if (len == 0) break;
            if (out.size() + len > kMaxBodyBytes) return false;
            while (raw.size() < i + len + 2) {
                const ssize_t n = ::recv(fd, tmp, sizeof(tmp), 0);
                if (n <= 0) return false;
                raw.append(tmp, size_t(n));
            }
            out.append(raw, i, len);
            i += len + 2;
        }
        req.body = std::move(out);
        return true;
    }
    while (req.body.size() < contentLength) {
        const ssize_t n = ::recv(fd, tmp, std::min(sizeof(tmp), contentLength - req.body.size()), 0);
        if (n <= 0) return false;
        req.body.append(tmp, size_t(n));
    }
    req.body.resize(contentLength);
    return true;
}

void sendAll(int fd, const std::string& data) {
    size_t off = 0;
    while (off < data.size()) {
        const ssize_t n = ::send(fd, data.data() + off, data.size() - off, MSG_NOSIGNAL);
        if (n <= 0) { if (errno == EINTR) continue; return; }
        off += size_t(n);
    }
}

const char* reason(int status) {
    switch (status) { case 200: return "OK"; case 400: return "Bad Request"; case 404: return "Not Found"; case 500: return "Internal Server Error"; case 501: return "Not Implemented"; default: return "Unknown"; }
}

// 200-style answer with a body (getCompleteResponse / getVerifyResponse, reference :262-270, :279-285)
std::string bodyResponse(const Request& req, int status, const std::string& body) {
    std::string out = req.version + " " + std::to_string(status) + " " + reason(status) + "\r\n";
    out += "Server: Beast\r\nContent-Type: text/json\r\nAccess-Control-Allow-Origin: *\r\n";
    if (req.wantsClose) out += "Connection: close\r\n";
    out += "Content-Length: " + std::to_string(body.size()) + "\r\n\r\n";
    out += body;
    return out;
}
// empty_body answers (400 for non-POST, 404 for unknown targets, reference :307-311, :350-354): only the CORS header
std::string emptyResponse(const Request& req, int status) {
    return req.version + " " + std::to_string(status) + " " + reason(status) + "\r\nAccess-Control-Allow-Origin: *\r\n\r\n";
}
std::string errorBody(const std::string& what) {
    json::Object o; o["error"] = what;
    return json::Value(std::move(o)).dump();
}

} // namespace

struct HttpFrontEnd::Impl {
    Server& server;
    int listenFd = -1;
    uint16_t boundPort = 0;
    std::atomic<bool> stopping{false};
    std::atomic<uint64_t> served{0};
    std::thread acceptor;
    std::vector<std::thread> io;
    std::mutex mu;
    std::condition_variable cv;
    std::deque<int> pending;

    explicit Impl(Server& s) : server(s) {}

    std::string handle(const Request& req) {
        if (req.method != "POST") return emptyResponse(req, 400);
        try {
            if (req.target == "/complete") {
                auto params = wire::parseCompleteParams(req.body);
                std::promise<Server::CompleteReponse> done;
                auto fut = done.get_future();
                server.completeText(std::move(params), [&done](Server::CompleteReponse r) { done.set_value(std::move(r)); });
                return bodyResponse(req, 200, wire::completeResponseJson(fut.get()));
            }
            if (req.target == "/verify_completion") {
                auto body = wire::parseVerifyBody(req.body);
                std::promise<float> done;
                auto fut = done.get_future();
                server.verify(std::move(body.request), std::move(body.response), [&done](float s) { done.set_value(s); });
                const float score = fut.get();
                if (std::isnan(score)) return bodyResponse(req, 500, errorBody("verification failed"));
                return bodyResponse(req, 200, wire::verifyResponseJson(score));
            }
            if (req.target == "/chat/completions" || req.target == "/chat/verify_completion") return emptyResponse(req, 501);
            return emptyResponse(req, 404);
        } catch (const json::ParseError& e) { return bodyResponse(req, 400, errorBody(e.what()));
        } catch (const json::TypeError& e) { return bodyResponse(req, 400, errorBody(e.what()));
        } catch (const std::exception& e) { return bodyResponse(req, 500, errorBody(e.what())); }
    }

    void serve(int fd) {
        Request req;
        if (readRequest(fd, req)) {
            sendAll(fd, handle(req));
            served++;
        }
        ::shutdown(fd, SHUT_WR);               // "Close the stream" (:357): one request per connection
        char sink[256];
        struct pollfd p{fd, POLLIN, 0};
        while (::poll(&p, 1, 200) > 0 && ::recv(fd, sink, sizeof(sink), 0) > 0) {}      // let the peer finish reading
        ::close(fd);
    }
    void ioLoop() {
        for (;;) {
            int fd;
            {
                std::unique_lock<std::mutex> lk(mu);
                cv.wait(lk, [&] { return stopping.load() || !pending.empty(); });
                if (pending.empty()) return;
                fd = pending.front();
                pending.pop_front();
            }
            serve(fd);
        }
    }
    void acceptLoop() {
        while (!stopping.load()) {
            struct pollfd p{listenFd, POLLIN, 0};
            const int r = ::poll(&p, 1, 100);
            if (r <= 0) continue;
            const int fd = ::accept(listenFd, nullptr, nullptr);
            if (fd < 0) continue;
            int one = 1;
            ::setsockopt(fd, IPPROTO_TCP, TCP_NODELAY, &one, sizeof(one));
            { std::lock_guard<std::mutex> lk(mu); pending.push_back(fd); }
            cv.notify_one();
        }
    }
};

HttpFrontEnd::HttpFrontEnd(Server& server, const std::string& host, uint16_t port, int ioThreads) : m_impl(std::make_unique<Impl>(server)) {
    Impl& s = *m_impl;
    sockaddr_in addr{};
    addr.sin_family = AF_INET;
    addr.sin_port = htons(port);
    if (::inet_pton(AF_INET, host.c_str(), &addr.sin_addr) != 1) throw std::invalid_argument("Invalid BLAMA_HOST");     // reference :389
    s.listenFd = ::socket(AF_INET, SOCK_STREAM, 0);
    if (s.listenFd < 0) throw std::runtime_error(std::string("socket: ") + std::strerror(errno));
    int one = 1;
    ::setsockopt(s.listenFd, SOL_SOCKET, SO_REUSEADDR, &one, sizeof(one));
    if (::bind(s.listenFd, reinterpret_cast<sockaddr*>(&addr), sizeof(addr)) != 0 || ::listen(s.listenFd, 128) != 0) {
        const std::string why = std::strerror(errno);
        ::close(s.listenFd);
        throw std::runtime_error("cannot listen on " + host + ":" + std::to_string(port) + ": " + why);
    }
    socklen_t len = sizeof(addr);
    ::getsockname(s.listenFd, reinterpret_cast<sockaddr*>(&addr), &len);
    s.boundPort = ntohs(addr.sin_port);
    for (int i = 0; i < std::max(1, ioThreads); ++i) s.io.emplace_back([&s] { s.ioLoop(); });
    s.acceptor = std::thread([&s] { s.acceptLoop(); });
}

HttpFrontEnd::~HttpFrontEnd() {
    Impl& s = *m_impl;
    s.stopping = true;
    s.cv.notify_all();
    if (s.acceptor.joinable()) s.acceptor.join();
    for (auto& t : s.io) if (t.joinable()) t.join();
    for (int fd : s.pending) ::close(fd);
    if (s.listenFd >= 0) ::close(s.listenFd);
}

uint16_t HttpFrontEnd::port() const noexcept { return m_impl->boundPort; }
uint64_t HttpFrontEnd::requestsServed() const noexcept { return m_impl->served.load(); }

} // namespace bl::llama::server
