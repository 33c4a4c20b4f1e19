// Http.hpp -- the HTTP/1.1 front end of the inference server (reference server/code/http/HttpServerMain.cpp:298-368).
//
// The reference uses Boost.Beast coroutines on 4 I/O threads; Boost is not part of this build, so this is a plain POSIX-socket
// server with the same observable contract:
//   * one request per connection, then the sending side is shut down (:357);
//   * only POST is served: anything else answers 400 with "Access-Control-Allow-Origin: *" and no body (:307-311);
//   * POST /complete            body = request JSON            -> 200 {"text", "tokenData"}          (:312-318)
//   * POST /verify_completion   body = {"request","response"}  -> 200 {"result": score}              (:328-337)
//   * any other target -> 404 with the CORS header and no body (:350-354); the two /chat/* routes (chat templates are outside
//     this build's scope) answer 501 the same way;
//   * 200 answers carry "Server: Beast", "Content-Type: text/json", "Access-Control-Allow-Origin: *", in that order, then
//     "Connection: close" when the request asked for it, then Content-Length (:262-270);
//   * 4 I/O threads (:445); the inference itself runs on the Server's worker threads, one per GPU replica.
// A body that does not parse terminates the reference process (uncaught exception in a detached coroutine); here it answers
// 400 with {"error": text}.
#pragma once
#include "Server.hpp"

#include <cstdint>
#include <memory>
#include <string>

namespace bl::llama::server {

class HttpFrontEnd {
public:
    // binds host:port (port 0 = any free port) and starts the acceptor + ioThreads I/O threads; throws std::runtime_error when
    // the address cannot be bound
    HttpFrontEnd(Server& server, const std::string& host, uint16_t port, int ioThreads = 4);
    ~HttpFrontEnd();                       // stops accepting, joins the threads
    HttpFrontEnd(const HttpFrontEnd&) = delete;
    HttpFrontEnd& operator=(const HttpFrontEnd&) = delete;

    uint16_t port() const noexcept;        // the bound port
    uint64_t requestsServed() const noexcept;

private:
    struct Impl;
    std::unique_ptr<Impl> m_impl;
};

} // namespace bl::llama::server
