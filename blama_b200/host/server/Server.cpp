// Server.cpp -- reference server/code/server/Server.cpp:45-161 with N workers (one per GPU replica).
#include "Server.hpp"

#include "../llama/Instance.hpp"
#include "../llama/LogitComparer.hpp"
#include "../llama/Model.hpp"
#include "../llama/Session.hpp"

#include <condition_variable>
#include <deque>
#include <mutex>
#include <thread>

namespace bl::llama::server {

struct Server::Impl {
    struct Worker {
        std::shared_ptr<Model> model;
        std::unique_ptr<Instance> instance;
        std::thread thread;
    };
    using Job = std::function<void(Worker&)>;

    std::vector<std::unique_ptr<Worker>> workers;
    std::mutex mu;
    std::condition_variable cv, idleCv;
    std::deque<Job> queue;
    size_t running = 0;
    bool stopping = false;

    explicit Impl(std::vector<std::shared_ptr<Model>> replicas) {
        for (auto& r : replicas) {
            auto w = std::make_unique<Worker>();
            w->model = std::move(r);
            w->instance = std::make_unique<Instance>(*w->model, Instance::InitParams{});
            w->instance->warmup();
            workers.push_back(std::move(w));
        }
        for (auto& w : workers) w->thread = std::thread([this, wp = w.get()] { loop(*wp); });
    }
    ~Impl() {
        { std::lock_guard<std::mutex> lk(mu); stopping = true; }
        cv.notify_all();
        for (auto& w : workers) if (w->thread.joinable()) w->thread.join();
    }
    void post(Job j) {
        { std::lock_guard<std::mutex> lk(mu); queue.push_back(std::move(j)); }
        cv.notify_one();
    }
    void loop(Worker& w) {
        for (;;) {
            Job job;
            {
                std::unique_lock<std::mutex> lk(mu);
                cv.wait(lk, [&] { return stopping || !queue.empty(); });
                if (queue.empty()) return;
                job = std::move(queue.front());
                queue.pop_front();
                ++running;
            }
            // the reference lets exceptions escape its io_context and terminate; here a failing
            // request is dropped after stopping its session so the worker survives
            try { job(w); } catch (...) { w.instance->stopSession(); }
            {
                std::lock_guard<std::mutex> lk(mu);
                --running;
                if (queue.empty() && running == 0) idleCv.notify_all();
            }
        }
    }
    void drain() {
        std::unique_lock<std::mutex> lk(mu);
        idleCv.wait(lk, [&] { return queue.empty() && running == 0; });
    }

    static CompleteReponse marshal(const Model& model, const std::vector<TokenPrediction>& preds) {
        CompleteReponse response;
        response.reserve(preds.size());
        for (const auto& p : preds) {
            auto& td = response.emplace_back();
            td.tokenStr = model.vocab().tokenToString(p.token);
            td.tokenId = uint32_t(p.token);
            td.logits.reserve(p.logits.size());
            for (const auto& l : p.logits) td.logits.push_back({uint32_t(l.token), l.logit});
        }
        return response;
    }
    static std::vector<TokenPrediction> unmarshal(const CompleteReponse& resp) {
        std::vector<TokenPrediction> preds;
        preds.reserve(resp.size());
        for (const auto& t : resp) {
            auto& p = preds.emplace_back();
            p.token = Token(t.tokenId);
            p.logits.reserve(t.logits.size());
            for (const auto& l : t.logits) p.logits.push_back({Token(l.tokenId), l.logit});
        }
        return preds;
    }

    void complete(std::vector<int32_t> prompt, bool tokenize, CompleteRequestParams params, std::function<void(CompleteReponse)> cb) {
        post([prompt = std::move(prompt), tokenize, params = std::move(params), cb = std::move(cb)](Worker& w) mutable {
            auto& session = w.instance->startSession({.seed = params.seed, .temperature = params.temperature, .topP = params.topP});
            if (tokenize) prompt = w.model->vocab().tokenize(params.prompt, true, true);
            session.setInitialPrompt(prompt);
            auto preds = session.complete({.prompt = {}, .suffix = {}, .maxTokens = int32_t(params.maxTokens)});
            cb(marshal(*w.model, preds));
            w.instance->stopSession();
        });
    }
    void verify(std::vector<int32_t> prompt, bool tokenize, CompleteRequestParams req, CompleteReponse resp, std::function<void(float)> cb) {
        post([prompt = std::move(prompt), tokenize, req = std::move(req), resp = std::move(resp), cb = std::move(cb)](Worker& w) mutable {
            auto& session = w.instance->startSession({.seed = req.seed, .temperature = req.temperature, .topP = req.topP});
            if (tokenize) prompt = w.model->vocab().tokenize(req.prompt, true, true);
            session.setInitialPrompt(prompt);
            auto orig = unmarshal(resp);
            auto mine = session.fillCtx(orig);
            // the reference pushes one metric at a time and re-sums the history on every push (Server.cpp:153-156); only the last
            // push's value is reported, which is the in-order double sum over all metrics: one push of the whole span
            std::vector<TokenPredictionView> pairs(orig.size());
            for (size_t i = 0; i < orig.size(); i++) pairs[i] = {&orig[i].logits, &mine[i].logits};
            const std::vector<ComparisonMetrics> ms = compareAll(pairs);
            MetricsAggregator agg;
            cb(ms.empty() ? 0.0f : agg.pushAndVerify(ms));
            w.instance->stopSession();
        });
    }
};

Server::Server(std::shared_ptr<Model> model) : m_impl(std::make_unique<Impl>(std::vector<std::shared_ptr<Model>>{std::move(model)})) {}
Server::Server(std::vector<std::shared_ptr<Model>> replicas) : m_impl(std::make_unique<Impl>(std::move(replicas))) {}
Server::~Server() = default;

void Server::completeText(CompleteRequestParams params, std::function<void(CompleteReponse)> cb) {
    m_impl->complete({}, true, std::move(params), std::move(cb));
}
void Server::verify(CompleteRequestParams req, CompleteReponse resp, std::function<void(float)> cb) {
    m_impl->verify({}, true, std::move(req), std::move(resp), std::move(cb));
}
void Server::completeTokens(std::vector<int32_t> prompt, CompleteRequestParams params, std::function<void(CompleteReponse)> cb) {
    m_impl->complete(std::move(prompt), false, std::move(params), std::move(cb));
}
void Server::verifyTokens(std::vector<int32_t> prompt, CompleteRequestParams req, CompleteReponse resp, std::function<void(float)> cb) {
    m_impl->verify(std::move(prompt), false, std::move(req), std::move(resp), std::move(cb));
}
size_t Server::workerCount() const noexcept { return m_impl->workers.size(); }
void Server::drain() { m_impl->drain(); }

} // namespace bl::llama::server
