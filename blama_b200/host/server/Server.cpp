// Server.cpp -- reference server/code/server/Server.cpp:45-161 with N workers (one per GPU replica).
#include "Server.hpp"

#include "../llama/Instance.hpp"
#include "../llama/LogitComparer.hpp"
#include "../llama/Model.hpp"
#include "../llama/Session.hpp"

#include <blama_b200.h>

#include <condition_variable>
#include <deque>
#include <limits>
#include <mutex>
#include <thread>

namespace bl::llama::server {

struct Server::Impl {
    struct Worker {
        std::shared_ptr<Model> model;
        std::unique_ptr<Instance> instance;
        std::thread thread;
        uint64_t requests = 0;      // guarded by Impl::mu
        double gpuMs = 0;
    };
    struct Job {
        std::function<void(Worker&)> run;
        std::function<void()> failed;      // delivers the request's "no result" answer to its callback
    };

    std::vector<std::unique_ptr<Worker>> workers;
    mutable std::mutex mu;
    std::condition_variable cv, idleCv;
    std::deque<Job> queue;
    size_t running = 0;
    bool stopping = false;
    std::function<void(const std::string&)> onError;

    Impl(std::vector<std::shared_ptr<Model>> replicas, Instance::InitParams ip) {
        for (auto& r : replicas) {
            auto w = std::make_unique<Worker>();
            w->model = std::move(r);
            w->instance = std::make_unique<Instance>(*w->model, ip);
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
            // CUDA events on the worker's own stream bracket the request: the GPU time of this replica, whatever the host did
            blk_ctx* ctx = w.instance->lctx();
            (void)blk_timer_start(ctx);
            std::string error;
            bool ok = true;
            try { job.run(w); }
            catch (const std::exception& e) { ok = false; error = e.what(); }
            catch (...) { ok = false; error = "unknown error"; }
            float ms = 0.0f;
            (void)blk_timer_stop(ctx, &ms);
            if (!ok) {
                // the reference lets the exception escape its io_context and dies; here the session is
                // stopped, the error reported and the request's callback answered, so that nobody waits forever
                w.instance->stopSession();
                std::function<void(const std::string&)> handler;
                { std::lock_guard<std::mutex> lk(mu); handler = onError; }
                try { if (handler) handler(error); } catch (...) {}
                try { if (job.failed) job.failed(); } catch (...) {}
            }
            {
                std::lock_guard<std::mutex> lk(mu);
                w.requests++; w.gpuMs += ms;
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
        auto cbp = std::make_shared<std::function<void(CompleteReponse)>>(std::move(cb));
        Job job;
        job.run = [prompt = std::move(prompt), tokenize, params = std::move(params), cbp](Worker& w) mutable {
            auto& session = w.instance->startSession({.seed = params.seed, .temperature = params.temperature, .topP = params.topP});
            if (tokenize) prompt = w.model->vocab().tokenize(params.prompt, true, true);
            session.setInitialPrompt(prompt);
            auto preds = session.complete({.prompt = {}, .suffix = {}, .maxTokens = int32_t(params.maxTokens)});
            auto response = marshal(*w.model, preds);
            w.instance->stopSession();
            (*cbp)(std::move(response));
        };
        job.failed = [cbp] { (*cbp)({}); };
        post(std::move(job));
    }
    void verify(std::vector<int32_t> prompt, bool tokenize, CompleteRequestParams req, CompleteReponse resp, std::function<void(float)> cb) {
        auto cbp = std::make_shared<std::function<void(float)>>(std::move(cb));
        Job job;
        job.run = [prompt = std::move(prompt), tokenize, req = std::move(req), resp = std::move(resp), cbp](Worker& w) mutable {
            auto& session = w.instance->startSession({.seed = req.seed, .temperature = req.temperature, .topP = req.topP});
            if (tokenize) prompt = w.model->vocab().tokenize(req.prompt, true, true);
            auto orig = unmarshal(resp);
            // setInitialPrompt + fillCtx (reference :135, :149) as one causal prefill over [prompt | response]
            auto mine = session.setInitialPromptAndFill(prompt, orig);
            // the reference pushes one metric at a time and re-sums the history on every push (Server.cpp:153-156); only the last
            // push's value is reported, which is the in-order double sum over all metrics: one push of the whole span
            std::vector<TokenPredictionView> pairs(orig.size());
            for (size_t i = 0; i < orig.size(); i++) pairs[i] = {&orig[i].logits, &mine[i].logits};
            const std::vector<ComparisonMetrics> ms = compareAll(pairs);
            MetricsAggregator agg;
            const float score = ms.empty() ? 0.0f : agg.pushAndVerify(ms);
            w.instance->stopSession();
            (*cbp)(score);
        };
        job.failed = [cbp] { (*cbp)(std::numeric_limits<float>::quiet_NaN()); };
        post(std::move(job));
    }
};

Server::Server(std::shared_ptr<Model> model) : m_impl(std::make_unique<Impl>(std::vector<std::shared_ptr<Model>>{std::move(model)}, Instance::InitParams{})) {}
Server::Server(std::vector<std::shared_ptr<Model>> replicas) : m_impl(std::make_unique<Impl>(std::move(replicas), Instance::InitParams{})) {}
Server::Server(std::vector<std::shared_ptr<Model>> replicas, Instance::InitParams ip) : m_impl(std::make_unique<Impl>(std::move(replicas), ip)) {}
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
void Server::setErrorHandler(std::function<void(const std::string&)> handler) {
    std::lock_guard<std::mutex> lk(m_impl->mu);
    m_impl->onError = std::move(handler);
}
std::vector<Server::WorkerStats> Server::workerStats() const {
    std::lock_guard<std::mutex> lk(m_impl->mu);
    std::vector<WorkerStats> out;
    for (const auto& w : m_impl->workers) out.push_back({w->model->params().device, w->requests, w->gpuMs});
    return out;
}
size_t Server::workerCount() const noexcept { return m_impl->workers.size(); }
void Server::drain() { m_impl->drain(); }

} // namespace bl::llama::server
