// Process-wide pool of host worker threads shared by the batch surface (tiles of calls) and by single calls (the second
// operand of a ct x ct call is inflated on a helper while the caller inflates the first).
#pragma once
#include <algorithm>
#include <condition_variable>
#include <deque>
#include <functional>
#include <memory>
#include <mutex>
#include <thread>

namespace fheb {

// Persistent host workers: spawning 2 x cores threads per batch costs more than a small batch itself.
// run(k, fn) executes fn on up to k pool threads plus the caller and returns when every started copy has finished;
// copies that no pool thread has picked up by the time the caller's own copy returns are withdrawn (fn drains a shared
// work counter, so nothing is lost), which also makes concurrent and nested batches deadlock-free.
class HostPool {
   public:
    static HostPool &get() {
        static HostPool *p = new HostPool();  // never destroyed: worker threads may outlive static destructors
        return *p;
    }
    void run(size_t helpers, const std::function<void()> &fn) {
        auto job = std::make_shared<Job>();
        job->fn = &fn;
        {
            std::lock_guard<std::mutex> lk(mu_);
            const size_t cap = 4 * (size_t)std::max(1u, std::thread::hardware_concurrency());
            while (threads_ < std::min(helpers, cap)) {
                std::thread([this] { loop(); }).detach();
                threads_++;
            }
            for (size_t i = 0; i < helpers; i++) queue_.push_back(job);
            job->queued = helpers;
        }
        cv_.notify_all();
        fn();
        std::unique_lock<std::mutex> lk(mu_);
        for (auto it = queue_.begin(); it != queue_.end();)  // withdraw the copies nobody started
            if (it->get() == job.get()) {
                it = queue_.erase(it);
                job->queued--;
            } else {
                ++it;
            }
        job->done_cv.wait(lk, [&] { return job->queued == 0; });
    }

   private:
    struct Job {
        const std::function<void()> *fn = nullptr;
        size_t queued = 0;  // copies in the queue or running (guarded by mu_)
        std::condition_variable done_cv;
    };
    void loop() {
        std::unique_lock<std::mutex> lk(mu_);
        for (;;) {
            cv_.wait(lk, [&] { return !queue_.empty(); });
            std::shared_ptr<Job> job = queue_.front();
            queue_.pop_front();
            lk.unlock();
            (*job->fn)();
            lk.lock();
            if (--job->queued == 0) job->done_cv.notify_all();
        }
    }
    std::mutex mu_;
    std::condition_variable cv_;
    std::deque<std::shared_ptr<Job>> queue_;
    size_t threads_ = 0;
};

}  // namespace fheb
