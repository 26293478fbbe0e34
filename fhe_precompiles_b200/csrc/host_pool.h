// Process-wide pool of host worker threads shared by the batch surface (tiles of calls) and by single calls (the second
// operand of a ct x ct call is inflated on a helper while the caller inflates the first).
#pragma once
#include <algorithm>
#include <condition_variable>
#include <deque>
#include <exception>
#include <functional>
#include <memory>
#include <mutex>
#include <thread>

namespace fheb {

// Persistent host workers: spawning 2 x cores threads per batch costs more than a small batch itself.
// run(k, fn) executes fn on up to k pool threads plus the caller and returns when every started copy has finished;
// copies that no pool thread has picked up by the time the caller's own copy returns are withdrawn (fn drains a shared
// work counter, so nothing is lost), which also makes concurrent and nested batches deadlock-free.
//
// Exceptions: fn may throw on any thread (a CUDA allocation failure, a missing libzstd, bad_alloc).  A pool thread never
// lets one escape (that would be std::terminate for the whole host process): the first exception of a job is kept in the
// job and rethrown on the CALLER once every copy has been withdrawn or has finished, so by the time run() unwinds nobody
// holds a pointer into the caller's frame any more.  The same withdrawal + wait runs when the caller's own copy throws.
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
        std::exception_ptr mine;
        try {
            fn();
        } catch (...) {
            mine = std::current_exception();
        }
        std::exception_ptr theirs;
        {
            std::unique_lock<std::mutex> lk(mu_);
            for (auto it = queue_.begin(); it != queue_.end();)  // withdraw the copies nobody started
                if (it->get() == job.get()) {
                    it = queue_.erase(it);
                    job->queued--;
                } else {
                    ++it;
                }
            job->done_cv.wait(lk, [&] { return job->queued == 0; });
            theirs = job->error;
            job->fn = nullptr;  // the frame `fn` lives in is about to go away
        }
        if (mine) std::rethrow_exception(mine);
        if (theirs) std::rethrow_exception(theirs);
    }

   private:
    struct Job {
        const std::function<void()> *fn = nullptr;
        size_t queued = 0;  // copies in the queue or running (guarded by mu_)
        std::exception_ptr error;  // first exception thrown by a pool-thread copy (guarded by mu_)
        std::condition_variable done_cv;
    };
    void loop() {
        std::unique_lock<std::mutex> lk(mu_);
        for (;;) {
            cv_.wait(lk, [&] { return !queue_.empty(); });
            std::shared_ptr<Job> job = queue_.front();
            queue_.pop_front();
            const std::function<void()> *fn = job->fn;
            lk.unlock();
            std::exception_ptr err;
            try {
                if (fn) (*fn)();
            } catch (...) {
                err = std::current_exception();
            }
            lk.lock();
            if (err && !job->error) job->error = err;
            if (--job->queued == 0) job->done_cv.notify_all();
        }
    }
    std::mutex mu_;
    std::condition_variable cv_;
    std::deque<std::shared_ptr<Job>> queue_;
    size_t threads_ = 0;
};

}  // namespace fheb
