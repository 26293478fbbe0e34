// CPU test harness of csrc/host_pool.h (test infrastructure): exceptions thrown by pool-thread copies or by the caller's own
// copy must reach the caller of run(), after every copy has been withdrawn or has finished -- never std::terminate, never a
// pool thread still running a function object of a dead stack frame.  Prints "ok" on success.
#include <atomic>
#include <chrono>
#include <cstdio>
#include <stdexcept>
#include <thread>

#include "host_pool.h"

using fheb::HostPool;

static int scenario(int who_throws) {
    // who_throws: 0 nobody, 1 a helper copy, 2 the caller's copy, 3 every copy
    std::atomic<int> started{0}, finished{0};
    const std::thread::id caller = std::this_thread::get_id();
    bool caught = false;
    {
        int frame_local[64];  // helpers touch the caller's frame: they must be done before run() unwinds
        for (int &x : frame_local) x = 0;
        const std::function<void()> fn = [&] {
            const int me = started.fetch_add(1);
            std::this_thread::sleep_for(std::chrono::milliseconds(5 + 3 * me));
            frame_local[me & 63]++;
            const bool is_caller = std::this_thread::get_id() == caller;
            finished.fetch_add(1);
            if (who_throws == 3 || (who_throws == 2 && is_caller) || (who_throws == 1 && !is_caller))
                throw std::runtime_error("boom");
        };
        try {
            HostPool::get().run(3, fn);
        } catch (const std::runtime_error &) {
            caught = true;
        }
        if (started.load() != finished.load()) return 1;  // a copy is still running after run() returned
    }
    if ((who_throws != 0) != caught) {
        // helpers may all have been withdrawn before starting: then scenario 1 legitimately throws nothing
        if (!(who_throws == 1 && started.load() == 1)) return 2;
    }
    return 0;
}

int main() {
    for (int rep = 0; rep < 20; rep++)
        for (int s = 0; s < 4; s++) {
            int rc = scenario(s);
            if (rc) {
                printf("scenario %d failed (%d)\n", s, rc);
                return 1;
            }
        }
    // the pool still works afterwards
    std::atomic<int> n{0};
    const std::function<void()> fn = [&] { n.fetch_add(1); };
    HostPool::get().run(4, fn);
    if (n.load() < 1) return 1;
    printf("ok\n");
    return 0;
}
