// copy_pool.h -- host threads that copy one byte range in parallel (plain C++, no CUDA: tests/host_emul builds it with g++).
#pragma once
#include <condition_variable>
#include <cstdint>
#include <cstring>
#include <mutex>
#include <thread>
#include <vector>

// Host -> device copies of the caller's arrays.  The harness hands over Rust Vecs: pageable memory, which cudaMemcpyAsync
// moves through the driver's bounce buffer with ONE host thread (measured on the B200 box: 8-9 GB/s for the bases of a
// one-call MSM, 14 GB/s for scalars; pinned memory: 53 GB/s).  copy_pool_t is a few host threads that fill the context's
// own pinned slots in parallel; the DMA of slot k overlaps the filling of slot k + 1.
struct copy_pool_t {
    std::vector<std::thread> th;
    std::mutex mu;
    std::condition_variable cv_go, cv_done;
    uint8_t *dst = nullptr; const uint8_t *src = nullptr; size_t bytes = 0;
    uint64_t gen = 0;
    int pending = 0;
    bool stop = false;
    void start(int p) {
        for (int i = 0; i < p; i++) th.emplace_back([this, i, p] {
            uint64_t seen = 0;
            for (;;) {
                uint8_t *d; const uint8_t *s; size_t len;
                {
                    std::unique_lock<std::mutex> lk(mu);
                    cv_go.wait(lk, [&] { return stop || gen != seen; });
                    if (stop) return;
                    seen = gen;
                    const size_t per = ((bytes + p - 1) / p + 63) & ~(size_t)63, lo = per * i < bytes ? per * i : bytes, hi = lo + per < bytes ? lo + per : bytes;
                    d = dst + lo; s = src + lo; len = hi - lo;
                }
                if (len) memcpy(d, s, len);
                std::lock_guard<std::mutex> lk(mu);
                if (--pending == 0) cv_done.notify_one();
            }
        });
    }
    void run(uint8_t *d, const uint8_t *s, size_t n) {
        std::unique_lock<std::mutex> lk(mu);
        dst = d; src = s; bytes = n; pending = (int)th.size(); gen++;
        cv_go.notify_all();
        cv_done.wait(lk, [&] { return pending == 0; });
    }
    ~copy_pool_t() {
        { std::lock_guard<std::mutex> lk(mu); stop = true; }
        cv_go.notify_all();
        for (auto &t : th) t.join();
    }
};
