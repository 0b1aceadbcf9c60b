// host_copy.hpp -- host-side helpers of the download path towards PAGEABLE caller memory (api.cu): a small pool of
// copy threads and the page-residency probe that decides when the bounce path pays.  Plain C++ (no CUDA), so that
// tests/test_host_copy.py can exercise it without a GPU.
#pragma once
#include <sys/mman.h>
#include <unistd.h>

#include <algorithm>
#include <condition_variable>
#include <cstdint>
#include <cstring>
#include <deque>
#include <mutex>
#include <thread>
#include <vector>

namespace {

// A few helper threads that copy one large block between host buffers in parallel.  A download into PAGEABLE caller
// memory (a Rust Vec, a numpy array) is what the reference's call sites hand over; cudaMemcpy into it runs at
// ~2 GB/s when the pages have never been touched (a fresh allocation) because the driver faults and pins them one
// by one.  The library instead downloads into its own pinned bounce buffers at link speed and copies from there
// with these threads, so the first-touch page faults of the destination are taken in parallel.
class ParallelCopier {
    struct Job { unsigned char *dst; const unsigned char *src; size_t n; };
    std::vector<std::thread> th_;
    std::mutex mu_;
    std::condition_variable cv_, done_;
    std::deque<Job> q_;
    size_t pending_ = 0;
    bool stop_ = false;
    void run()
    {
        std::unique_lock<std::mutex> lk(mu_);
        for (;;) {
            cv_.wait(lk, [&] { return stop_ || !q_.empty(); });
            if (q_.empty()) return;
            const Job j = q_.front();
            q_.pop_front();
            lk.unlock();
            std::memcpy(j.dst, j.src, j.n);
            lk.lock();
            if (--pending_ == 0) done_.notify_all();
        }
    }

public:
    explicit ParallelCopier(unsigned helpers)
    {
        try {
            for (unsigned i = 0; i < helpers; ++i) th_.emplace_back([this] { run(); });
        } catch (...) {
            // no more threads to be had: the ones that started (possibly none) and the caller do the copying
        }
    }
    ~ParallelCopier()
    {
        {
            std::lock_guard<std::mutex> lk(mu_);
            stop_ = true;
        }
        cv_.notify_all();
        for (auto &t : th_) t.join();
    }
    // copies n bytes; returns when all of them are in dst (the caller copies a share itself)
    void copy(void *dst_v, const void *src_v, size_t n)
    {
        unsigned char *dst = static_cast<unsigned char *>(dst_v);
        const unsigned char *src = static_cast<const unsigned char *>(src_v);
        const size_t parts = th_.size() + 1;
        size_t share = (n / parts + 4095) & ~size_t(4095);          // whole pages per thread
        if (th_.empty() || n < (size_t(256) << 10)) share = n;
        size_t done = std::min(share, n);                            // [0, done) is the caller's own share
        {
            std::lock_guard<std::mutex> lk(mu_);
            for (size_t o = done; o < n; o += share) {
                q_.push_back(Job{dst + o, src + o, std::min(share, n - o)});
                ++pending_;
            }
        }
        cv_.notify_all();
        std::memcpy(dst, src, done);
        std::unique_lock<std::mutex> lk(mu_);
        done_.wait(lk, [&] { return pending_ == 0; });
    }
};

// true if most of the pages of [p, p + bytes) have never been touched (a fresh allocation): sampled with mincore
inline bool host_is_fresh(const void *p, size_t bytes)
{
    const long page = sysconf(_SC_PAGESIZE);
    if (page <= 0 || bytes < (size_t)page * 2) return false;
    const uintptr_t lo = reinterpret_cast<uintptr_t>(p) & ~(uintptr_t)(page - 1);
    const size_t pages = (reinterpret_cast<uintptr_t>(p) + bytes - lo + page - 1) / page;
    unsigned absent = 0, seen = 0;
    for (unsigned k = 0; k < 16; ++k) {
        unsigned char vec = 0;
        void *addr = reinterpret_cast<void *>(lo + (pages * k / 16) * (size_t)page);
        if (mincore(addr, (size_t)page, &vec) != 0) continue;      // not mapped the way mincore likes: treat as touched
        ++seen;
        if (!(vec & 1u)) ++absent;
    }
    return seen >= 8 && absent * 2 > seen;
}

}  // namespace
