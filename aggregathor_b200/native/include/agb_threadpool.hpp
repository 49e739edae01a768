// Host thread pool used by the CPU (gloo-mode) aggregation rules.
//
// Role of the reference's `native/so_threadpool` + `include/threadpool.hpp`
// (threadpool.cpp:67-166, threadpool.hpp:166-239), redesigned: instead of a job
// queue with one condition variable per job, a `parallel_for` publishes ONE
// range descriptor and the workers claim fixed-size chunks with an atomic
// counter (dynamic load balance, no allocation per call). Chunk boundaries only
// depend on (begin, end, grain), never on timing, so reductions that store one
// partial per chunk and fold them in chunk order are bit-reproducible — the
// reference's CAS-loop accumulation (operations.hpp:46-58) is not.
#pragma once

#include <atomic>
#include <condition_variable>
#include <cstddef>
#include <functional>
#include <mutex>
#include <thread>
#include <vector>

namespace agb {

class ThreadPool {
public:
    using Body = std::function<void(size_t chunk, size_t begin, size_t end)>;
    explicit ThreadPool(size_t nbworkers = 0);
    ~ThreadPool();
    ThreadPool(ThreadPool const&) = delete;
    ThreadPool& operator=(ThreadPool const&) = delete;
    size_t size() const noexcept { return workers.size() + 1; }
    // Number of chunks `run` will produce for this range (chunk ids are 0..count-1).
    static size_t chunk_count(size_t begin, size_t end, size_t grain) noexcept {
        return end <= begin ? 0 : (end - begin + grain - 1) / grain;
    }
    // Blocking parallel loop; the calling thread participates. Re-entrant calls run inline.
    void run(size_t begin, size_t end, size_t grain, Body const& body);
private:
    void worker_loop();
    void drain();
    std::vector<std::thread> workers;
    std::mutex lock;
    std::condition_variable wake, done;
    Body const* body = nullptr;
    size_t begin = 0, end = 0, grain = 1, chunks = 0;
    std::atomic<size_t> next{0};
    size_t active = 0;      // workers currently inside drain()
    unsigned long epoch = 0;
    bool stopping = false;
    std::mutex serialize;   // one parallel region at a time
};

// Process-wide pool (size = hardware concurrency, overridable with AGB_NUM_THREADS).
ThreadPool& global_pool();

template<class Func> inline void parallel_for(size_t begin, size_t end, size_t grain, Func&& f) {
    global_pool().run(begin, end, grain, [&](size_t, size_t b, size_t e) { f(b, e); });
}

} // namespace agb
