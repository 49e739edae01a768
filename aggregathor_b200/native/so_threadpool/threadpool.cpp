// See include/agb_threadpool.hpp.
#include <agb_threadpool.hpp>

#include <cstdlib>

namespace agb {

static thread_local bool inside_region = false;

ThreadPool::ThreadPool(size_t nbworkers) {
    if (nbworkers == 0) {
        if (char const* env = std::getenv("AGB_NUM_THREADS"))
            nbworkers = static_cast<size_t>(std::strtoul(env, nullptr, 10));
        if (nbworkers == 0)
            nbworkers = std::thread::hardware_concurrency();
        if (nbworkers == 0)
            nbworkers = 4;
    }
    for (size_t i = 1; i < nbworkers; ++i)
        workers.emplace_back([this] { worker_loop(); });
}

ThreadPool::~ThreadPool() {
    {
        std::lock_guard<std::mutex> guard{lock};
        stopping = true;
    }
    wake.notify_all();
    for (auto& worker: workers)
        worker.join();
}

void ThreadPool::drain() {
    for (;;) {
        size_t chunk = next.fetch_add(1, std::memory_order_relaxed);
        if (chunk >= chunks)
            return;
        size_t b = begin + chunk * grain;
        size_t e = b + grain < end ? b + grain : end;
        (*body)(chunk, b, e);
    }
}

void ThreadPool::worker_loop() {
    inside_region = true; // nested parallel_for from a worker runs inline
    unsigned long seen = 0;
    std::unique_lock<std::mutex> guard{lock};
    for (;;) {
        wake.wait(guard, [&] { return stopping || epoch != seen; });
        if (stopping)
            return;
        seen = epoch;
        ++active;
        guard.unlock();
        drain();
        guard.lock();
        if (--active == 0)
            done.notify_all();
    }
}

void ThreadPool::run(size_t b, size_t e, size_t g, Body const& f) {
    if (g == 0)
        g = 1;
    size_t count = chunk_count(b, e, g);
    if (count == 0)
        return;
    if (inside_region || count == 1 || workers.empty()) { // inline, same chunking (=> same results)
        for (size_t chunk = 0; chunk < count; ++chunk) {
            size_t cb = b + chunk * g;
            f(chunk, cb, cb + g < e ? cb + g : e);
        }
        return;
    }
    std::lock_guard<std::mutex> region{serialize};
    {
        std::unique_lock<std::mutex> guard{lock};
        done.wait(guard, [&] { return active == 0; }); // late wakers of the previous region must be out of drain()
        body = &f; begin = b; end = e; grain = g; chunks = count;
        next.store(0, std::memory_order_relaxed);
        ++epoch;
    }
    wake.notify_all();
    inside_region = true;
    drain();
    inside_region = false;
    std::unique_lock<std::mutex> guard{lock};
    // Every chunk has been claimed; wait for the workers still executing theirs. A worker that
    // has not woken up yet will find no chunk left, which is fine: `active` only counts
    // workers inside drain(), and late wakers see next >= chunks of *this* epoch or a newer one.
    done.wait(guard, [&] { return active == 0; });
    body = nullptr;
}

ThreadPool& global_pool() {
    static ThreadPool pool;
    return pool;
}

} // namespace agb
