// fs_host_par.hpp -- the few host loops over a whole (replicated) mesh run on a handful of threads.
#pragma once
#include <algorithm>
#include <cstdint>
#include <cstdlib>
#include <thread>
#include <vector>

namespace fs {

// threads a set-up loop may use: FS_HOST_THREADS, else min(8, cores / ranks sharing the box)
inline int host_threads(int ranks_on_box = 1)
{
    if (const char *e = getenv("FS_HOST_THREADS")) return std::max(1, std::min(64, atoi(e)));
    const unsigned hw = std::thread::hardware_concurrency();
    return (int)std::max<int64_t>(1, std::min<int64_t>(8, (int64_t)(hw ? hw : 1) / std::max(1, ranks_on_box)));
}

// fn(t, begin, end) on T contiguous pieces of [0, n); pieces are independent; returns the number of pieces used
template <class F>
inline int parallel_chunks(int64_t n, int threads, F fn)
{
    const int64_t T = std::max<int64_t>(1, std::min<int64_t>(threads, n / 32768));
    if (T == 1) {
        fn(0, (int64_t)0, n);
        return 1;
    }
    std::vector<std::thread> pool;
    for (int64_t t = 0; t < T; t++) pool.emplace_back(fn, (int)t, n * t / T, n * (t + 1) / T);
    for (std::thread &th : pool) th.join();
    return (int)T;
}

}  // namespace fs
