// hostio.cu -- host side of the copy-engine prefetch / write-back of the look-ahead cache.
//
// The reference's Prefetcher gathers the rows of a window's unique ids out of the CPU master tables with a worker pool
// and ships them to the trainers (cache_manager.py:27-46, `emb_tables_cpu.emb_l[i].weight[unique_idxs]`), and its
// eviction manager writes evicted rows back with `weight[idxs] = rows` (cache_manager.py:48-64).  The SM-driven
// alternative (rows_kernel<MODE 1 / 5> in move.cu: zero-copy loads / stores of the pinned master) needs no host thread,
// but every system-memory access it has in flight slows the training kernels that share the GPU.
// Here the scattered side of the transfer runs on HOST threads -- gather into / scatter out of a pinned, contiguous
// staging chunk -- and the PCIe side is a plain cudaMemcpyAsync on a copy engine (north_star (1): "streams missed rows
// from pinned host master tables with cudaMemcpyAsync on a side stream while writing evicted dirty lines back").
// No arithmetic happens here except the optional (W + row) / 2 of --average-on-writeback.
#include <thread>

#include "common.cuh"

namespace {

template <typename Fn>
void parallel_rows(int64_t n, int threads, Fn&& fn) {
    if (threads < 1) threads = 1;
    if (n < 4096 || threads == 1) {
        fn(0, n);
        return;
    }
    std::vector<std::thread> pool;
    const int64_t per = (n + threads - 1) / threads;
    for (int t = 0; t < threads; ++t) {
        const int64_t lo = t * per, hi = lo + per < n ? lo + per : n;
        if (lo >= hi) break;
        pool.emplace_back([=, &fn] { fn(lo, hi); });
    }
    for (auto& th : pool) th.join();
}

}  // namespace

// dst[i, :] = master[ids[i], :] for i in [0, n): ids ascending or not, all host pointers
extern "C" int cdlrm_host_gather_rows(const float* master, int64_t n_rows, int dim, const int64_t* ids, int64_t n,
                                      float* dst, int threads) {
    ARG_CHECK(master && ids && dst && dim > 0 && n >= 0 && n_rows > 0);
    const size_t row_b = (size_t)dim * sizeof(float);
    bool bad = false;
    parallel_rows(n, threads, [&](int64_t lo, int64_t hi) {
        for (int64_t i = lo; i < hi; ++i) {
            // a few rows ahead: the table is far larger than the caches, every row is a DRAM miss
            if (i + 8 < hi) __builtin_prefetch(master + ids[i + 8] * (int64_t)dim, 0, 0);
            const int64_t id = ids[i];
            if ((uint64_t)id >= (uint64_t)n_rows) { bad = true; continue; }
            memcpy(dst + i * (int64_t)dim, master + id * (int64_t)dim, row_b);
        }
    });
    if (bad) {
        cdlrm_set_error("cdlrm_host_gather_rows: id outside its table");
        return CDLRM_ERR_ARG;
    }
    return CDLRM_OK;
}

// master[ids[i], :] = src[i, :] (or the mean of the two with `average`) for every i with primary[i] != 0 (primary ==
// NULL: every i).  Duplicate ids carry identical rows and exactly one of them is primary (plan.cu: lists_kernel), so
// rows are disjoint across threads.
extern "C" int cdlrm_host_scatter_rows(float* master, int64_t n_rows, int dim, const int64_t* ids, const uint8_t* primary,
                                       int64_t n, const float* src, int average, int threads) {
    ARG_CHECK(master && ids && src && dim > 0 && n >= 0 && n_rows > 0);
    const size_t row_b = (size_t)dim * sizeof(float);
    bool bad = false;
    parallel_rows(n, threads, [&](int64_t lo, int64_t hi) {
        for (int64_t i = lo; i < hi; ++i) {
            if (primary && !primary[i]) continue;
            const int64_t id = ids[i];
            if ((uint64_t)id >= (uint64_t)n_rows) { bad = true; continue; }
            float* w = master + id * (int64_t)dim;
            const float* r = src + i * (int64_t)dim;
            if (average) {
                for (int c = 0; c < dim; ++c) w[c] = (w[c] + r[c]) / 2;
            } else {
                memcpy(w, r, row_b);
            }
        }
    });
    if (bad) {
        cdlrm_set_error("cdlrm_host_scatter_rows: id outside its table");
        return CDLRM_ERR_ARG;
    }
    return CDLRM_OK;
}

// plain cudaMemcpyAsync between a (pinned) host chunk and device memory on `stream`: kind 1 = host to device,
// 2 = device to host.  Runs on a copy engine; the destination may be any device address (a peer-readable shard too).
extern "C" int cdlrm_copy_async(int device, void* dst, const void* src, int64_t bytes, int kind, cdlrm_stream stream) {
    ARG_CHECK(dst && src && bytes >= 0 && (kind == 1 || kind == 2));
    if (bytes == 0) return CDLRM_OK;
    CU_CHECK(cudaSetDevice(device));
    // in pieces: a copy engine does not preempt a copy, and the training step's own small copies (inputs in, loss out)
    // would otherwise wait for a whole 128 MB chunk (2.3 ms at 55 GB/s, measured as 2.5-3 ms spikes of single steps)
    constexpr int64_t PIECE = 4 << 20;
    for (int64_t o = 0; o < bytes; o += PIECE) {
        const int64_t m = bytes - o < PIECE ? bytes - o : PIECE;
        CU_CHECK(cudaMemcpyAsync((char*)dst + o, (const char*)src + o, (size_t)m,
                                 kind == 1 ? cudaMemcpyHostToDevice : cudaMemcpyDeviceToHost, (cudaStream_t)stream));
    }
    return CDLRM_OK;
}
