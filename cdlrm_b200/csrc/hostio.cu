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
#include <atomic>
#include <condition_variable>
#include <functional>
#include <mutex>
#include <thread>

#include "common.cuh"

namespace {

// Persistent worker pool: a transfer is hundreds of 8 MB chunks, and creating threads per chunk maps and unmaps thread
// stacks (glibc caches 40 MB of them: from 6 threads on every chunk paid an mmap / munmap, i.e. the process-wide
// mmap lock, and the thread that enqueues the training steps stalled 15-20 ms at a time beside it).  Workers are
// created once, sleep on a condition variable and are never joined (the pool lives as long as the process).
class RowPool {
public:
    void run(int64_t n, int threads, const std::function<void(int64_t, int64_t)>& fn) {
        if (threads < 1) threads = 1;
        if (n < 4096 || threads == 1) {
            fn(0, n);
            return;
        }
        std::unique_lock<std::mutex> call(call_mu_);        // one transfer at a time uses the pool
        const int64_t per = (n + threads - 1) / threads;
        int parts = (int)((n + per - 1) / per);
        {
            std::unique_lock<std::mutex> lk(mu_);
            while ((int)workers_.size() < parts - 1) {
                const int id = (int)workers_.size();
                workers_.emplace_back([this, id] { loop(id); });
                workers_.back().detach();
            }
            fn_ = &fn;
            n_ = n;
            per_ = per;
            parts_ = parts;
            pending_ = parts - 1;
            ++gen_;
        }
        cv_.notify_all();
        fn(0, per < n ? per : n);                            // the caller is worker 0
        std::unique_lock<std::mutex> lk(mu_);
        done_.wait(lk, [this] { return pending_ == 0; });
        fn_ = nullptr;
    }

private:
    void loop(int id) {
        uint64_t seen = 0;
        for (;;) {
            const std::function<void(int64_t, int64_t)>* fn;
            int64_t lo, hi;
            {
                std::unique_lock<std::mutex> lk(mu_);
                cv_.wait(lk, [&] { return gen_ != seen; });
                seen = gen_;
                if (id + 1 >= parts_) continue;              // not needed for this call
                fn = fn_;
                lo = (int64_t)(id + 1) * per_;
                hi = lo + per_ < n_ ? lo + per_ : n_;
            }
            if (lo < hi) (*fn)(lo, hi);
            std::unique_lock<std::mutex> lk(mu_);
            if (--pending_ == 0) done_.notify_all();
        }
    }
    std::mutex call_mu_, mu_;
    std::condition_variable cv_, done_;
    std::vector<std::thread> workers_;
    const std::function<void(int64_t, int64_t)>* fn_ = nullptr;
    int64_t n_ = 0, per_ = 0;
    int parts_ = 0, pending_ = 0;
    uint64_t gen_ = 0;
};

RowPool& pool() {
    static RowPool* p = new RowPool();       // never destroyed: its detached workers may outlive static destructors
    return *p;
}

template <typename Fn>
void parallel_rows(int64_t n, int threads, Fn&& fn) {
    pool().run(n, threads, std::function<void(int64_t, int64_t)>(fn));
}

bool gather_rows(const float* master, int64_t n_rows, int dim, const int64_t* ids, int64_t n, float* dst, int threads) {
    const size_t row_b = (size_t)dim * sizeof(float);
    std::atomic<bool> bad{false};
    parallel_rows(n, threads, [&](int64_t lo, int64_t hi) {
        for (int64_t i = lo; i < hi; ++i) {
            // a few rows ahead: the table is far larger than the caches, every row is a DRAM miss
            if (i + 8 < hi) __builtin_prefetch(master + ids[i + 8] * (int64_t)dim, 0, 0);
            const int64_t id = ids[i];
            if ((uint64_t)id >= (uint64_t)n_rows) { bad.store(true, std::memory_order_relaxed); continue; }
            memcpy(dst + i * (int64_t)dim, master + id * (int64_t)dim, row_b);
        }
    });
    return !bad.load();
}

bool scatter_rows(float* master, int64_t n_rows, int dim, const int64_t* ids, const uint8_t* primary, int64_t n,
                  const float* src, int average, int threads) {
    const size_t row_b = (size_t)dim * sizeof(float);
    std::atomic<bool> bad{false};
    parallel_rows(n, threads, [&](int64_t lo, int64_t hi) {
        for (int64_t i = lo; i < hi; ++i) {
            if (primary && !primary[i]) continue;
            const int64_t id = ids[i];
            if ((uint64_t)id >= (uint64_t)n_rows) { bad.store(true, std::memory_order_relaxed); continue; }
            float* w = master + id * (int64_t)dim;
            const float* r = src + i * (int64_t)dim;
            if (average) {
                for (int c = 0; c < dim; ++c) w[c] = (w[c] + r[c]) / 2;
            } else {
                memcpy(w, r, row_b);
            }
        }
    });
    return !bad.load();
}

// cudaMemcpyAsync in pieces: a copy engine does not preempt a copy, and the training step's own small copies (inputs
// in, loss out) would otherwise wait for a whole chunk
cudaError_t copy_pieces(void* dst, const void* src, int64_t bytes, cudaMemcpyKind kind, cudaStream_t s) {
    constexpr int64_t PIECE = 4 << 20;
    for (int64_t o = 0; o < bytes; o += PIECE) {
        const int64_t m = bytes - o < PIECE ? bytes - o : PIECE;
        cudaError_t e = cudaMemcpyAsync((char*)dst + o, (const char*)src + o, (size_t)m, kind, s);
        if (e != cudaSuccess) return e;
    }
    return cudaSuccess;
}

struct EventPair {
    cudaEvent_t ev[2] = {nullptr, nullptr};
    bool used[2] = {false, false};
    cudaError_t init() {
        for (int b = 0; b < 2; ++b) {
            cudaError_t e = cudaEventCreateWithFlags(&ev[b], cudaEventDisableTiming);
            if (e != cudaSuccess) return e;
        }
        return cudaSuccess;
    }
    ~EventPair() {
        for (int b = 0; b < 2; ++b)
            if (ev[b]) cudaEventDestroy(ev[b]);
    }
};

}  // namespace

// dst[i, :] = master[ids[i], :] for i in [0, n): ids ascending or not, all host pointers
extern "C" int cdlrm_host_gather_rows(const float* master, int64_t n_rows, int dim, const int64_t* ids, int64_t n,
                                      float* dst, int threads) {
    ARG_CHECK(master && ids && dst && dim > 0 && n >= 0 && n_rows > 0);
    if (!gather_rows(master, n_rows, dim, ids, n, dst, threads)) {
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
    if (!scatter_rows(master, n_rows, dim, ids, primary, n, src, average, threads)) {
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
    CU_CHECK(copy_pieces(dst, src, bytes, kind == 1 ? cudaMemcpyHostToDevice : cudaMemcpyDeviceToHost, (cudaStream_t)stream));
    return CDLRM_OK;
}

// ---- whole transfers in ONE call ------------------------------------------------------------------------------------
// A window's prefetch is ~1000 chunks and its write-back ~300.  Driven chunk by chunk from Python, every chunk took the
// interpreter lock three or four times in the planner's thread while the main thread -- busy enqueueing training steps --
// held it: the transfer crawled (12 GB/s) and the main thread lost 5 ms switch intervals.  Here the chunk loop is native:
// the caller drops the interpreter lock once for the whole transfer.
//
// Prefetch: for every job j (a table's id list) dst_j[i, :] = masters[j][ids_j[i], :].  Host threads gather chunk c into
// one of the two pinned staging chunks while the copy engine moves chunk c-1 into HBM.  ids_j: HOST int64, dst_j: DEVICE.
extern "C" int cdlrm_host_prefetch_rows(int device, int n_jobs, const float* const* masters, const int64_t* n_rows, int dim,
                                        const int64_t* const* ids, const int64_t* counts, float* const* dst,
                                        float* chunk0, float* chunk1, int64_t chunk_rows, int threads,
                                        cdlrm_stream stream) {
    ARG_CHECK(n_jobs >= 0 && dim > 0 && chunk0 && chunk1 && chunk_rows > 0);
    if (n_jobs == 0) return CDLRM_OK;
    ARG_CHECK(masters && n_rows && ids && counts && dst);
    CU_CHECK(cudaSetDevice(device));
    cudaStream_t s = (cudaStream_t)stream;
    EventPair ep;
    CU_CHECK(ep.init());
    float* chunk[2] = {chunk0, chunk1};
    int b = 0;
    bool ok = true;
    for (int j = 0; j < n_jobs && ok; ++j) {
        ARG_CHECK(counts[j] >= 0 && (counts[j] == 0 || (masters[j] && ids[j] && dst[j] && n_rows[j] > 0)));
        for (int64_t a = 0; a < counts[j]; a += chunk_rows, b ^= 1) {
            const int64_t m = counts[j] - a < chunk_rows ? counts[j] - a : chunk_rows;
            if (ep.used[b]) CU_CHECK(cudaEventSynchronize(ep.ev[b]));      // the copy that last read this chunk is done
            if (!gather_rows(masters[j], n_rows[j], dim, ids[j] + a, m, chunk[b], threads)) {
                ok = false;
                break;
            }
            CU_CHECK(copy_pieces(dst[j] + a * (int64_t)dim, chunk[b], m * (int64_t)dim * 4, cudaMemcpyHostToDevice, s));
            CU_CHECK(cudaEventRecord(ep.ev[b], s));
            ep.used[b] = true;
        }
    }
    for (int q = 0; q < 2; ++q)                 // the staging chunks are free again when this returns
        if (ep.used[q]) CU_CHECK(cudaEventSynchronize(ep.ev[q]));
    if (!ok) {
        cdlrm_set_error("cdlrm_host_prefetch_rows: id outside its table");
        return CDLRM_ERR_ARG;
    }
    return CDLRM_OK;
}

// Write-back: for every job j masters[j][ids_j[i], :] = src_j[i, :] (mean of the two with `average`) for the i whose
// primary_j[i] != 0.  The copy engine brings chunk c out of HBM while the host threads scatter chunk c-1 into the master.
// ids_j, primary_j: HOST; src_j: DEVICE rows.  Returns when every row is in the master.
extern "C" int cdlrm_host_writeback_rows(int device, int n_jobs, float* const* masters, const int64_t* n_rows, int dim,
                                         const int64_t* const* ids, const uint8_t* const* primary, const int64_t* counts,
                                         const float* const* src, float* chunk0, float* chunk1, int64_t chunk_rows,
                                         int average, int threads, cdlrm_stream stream) {
    ARG_CHECK(n_jobs >= 0 && dim > 0 && chunk0 && chunk1 && chunk_rows > 0);
    if (n_jobs == 0) return CDLRM_OK;
    ARG_CHECK(masters && n_rows && ids && counts && src);
    CU_CHECK(cudaSetDevice(device));
    cudaStream_t s = (cudaStream_t)stream;
    EventPair ep;
    CU_CHECK(ep.init());
    float* chunk[2] = {chunk0, chunk1};
    struct Piece { int j; int64_t a, m; int b; };
    Piece prev{-1, 0, 0, 0};
    bool ok = true;
    auto scatter = [&](const Piece& p) -> int {
        CU_CHECK(cudaEventSynchronize(ep.ev[p.b]));                        // the chunk has landed in host memory
        if (!scatter_rows(masters[p.j], n_rows[p.j], dim, ids[p.j] + p.a, primary && primary[p.j] ? primary[p.j] + p.a : nullptr,
                          p.m, chunk[p.b], average, threads))
            ok = false;
        return CDLRM_OK;
    };
    int b = 0;
    for (int j = 0; j < n_jobs; ++j) {
        ARG_CHECK(counts[j] >= 0 && (counts[j] == 0 || (masters[j] && ids[j] && src[j] && n_rows[j] > 0)));
        for (int64_t a = 0; a < counts[j]; a += chunk_rows, b ^= 1) {
            const int64_t m = counts[j] - a < chunk_rows ? counts[j] - a : chunk_rows;
            // chunk[b] was scattered before the previous copy was issued (two chunks alternate)
            CU_CHECK(copy_pieces(chunk[b], src[j] + a * (int64_t)dim, m * (int64_t)dim * 4, cudaMemcpyDeviceToHost, s));
            CU_CHECK(cudaEventRecord(ep.ev[b], s));
            if (prev.j >= 0) {
                const int rc = scatter(prev);
                if (rc != CDLRM_OK) return rc;
            }
            prev = Piece{j, a, m, b};
        }
    }
    if (prev.j >= 0) {
        const int rc = scatter(prev);
        if (rc != CDLRM_OK) return rc;
    }
    if (!ok) {
        cdlrm_set_error("cdlrm_host_writeback_rows: id outside its table");
        return CDLRM_ERR_ARG;
    }
    return CDLRM_OK;
}
