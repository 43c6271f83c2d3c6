// plan.cu -- look-ahead window planner: the torch.unique of
// Prefetcher.process_batch_slice (cache_manager.py:27-46) and the decision part of
// CacheEmbeddings (main_no_ddp.py:155-204), on the GPU, bit-exact.
//
// Tag evolution depends only on (index stream, RNG), never on embedding values, so
// the plan for window w+1 is computed on a side stream while window w trains.
//
// Phase A (per table): bitmap of the window's ids -> ascending unique list
// (sort-free, = torch.unique) -> probe against the plan tags with a ballot over the
// tag line, pin hit ways -> drop misses whose set is fully pinned -> stable
// compaction of the survivors (ascending id = the reference's row order).
// Host: R_k survivor rows -> draws q[R_k, ways] from the torch-compatible mt19937
// stream (cdlrm_rng_*), table 0..T-1, exactly as Categorical.sample() consumes it.
// Phase B (per table): way = argmax(probs/q) among un-pinned ways -> claims
// (last survivor wins a contested slot, = single-thread index_put_) -> evict list
// (old tag != -1, survivor order, duplicates kept) -> winners update the plan tags
// and form the fill list.
#include "common.cuh"
#include "compact.cuh"
#include "expdraw.cuh"

namespace {

constexpr int CNT_U = 0, CNT_HIT = 1, CNT_DROP = 2, CNT_ROWS = 3, CNT_E = 4, CNT_F = 5;

// ---- A5: probe unique ids against the plan tags, pin the hit ways ----------------------------
template <int GW>
__global__ void __launch_bounds__(256) plan_probe_kernel(const int64_t* __restrict__ uniq,
                                                         const unsigned long long* __restrict__ d_U,
                                                         const int64_t* __restrict__ tags, int64_t S, int ways,
                                                         unsigned long long* __restrict__ pin,
                                                         uint8_t* __restrict__ state,
                                                         unsigned long long* __restrict__ d_hits) {
    constexpr int GPW = 32 / GW, NG = 8 * GPW, IPG = 256 / NG;
    const int64_t U = (int64_t)*d_U;
    const int64_t cta0 = (int64_t)blockIdx.x * 256;
    if (cta0 >= U) return;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int gl = lane % GW, gidx = lane / GW, group = warp * GPW + gidx;
    const uint32_t gmask = GW == 32 ? 0xffffffffu : ((1u << GW) - 1u);
    int hits = 0;
    for (int i = 0; i < IPG; ++i) {
        const int64_t u = cta0 + group * IPG + i;
        const bool valid = u < U;
        const int64_t id = valid ? uniq[u] : 0;
        const int64_t s = set_index(id, S);
        int way = -1;
        for (int w0 = 0; w0 < ways; w0 += GW) {
            const int w = w0 + gl;
            const bool m = valid && w < ways && tags[s * ways + w] == id;
            const uint32_t gb = (__ballot_sync(0xffffffffu, m) >> (gidx * GW)) & gmask;
            if (gb && way < 0) way = w0 + __ffs(gb) - 1;
        }
        if (gl == 0 && valid) {
            if (way >= 0) {
                atomicOr(pin + s, 1ull << way);
                state[u] = (uint8_t)way;
                ++hits;
            } else {
                state[u] = 255;
            }
        }
    }
    __shared__ int s_hits;
    if (threadIdx.x == 0) s_hits = 0;
    __syncthreads();
    if (hits) atomicAdd(&s_hits, hits);
    __syncthreads();
    if (threadIdx.x == 0 && s_hits) atomicAdd(d_hits, (unsigned long long)s_hits);
}

__device__ __forceinline__ bool is_survivor(int64_t u, int64_t U, const int64_t* uniq, const uint8_t* state,
                                            const unsigned long long* pin, int64_t S, unsigned long long full,
                                            bool& dropped) {
    dropped = false;
    if (u >= U || state[u] != 255) return false;
    const int64_t s = set_index(uniq[u], S);
    if ((pin[s] & full) == full) {  // every way pinned by this window's hits (main_no_ddp.py:173-180)
        dropped = true;
        return false;
    }
    return true;
}

// ---- A6/A8: stable compaction of the survivors ---------------------------------------------------
template <bool EMIT>
__global__ void __launch_bounds__(256) surv_kernel(const int64_t* __restrict__ uniq,
                                                   const unsigned long long* __restrict__ d_U,
                                                   const uint8_t* __restrict__ state,
                                                   const unsigned long long* __restrict__ pin, int64_t S,
                                                   unsigned long long full, int32_t* __restrict__ blocksum,
                                                   int32_t* __restrict__ surv,
                                                   unsigned long long* __restrict__ d_drop) {
    __shared__ int s_w[33];
    __shared__ int s_drop;
    const int64_t U = (int64_t)*d_U;
    const int64_t u0 = (int64_t)blockIdx.x * TILE + threadIdx.x * 4;
    if (threadIdx.x == 0) s_drop = 0;
    bool sv[4];
    int c = 0, nd = 0;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        bool d;
        sv[q] = is_survivor(u0 + q, U, uniq, state, pin, S, full, d);
        c += sv[q];
        nd += d;
    }
    int total;
    const int ex = block_excl_scan<256>(c, s_w, total);
    if (!EMIT) {
        if (nd) atomicAdd(&s_drop, nd);
        __syncthreads();
        if (threadIdx.x == 0) {
            blocksum[blockIdx.x] = total;
            if (s_drop) atomicAdd(d_drop, (unsigned long long)s_drop);
        }
    } else {
        int off = blocksum[blockIdx.x] + ex;
#pragma unroll
        for (int q = 0; q < 4; ++q)
            if (sv[q]) surv[off++] = (int32_t)(u0 + q);
    }
}

// ---- B1: way = argmax(probs / q) among un-pinned ways; claim the slot -----------------------------
// RAW: q holds the raw mt19937 word pairs (uint2 per draw) of the device-resident stream and the
// exponential transform is applied here; otherwise q holds host-drawn float32 exponentials.
template <int GW, bool RAW>
__global__ void __launch_bounds__(256) select_kernel(const int64_t* __restrict__ uniq,
                                                     const int32_t* __restrict__ surv, int64_t R,
                                                     const void* __restrict__ q,
                                                     const unsigned long long* __restrict__ pin,
                                                     const int64_t* __restrict__ tags, int64_t S, int ways,
                                                     unsigned long long full, int32_t* __restrict__ slot_out,
                                                     int64_t* __restrict__ old_out, int32_t* __restrict__ claim) {
    constexpr int GPW = 32 / GW, NG = 8 * GPW;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int gl = lane % GW, gidx = lane / GW, group = warp * GPW + gidx;
    const int64_t r = (int64_t)blockIdx.x * NG + group;
    const bool valid = r < R;
    int64_t id = 0, s = 0;
    unsigned long long avail = full;
    if (valid) {
        id = uniq[surv[r]];
        s = set_index(id, S);
        avail = ~pin[s] & full;
    }
    // probs = avail / avail.sum(-1)  (Categorical normalisation), float32 IEEE division
    const float p = __fdiv_rn(1.0f, (float)__popcll(avail));
    float best = -1.0f;
    int bw = 0x7fffffff;
    for (int w = gl; w < ways; w += GW) {
        float qv = 1.0f;
        if (valid) {
            if (RAW) qv = exp_draw_from_raw(reinterpret_cast<const uint2*>(q)[r * ways + w]);
            else qv = reinterpret_cast<const float*>(q)[r * ways + w];
        }
        const float v = __fdiv_rn(((avail >> w) & 1ull) ? p : 0.0f, qv);
        if (v > best) {  // strict: first index wins ties (torch.argmax)
            best = v;
            bw = w;
        }
    }
#pragma unroll
    for (int o = GW / 2; o >= 1; o >>= 1) {
        const float ov = __shfl_xor_sync(0xffffffffu, best, o);
        const int ow = __shfl_xor_sync(0xffffffffu, bw, o);
        if (ov > best || (ov == best && ow < bw)) {
            best = ov;
            bw = ow;
        }
    }
    if (valid && gl == 0) {
        const int64_t slot = S * bw + s;
        slot_out[r] = (int32_t)slot;
        old_out[r] = tags[s * ways + bw];   // tag BEFORE this window's writes (main_no_ddp.py:190,196)
        atomicMax(claim + slot, (int32_t)r);  // last survivor wins (single-thread index_put_, :204)
    }
}

// ---- B2/B4: evict + fill lists -------------------------------------------------------------------------
template <bool EMIT>
__global__ void __launch_bounds__(256) lists_kernel(int64_t R, const int64_t* __restrict__ uniq,
                                                    const int32_t* __restrict__ surv,
                                                    const int32_t* __restrict__ slot, const int64_t* __restrict__ old,
                                                    int32_t* __restrict__ claim, uint8_t* __restrict__ flag,
                                                    uint8_t* __restrict__ state,
                                                    int32_t* __restrict__ bsE, int32_t* __restrict__ bsF,
                                                    int64_t* __restrict__ tags, int64_t S, int ways,
                                                    int64_t* __restrict__ evict_ids, int32_t* __restrict__ evict_slots,
                                                    uint8_t* __restrict__ evict_primary,
                                                    int64_t* __restrict__ fill_ids, int32_t* __restrict__ fill_slots,
                                                    bool primary_only) {
    __shared__ int s_w[33];
    const int64_t r0 = (int64_t)blockIdx.x * TILE + threadIdx.x * 4;
    uint8_t f[4];
    int ce = 0, cf = 0;
#pragma unroll
    for (int qq = 0; qq < 4; ++qq) {
        const int64_t r = r0 + qq;
        f[qq] = 0;
        if (r < R) {
            if (!EMIT) {
                const bool win = claim[slot[r]] == (int32_t)r;
                // (primary_only: a (set, way) claimed by several ids of the window is listed once, by its winner -- the
                // entry whose row the write-back uses; the reference lists it once per claimant, main_no_ddp.py:190-199)
                const bool ev = old[r] != -1 && (win || !primary_only);
                f[qq] = (uint8_t)((ev ? 1 : 0) | (win ? 2 : 0));
                flag[r] = f[qq];
            } else {
                f[qq] = flag[r];
            }
        }
        ce += f[qq] & 1;
        cf += (f[qq] >> 1) & 1;
    }
    int totE, totF;
    const int exE = block_excl_scan<256>(ce, s_w, totE);
    const int exF = block_excl_scan<256>(cf, s_w, totF);
    if (!EMIT) {
        if (threadIdx.x == 0) {
            bsE[blockIdx.x] = totE;
            bsF[blockIdx.x] = totF;
        }
        return;
    }
    int64_t oe = (int64_t)bsE[blockIdx.x] + exE, of = (int64_t)bsF[blockIdx.x] + exF;
#pragma unroll
    for (int qq = 0; qq < 4; ++qq) {
        const int64_t r = r0 + qq;
        if (r >= R) continue;
        const int32_t sl = slot[r];
        if (f[qq] & 1) {
            evict_ids[oe] = old[r];
            evict_slots[oe] = sl;
            evict_primary[oe] = (f[qq] >> 1) & 1;
            ++oe;
        }
        if (f[qq] & 2) {
            const int32_t u = surv[r];
            const int64_t id = uniq[u];
            state[u] = 254;   // cached after this install: not a loser
            fill_ids[of] = id;
            fill_slots[of] = sl;
            ++of;
            const int64_t way = sl / S, s = sl - way * S;
            tags[s * ways + way] = id;  // main_no_ddp.py:204
        }
        claim[sl] = -1;  // leave the claim array clean for the next table / window
    }
}

// ---- losers: unique ids of the window that stay un-cached (state 255), ascending ------------------
template <bool EMIT>
__global__ void __launch_bounds__(256) losers_kernel(const int64_t* __restrict__ uniq, int64_t U,
                                                     const uint8_t* __restrict__ state, const uint32_t* __restrict__ own,
                                                     int32_t* __restrict__ blocksum, int64_t* __restrict__ out) {
    __shared__ int s_w[33];
    const int64_t u0 = (int64_t)blockIdx.x * TILE + threadIdx.x * 4;
    bool lose[4];
    int c = 0;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        lose[q] = (u0 + q < U) && state[u0 + q] == 255;
        if (lose[q] && own) {          // data-parallel ranks: only the ids this rank's own batches contain
            const int64_t id = uniq[u0 + q];
            lose[q] = (own[id >> 5] >> (id & 31)) & 1u;
        }
        c += lose[q];
    }
    int total;
    const int ex = block_excl_scan<256>(c, s_w, total);
    if (!EMIT) {
        if (threadIdx.x == 0) blocksum[blockIdx.x] = total;
    } else {
        int64_t off = (int64_t)blocksum[blockIdx.x] + ex;
#pragma unroll
        for (int q = 0; q < 4; ++q)
            if (lose[q]) out[off++] = uniq[u0 + q];
    }
}

__global__ void fill_i32_kernel(int32_t* p, int64_t n, int32_t v) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (; i < n; i += stride) p[i] = v;
}

inline int pow2_ceil(int v) {
    int p = 1;
    while (p < v) p <<= 1;
    return p;
}

inline size_t align256(size_t v) { return (v + 255) & ~(size_t)255; }

struct Carve {
    char* base;
    size_t off = 0;
    template <typename Tp>
    Tp* take(size_t count) {
        Tp* p = base ? reinterpret_cast<Tp*>(base + off) : nullptr;
        off += align256(count * sizeof(Tp));
        return p;
    }
};

// One carve-up shared by the size query and the bind.
size_t carve_workspace(const cdlrm_ctx* c, int64_t N, char* base, std::vector<PlanTable>* pt,
                       std::vector<unsigned long long*>* pins, cdlrm_ctx* out) {
    Carve cv{base};
    int64_t umax_max = 0, sets_max = 0, rows_max = 0, words_max = 0;
    for (int k = 0; k < c->T; ++k) {
        const TableDesc& t = c->tabs[k];
        const int64_t umax = t.n_rows < N ? t.n_rows : N;
        const int64_t words = (t.n_rows + 31) / 32;
        PlanTable p;
        p.bitmap = cv.take<uint32_t>(words);
        p.own = cv.take<uint32_t>(words);
        p.uniq = cv.take<int64_t>(umax);
        p.surv = cv.take<int32_t>(umax);
        p.state = cv.take<uint8_t>(umax);
        p.umax = umax;
        unsigned long long* pin = cv.take<unsigned long long>(t.num_sets);
        if (pt) pt->push_back(p);
        if (pins) pins->push_back(pin);
        umax_max = umax > umax_max ? umax : umax_max;
        sets_max = t.num_sets > sets_max ? t.num_sets : sets_max;
        rows_max = t.cache_rows > rows_max ? t.cache_rows : rows_max;
        words_max = words > words_max ? words : words_max;
    }
    const int64_t items_max = umax_max > words_max ? umax_max : words_max;
    const int64_t nblk = (items_max + TILE - 1) / TILE + 1;
    int32_t* bs1 = cv.take<int32_t>(nblk);
    int32_t* bs2 = cv.take<int32_t>(nblk);
    int32_t* claim = cv.take<int32_t>(rows_max);
    int32_t* slot = cv.take<int32_t>(umax_max);
    int64_t* old = cv.take<int64_t>(umax_max);
    uint8_t* flag = cv.take<uint8_t>(umax_max);
    if (out) {
        out->p_blocksum = bs1;
        out->p_claim = claim;
        out->p_slot = slot;
        out->p_old = old;
        out->p_flag = flag;
        out->p_blocksum2 = bs2;
        out->umax_max = umax_max;
        out->sets_max = sets_max;
        out->rows_max = rows_max;
    }
    return cv.off;
}

__global__ void set_u64_kernel(unsigned long long* p, unsigned long long v) { *p = v; }

}  // namespace

extern "C" int64_t cdlrm_plan_workspace_bytes(const cdlrm_ctx* c, int64_t window_len) {
    if (!c || window_len <= 0) return -1;
    return (int64_t)carve_workspace(c, window_len, nullptr, nullptr, nullptr, nullptr);
}

extern "C" int cdlrm_plan_bind_workspace(cdlrm_ctx* c, void* ws, int64_t bytes, int64_t window_len) {
    ARG_CHECK(c && ws && window_len > 0);
    ARG_CHECK(((uintptr_t)ws & 255) == 0);
    const int64_t need = cdlrm_plan_workspace_bytes(c, window_len);
    if (bytes < need) {
        cdlrm_set_error("planner workspace too small: %lld < %lld bytes", (long long)bytes, (long long)need);
        return CDLRM_ERR_ARG;
    }
    CU_CHECK(cudaSetDevice(c->device));
    c->ptabs.clear();
    std::vector<unsigned long long*>& pins = c->pins;
    pins.clear();
    carve_workspace(c, window_len, (char*)ws, &c->ptabs, &pins, c);
    c->plan_window_len = window_len;
    c->plan_ws = (char*)ws;
    // bitmaps start clean (the emit kernel keeps them clean); claims start at -1
    CU_CHECK(cudaDeviceSynchronize());
    for (int k = 0; k < c->T; ++k) {
        CU_CHECK(cudaMemset(c->ptabs[k].bitmap, 0, sizeof(uint32_t) * ((c->tabs[k].n_rows + 31) / 32)));
        CU_CHECK(cudaMemset(c->ptabs[k].own, 0, sizeof(uint32_t) * ((c->tabs[k].n_rows + 31) / 32)));
    }
    c->own_marked = false;
    fill_i32_kernel<<<1184, 256>>>(c->p_claim, c->rows_max, -1);
    CU_CHECK(cudaGetLastError());
    CU_CHECK(cudaDeviceSynchronize());
    return CDLRM_OK;
}

extern "C" const int64_t* cdlrm_plan_unique_ptr(const cdlrm_ctx* c, int table) {
    if (!c || table < 0 || table >= c->T || c->ptabs.empty()) return nullptr;
    return c->ptabs[table].uniq;
}

extern "C" int cdlrm_plan_copy_unique(cdlrm_ctx* c, int table, int64_t* out, int64_t n, cdlrm_stream stream) {
    ARG_CHECK(c && table >= 0 && table < c->T && n >= 0);
    if (c->ptabs.empty()) {
        cdlrm_set_error("planner workspace not bound");
        return CDLRM_ERR_STATE;
    }
    ARG_CHECK(n <= c->ptabs[table].umax);
    if (n == 0) return CDLRM_OK;
    ARG_CHECK(out);
    CU_CHECK(cudaSetDevice(c->device));
    CU_CHECK(cudaMemcpyAsync(out, c->ptabs[table].uniq, sizeof(int64_t) * n, cudaMemcpyDeviceToDevice, (cudaStream_t)stream));
    return CDLRM_OK;
}

static int plan_impl(cdlrm_ctx* c, const int64_t* win_ids, int64_t ld, int64_t n, const int64_t* h_uniq_len,
                     int64_t* h_counts, bool unique_only, cdlrm_stream stream);

extern "C" int cdlrm_plan_unique(cdlrm_ctx* c, const int64_t* win_ids, int64_t ld, int64_t n, int64_t* h_counts,
                                 cdlrm_stream stream) {
    return plan_impl(c, win_ids, ld, n, nullptr, h_counts, true, stream);
}

extern "C" int cdlrm_plan_phase_a(cdlrm_ctx* c, const int64_t* win_ids, int64_t ld, int64_t n,
                                  const int64_t* h_uniq_len, int64_t* h_counts, cdlrm_stream stream) {
    return plan_impl(c, win_ids, ld, n, h_uniq_len, h_counts, false, stream);
}

extern "C" int cdlrm_plan_mark_ids(cdlrm_ctx* c, const int64_t* ids, int64_t ld, int64_t n, cdlrm_stream stream) {
    ARG_CHECK(c && n >= 0);
    if (n == 0) return CDLRM_OK;
    ARG_CHECK(ids);
    if (c->ptabs.empty()) {
        cdlrm_set_error("planner workspace not bound");
        return CDLRM_ERR_STATE;
    }
    cudaStream_t s = (cudaStream_t)stream;
    CU_CHECK(cudaSetDevice(c->device));
    for (int k = 0; k < c->T; ++k) {
        const int g1 = (int)((n + 1023) / 1024 < 148 * 16 ? (n + 1023) / 1024 : 148 * 16);
        LAUNCH(K_PLAN_BITMAP_SET, s, bitmap_set_kernel<<<g1 > 0 ? g1 : 1, 256, 0, s>>>(ids + k * ld, n, c->ptabs[k].bitmap, c->tabs[k].n_rows, c->d_flags));
    }
    CU_CHECK(cudaGetLastError());
    return CDLRM_OK;
}

// Eviction lists of the next plans: every claimant of a replaced (set, way) (0: the reference's lists, main_no_ddp.py:190-199,
// needed by callers that return its eviction_data) or only the winner of each (1: what the write-back needs -- under cache
// pressure the full list is several times longer than the fill list: 27 M entries for 6.8 M fills at 2 GPUs).
extern "C" int cdlrm_plan_set_primary_evictions(cdlrm_ctx* c, int on) {
    ARG_CHECK(c);
    c->primary_evictions = on != 0;
    return CDLRM_OK;
}

// Own-id bitmaps: the ids of the window that THIS rank's own batches contain (a data-parallel rank trains on its slice of
// every global batch).  With them marked, cdlrm_plan_losers lists only the un-cached ids this rank itself will look up:
// a rank-private loser store of the size of a one-GPU run instead of the union over all ranks (42 GB at 8 GPUs).
extern "C" int cdlrm_plan_mark_own_ids(cdlrm_ctx* c, const int64_t* ids, int64_t ld, int64_t n, cdlrm_stream stream) {
    ARG_CHECK(c && n >= 0);
    if (n == 0) return CDLRM_OK;
    ARG_CHECK(ids);
    if (c->ptabs.empty()) {
        cdlrm_set_error("planner workspace not bound");
        return CDLRM_ERR_STATE;
    }
    cudaStream_t s = (cudaStream_t)stream;
    CU_CHECK(cudaSetDevice(c->device));
    for (int k = 0; k < c->T; ++k) {
        const int g1 = (int)((n + 1023) / 1024 < 148 * 16 ? (n + 1023) / 1024 : 148 * 16);
        LAUNCH(K_PLAN_BITMAP_SET, s, bitmap_set_kernel<<<g1 > 0 ? g1 : 1, 256, 0, s>>>(ids + k * ld, n, c->ptabs[k].own, c->tabs[k].n_rows, c->d_flags));
    }
    CU_CHECK(cudaGetLastError());
    c->own_marked = true;
    return CDLRM_OK;
}

// Window scan sharded over the ranks of a node: every rank marks its own share of the window's steps, then ORs the
// id bitmaps of all the others into its own, reading them in place over NVLink (peer workspaces mapped with
// cdlrm_peer_open: same carve-up on every rank, so a table's bitmap sits at the same offset everywhere).
namespace {
struct PeerWs { const char* p[CDLRM_MAX_PEERS]; };
__global__ void __launch_bounds__(256) or_peer_bitmaps_kernel(uint32_t* __restrict__ mine, int64_t off, PeerWs peers, int world,
                                                              int rank, int64_t words) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < words; i += stride) {
        uint32_t v = mine[i];
        for (int r = 0; r < world; ++r)
            if (r != rank) v |= __ldcv(reinterpret_cast<const uint32_t*>(peers.p[r] + off) + i);
        mine[i] = v;
    }
}
}  // namespace

extern "C" int cdlrm_plan_or_peer_bitmaps(cdlrm_ctx* c, const void* const* h_peer_ws, int world, int rank,
                                          cdlrm_stream stream) {
    ARG_CHECK(c && h_peer_ws && world >= 1 && world <= CDLRM_MAX_PEERS && rank >= 0 && rank < world);
    if (c->ptabs.empty() || !c->plan_ws) {
        cdlrm_set_error("planner workspace not bound");
        return CDLRM_ERR_STATE;
    }
    cudaStream_t s = (cudaStream_t)stream;
    CU_CHECK(cudaSetDevice(c->device));
    PeerWs pw;
    for (int r = 0; r < CDLRM_MAX_PEERS; ++r) pw.p[r] = nullptr;
    for (int r = 0; r < world; ++r) {
        ARG_CHECK(r == rank || h_peer_ws[r]);
        pw.p[r] = (const char*)h_peer_ws[r];
    }
    for (int k = 0; k < c->T; ++k) {
        const int64_t words = (c->tabs[k].n_rows + 31) / 32;
        const int64_t off = (char*)c->ptabs[k].bitmap - c->plan_ws;
        int64_t blocks = (words + 255) / 256;
        if (blocks > 148 * 4) blocks = 148 * 4;
        LAUNCH(K_PLAN_BITMAP_SET, s, or_peer_bitmaps_kernel<<<(int)blocks, 256, 0, s>>>(c->ptabs[k].bitmap, off, pw, world, rank, words));
    }
    CU_CHECK(cudaGetLastError());
    return CDLRM_OK;
}

static int plan_impl(cdlrm_ctx* c, const int64_t* win_ids, int64_t ld, int64_t n, const int64_t* h_uniq_len,
                     int64_t* h_counts, bool unique_only, cdlrm_stream stream) {
    ARG_CHECK(c && h_counts);
    ARG_CHECK(win_ids || !h_uniq_len);       // win_ids == NULL: the bitmaps were filled by cdlrm_plan_mark_ids
    ARG_CHECK(n >= 0);
    if (c->ptabs.empty()) {
        cdlrm_set_error("planner workspace not bound");
        return CDLRM_ERR_STATE;
    }
    ARG_CHECK(n <= c->plan_window_len);
    cudaStream_t s = (cudaStream_t)stream;
    CU_CHECK(cudaSetDevice(c->device));
    if (!unique_only) {
        for (int k = 0; k < c->T; ++k)
            if (!c->tabs[k].plan_tags) {
                cdlrm_set_error("table %d: tags not bound", k);
                return CDLRM_ERR_STATE;
            }
        int rc = cdlrm_sync_tabs(c, s);
        if (rc) return rc;
    }
    std::vector<unsigned long long*>& pins = c->pins;
    unsigned long long* cnt = reinterpret_cast<unsigned long long*>(c->p_counts);
    CU_CHECK(cudaMemsetAsync(cnt, 0, sizeof(int64_t) * c->T * 8, s));
    const int gw = pow2_ceil(c->ways) > 32 ? 32 : pow2_ceil(c->ways);
    const unsigned long long full = c->ways == 64 ? ~0ull : ((1ull << c->ways) - 1ull);
    for (int k = 0; k < c->T; ++k) {
        const TableDesc& t = c->tabs[k];
        const PlanTable& p = c->ptabs[k];
        unsigned long long* ck = cnt + k * 8;
        int64_t ubound;  // host-side upper bound of U_k (U_k itself stays on the device)
        if (h_uniq_len) {
            ARG_CHECK(h_uniq_len[k] >= 0 && h_uniq_len[k] <= p.umax);
            ubound = h_uniq_len[k];
            if (ubound)
                CU_CHECK(cudaMemcpyAsync(p.uniq, win_ids + k * ld, sizeof(int64_t) * ubound, cudaMemcpyDeviceToDevice, s));
            LAUNCH(K_MISC, s, set_u64_kernel<<<1, 1, 0, s>>>(ck + CNT_U, (unsigned long long)ubound));
        } else {
            ubound = p.umax < n ? p.umax : n;
            if (n > 0) {
                const int64_t nwords = (t.n_rows + 31) / 32;
                int g1 = (int)((n + 1023) / 1024 < 148 * 16 ? (n + 1023) / 1024 : 148 * 16);
                if (win_ids) LAUNCH(K_PLAN_BITMAP_SET, s, bitmap_set_kernel<<<g1 > 0 ? g1 : 1, 256, 0, s>>>(win_ids + k * ld, n, p.bitmap, t.n_rows, c->d_flags));
                const int nblk = (int)((nwords + TILE - 1) / TILE);
                LAUNCH(K_PLAN_COMPACT, s, bitmap_count_kernel<<<nblk, 256, 0, s>>>(p.bitmap, nwords, c->p_blocksum));
                LAUNCH(K_PLAN_COMPACT, s, scan_tiles_kernel<<<1, 1024, 0, s>>>(c->p_blocksum, nblk, ck + CNT_U));
                LAUNCH(K_PLAN_COMPACT, s, (bitmap_emit_kernel<int64_t, true><<<nblk, 256, 0, s>>>(p.bitmap, nwords, c->p_blocksum, p.uniq)));
            }
        }
        if (unique_only) continue;
        CU_CHECK(cudaMemsetAsync(pins[k], 0, sizeof(unsigned long long) * t.num_sets, s));
        if (ubound > 0) {
            const int g5 = (int)((ubound + 255) / 256);
#define LAUNCH_PP(GW) LAUNCH(K_PLAN_PROBE, s, plan_probe_kernel<GW><<<g5, 256, 0, s>>>(p.uniq, ck + CNT_U, t.plan_tags, t.num_sets, c->ways, pins[k], p.state, ck + CNT_HIT))
            switch (gw) {
                case 1: LAUNCH_PP(1); break;
                case 2: LAUNCH_PP(2); break;
                case 4: LAUNCH_PP(4); break;
                case 8: LAUNCH_PP(8); break;
                case 16: LAUNCH_PP(16); break;
                default: LAUNCH_PP(32); break;
            }
#undef LAUNCH_PP
            const int nblk = (int)((ubound + TILE - 1) / TILE);
            LAUNCH(K_PLAN_SURV, s, surv_kernel<false><<<nblk, 256, 0, s>>>(p.uniq, ck + CNT_U, p.state, pins[k], t.num_sets, full, c->p_blocksum, p.surv, ck + CNT_DROP));
            LAUNCH(K_PLAN_COMPACT, s, scan_tiles_kernel<<<1, 1024, 0, s>>>(c->p_blocksum, nblk, ck + CNT_ROWS));
            LAUNCH(K_PLAN_SURV, s, surv_kernel<true><<<nblk, 256, 0, s>>>(p.uniq, ck + CNT_U, p.state, pins[k], t.num_sets, full, c->p_blocksum, p.surv, ck + CNT_DROP));
        }
        CU_CHECK(cudaGetLastError());
    }
    // counts -> pinned host: [k*4 + {U, hits, dropped, rows}]
    for (int k = 0; k < c->T; ++k)
        CU_CHECK(cudaMemcpyAsync(h_counts + k * 4, c->p_counts + k * 8, sizeof(int64_t) * 4, cudaMemcpyDeviceToHost, s));
    return CDLRM_OK;
}

struct cdlrm_rngdev;
extern "C" int cdlrm_rngdev_raw(cdlrm_rngdev* r, uint32_t* d_out, int64_t n_draws, cdlrm_stream stream);

static int phase_b_impl(cdlrm_ctx* c, const float* q, cdlrm_rngdev* rng, uint32_t* raw, int64_t raw_draws,
                        const int64_t* h_rows, int64_t* evict_ids,
                                  int32_t* evict_slots, uint8_t* evict_primary, int64_t* fill_ids,
                                  int32_t* fill_slots, int64_t* h_counts2, cdlrm_stream stream) {
    ARG_CHECK(c && h_rows && h_counts2);
    if (c->ptabs.empty()) {
        cdlrm_set_error("planner workspace not bound");
        return CDLRM_ERR_STATE;
    }
    cudaStream_t s = (cudaStream_t)stream;
    CU_CHECK(cudaSetDevice(c->device));
    std::vector<unsigned long long*>& pins = c->pins;
    unsigned long long* cnt = reinterpret_cast<unsigned long long*>(c->p_counts);
    const int gw = pow2_ceil(c->ways) > 32 ? 32 : pow2_ceil(c->ways);
    const unsigned long long full = c->ways == 64 ? ~0ull : ((1ull << c->ways) - 1ull);
    int32_t* bsE = c->p_blocksum;
    int32_t* bsF = c->p_blocksum2;
    int64_t off = 0;
    for (int k = 0; k < c->T; ++k) {
        const TableDesc& t = c->tabs[k];
        const PlanTable& p = c->ptabs[k];
        const int64_t R = h_rows[k];
        ARG_CHECK(R >= 0 && R <= p.umax);
        unsigned long long* ck = cnt + k * 8;
        if (R > 0) {
            ARG_CHECK((q || rng) && evict_ids && evict_slots && evict_primary && fill_ids && fill_slots);
            const int NG = 256 / gw;
            const int g1 = (int)((R + NG - 1) / NG);
            const void* qk = q ? (const void*)(q + off * c->ways) : (const void*)raw;
            if (!q) {   // device stream: the draws of table k, in table order (split-invariant stream)
                ARG_CHECK(raw && R * c->ways <= raw_draws);
                int rc = cdlrm_rngdev_raw(rng, raw, R * c->ways, stream);
                if (rc) return rc;
            }
#define LAUNCH_SEL(GW) do { if (q) LAUNCH(K_PLAN_SELECT, s, (select_kernel<GW, false><<<g1, 256, 0, s>>>(p.uniq, p.surv, R, qk, pins[k], t.plan_tags, t.num_sets, c->ways, full, c->p_slot, c->p_old, c->p_claim))); else LAUNCH(K_PLAN_SELECT, s, (select_kernel<GW, true><<<g1, 256, 0, s>>>(p.uniq, p.surv, R, qk, pins[k], t.plan_tags, t.num_sets, c->ways, full, c->p_slot, c->p_old, c->p_claim))); } while (0)
            switch (gw) {
                case 1: LAUNCH_SEL(1); break;
                case 2: LAUNCH_SEL(2); break;
                case 4: LAUNCH_SEL(4); break;
                case 8: LAUNCH_SEL(8); break;
                case 16: LAUNCH_SEL(16); break;
                default: LAUNCH_SEL(32); break;
            }
#undef LAUNCH_SEL
            const int nblk = (int)((R + TILE - 1) / TILE);
            LAUNCH(K_PLAN_LISTS, s, lists_kernel<false><<<nblk, 256, 0, s>>>(R, p.uniq, p.surv, c->p_slot, c->p_old, c->p_claim, c->p_flag, p.state, bsE, bsF, t.plan_tags, t.num_sets, c->ways, nullptr, nullptr, nullptr, nullptr, nullptr, c->primary_evictions));
            LAUNCH(K_PLAN_COMPACT, s, scan_tiles_kernel<<<1, 1024, 0, s>>>(bsE, nblk, ck + CNT_E));
            LAUNCH(K_PLAN_COMPACT, s, scan_tiles_kernel<<<1, 1024, 0, s>>>(bsF, nblk, ck + CNT_F));
            LAUNCH(K_PLAN_LISTS, s, lists_kernel<true><<<nblk, 256, 0, s>>>(R, p.uniq, p.surv, c->p_slot, c->p_old, c->p_claim, c->p_flag, p.state, bsE, bsF, t.plan_tags, t.num_sets, c->ways, evict_ids + off, evict_slots + off, evict_primary + off, fill_ids + off, fill_slots + off, c->primary_evictions));
            CU_CHECK(cudaGetLastError());
        }
        off += R;
    }
    for (int k = 0; k < c->T; ++k)
        CU_CHECK(cudaMemcpyAsync(h_counts2 + k * 2, c->p_counts + k * 8 + CNT_E, sizeof(int64_t) * 2, cudaMemcpyDeviceToHost, s));
    return CDLRM_OK;
}

extern "C" int cdlrm_plan_phase_b(cdlrm_ctx* c, const float* q, const int64_t* h_rows, int64_t* evict_ids,
                                  int32_t* evict_slots, uint8_t* evict_primary, int64_t* fill_ids,
                                  int32_t* fill_slots, int64_t* h_counts2, cdlrm_stream stream) {
    int64_t total = 0;
    if (c && h_rows)
        for (int k = 0; k < c->T; ++k) total += h_rows[k];
    ARG_CHECK(q || total == 0);
    return phase_b_impl(c, q, nullptr, nullptr, 0, h_rows, evict_ids, evict_slots, evict_primary, fill_ids, fill_slots,
                        h_counts2, stream);
}

extern "C" int cdlrm_plan_phase_b_dev(cdlrm_ctx* c, cdlrm_rngdev* rng, uint32_t* raw_scratch, int64_t raw_draws,
                                      const int64_t* h_rows, int64_t* evict_ids, int32_t* evict_slots,
                                      uint8_t* evict_primary, int64_t* fill_ids, int32_t* fill_slots,
                                      int64_t* h_counts2, cdlrm_stream stream) {
    ARG_CHECK(rng);
    return phase_b_impl(c, nullptr, rng, raw_scratch, raw_draws, h_rows, evict_ids, evict_slots, evict_primary,
                        fill_ids, fill_slots, h_counts2, stream);
}

// Ascending list of the window's un-cached ids per table (call after phase B on the same
// stream).  h_uniq[k] = unique count of phase A; the list of table k is written at
// loser_ids + h_off[k] (capacity dropped_k + rows_k); h_counts3[k] = its length.
extern "C" int cdlrm_plan_losers(cdlrm_ctx* c, const int64_t* h_uniq, const int64_t* h_off, int64_t* loser_ids,
                                 int64_t* h_counts3, cdlrm_stream stream) {
    ARG_CHECK(c && h_uniq && h_off && h_counts3);
    if (c->ptabs.empty()) {
        cdlrm_set_error("planner workspace not bound");
        return CDLRM_ERR_STATE;
    }
    cudaStream_t s = (cudaStream_t)stream;
    CU_CHECK(cudaSetDevice(c->device));
    unsigned long long* cnt = reinterpret_cast<unsigned long long*>(c->p_counts);
    for (int k = 0; k < c->T; ++k) {
        const PlanTable& p = c->ptabs[k];
        const int64_t U = h_uniq[k];
        ARG_CHECK(U >= 0 && U <= p.umax);
        unsigned long long* ck = cnt + k * 8 + 6;
        if (U == 0) {
            CU_CHECK(cudaMemsetAsync(ck, 0, sizeof(unsigned long long), s));
            continue;
        }
        ARG_CHECK(loser_ids);
        const int nblk = (int)((U + TILE - 1) / TILE);
        const uint32_t* own = c->own_marked ? p.own : nullptr;
        LAUNCH(K_PLAN_LISTS, s, losers_kernel<false><<<nblk, 256, 0, s>>>(p.uniq, U, p.state, own, c->p_blocksum, nullptr));
        LAUNCH(K_PLAN_COMPACT, s, scan_tiles_kernel<<<1, 1024, 0, s>>>(c->p_blocksum, nblk, ck));
        LAUNCH(K_PLAN_LISTS, s, losers_kernel<true><<<nblk, 256, 0, s>>>(p.uniq, U, p.state, own, c->p_blocksum, loser_ids + h_off[k]));
    }
    if (c->own_marked) {        // the own-id bitmaps are clean again for the next window
        for (int k = 0; k < c->T; ++k)
            CU_CHECK(cudaMemsetAsync(c->ptabs[k].own, 0, sizeof(uint32_t) * ((c->tabs[k].n_rows + 31) / 32), s));
        c->own_marked = false;
    }
    CU_CHECK(cudaGetLastError());
    for (int k = 0; k < c->T; ++k)
        CU_CHECK(cudaMemcpyAsync(h_counts3 + k, c->p_counts + k * 8 + 6, sizeof(int64_t), cudaMemcpyDeviceToHost, s));
    return CDLRM_OK;
}
