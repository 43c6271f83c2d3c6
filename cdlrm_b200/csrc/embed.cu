// embed.cu -- per-step cache path: probe + slot + gather + sum-pool forward
// (Embedding_Table_Cache_Group.forward, model_no_ddp.py:149-212) and the
// de-duplicated sparse-SGD backward (EmbeddingBag backward + optimizer_embeds.step(),
// main_no_ddp.py:376,409,413).  All tables of a call are covered by one launch per
// stage (grid.y = table).  HBM-bound integer/row-copy work: no tensor cores.
#include "common.cuh"

namespace {

constexpr int FWD_CHUNK = 256;  // ids per CTA in the probe; rows per CTA in the gather
constexpr int CH = 8;           // max contributions merged by one chunk in the backward

template <int VEC> struct VecT;
template <> struct VecT<4> { using type = float4; };
template <> struct VecT<2> { using type = float2; };
template <> struct VecT<1> { using type = float; };

__device__ __forceinline__ float4 vadd(float4 a, float4 b) { return make_float4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w); }
__device__ __forceinline__ float2 vadd(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ float vadd(float a, float b) { return a + b; }
__device__ __forceinline__ float4 vfma(float s, float4 a, float4 b) { return make_float4(fmaf(s, a.x, b.x), fmaf(s, a.y, b.y), fmaf(s, a.z, b.z), fmaf(s, a.w, b.w)); }
__device__ __forceinline__ float2 vfma(float s, float2 a, float2 b) { return make_float2(fmaf(s, a.x, b.x), fmaf(s, a.y, b.y)); }
__device__ __forceinline__ float vfma(float s, float a, float b) { return fmaf(s, a, b); }
__device__ __forceinline__ float4 vscale(float s, float4 a) { return make_float4(s * a.x, s * a.y, s * a.z, s * a.w); }
__device__ __forceinline__ float2 vscale(float s, float2 a) { return make_float2(s * a.x, s * a.y); }
__device__ __forceinline__ float vscale(float s, float a) { return s * a; }
__device__ __forceinline__ void vzero(float4& a) { a = make_float4(0.f, 0.f, 0.f, 0.f); }
__device__ __forceinline__ void vzero(float2& a) { a = make_float2(0.f, 0.f); }
__device__ __forceinline__ void vzero(float& a) { a = 0.f; }

__device__ __forceinline__ void red_add(float4* p, float4 v) {
    // sm_90+ vectorised reduction: one 16-byte red per lane
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
__device__ __forceinline__ void red_add(float2* p, float2 v) {
    asm volatile("red.global.add.v2.f32 [%0], {%1, %2};" ::"l"(p), "f"(v.x), "f"(v.y) : "memory");
}
__device__ __forceinline__ void red_add(float* p, float v) { atomicAdd(p, v); }

__device__ __forceinline__ int warp_sum(int v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// ------------------------------------------------------------------------------
// K1: probe.  A group of GW lanes (GW = pow2 >= ways, <= 32) probes one id: lane w
// loads tag w of the set (one coalesced 8*ways-byte line), the hit way comes from a
// warp ballot.  Hits get their final slot; misses get -(1 + ordinal inside this
// chunk) and the chunk's miss count is published for K2's cross-chunk prefix.
// model_no_ddp.py:166-174.
// ------------------------------------------------------------------------------
template <int GW>
struct ProbeCfg { static constexpr int NT = GW >= 4 ? 1024 : 256; };   // 4 ids per group -> one batch in flight

template <int GW>
__global__ void __launch_bounds__(ProbeCfg<GW>::NT) probe_kernel(const TableDesc* __restrict__ tabs, int tb,
                                                    const int64_t* __restrict__ ids, int64_t ld_ids, int n_idx,
                                                    int32_t* __restrict__ slots, int64_t ld_slots,
                                                    int32_t* __restrict__ miss_cnt, int chunks, int ways) {
    constexpr int GPW = 32 / GW;                          // groups per warp
    constexpr int NG = ProbeCfg<GW>::NT / 32 * GPW;    // groups per CTA
    constexpr int IPG = FWD_CHUNK / NG;                   // ids per group
    static_assert(IPG >= 1 && IPG <= 32, "miss mask is 32 bits");
    const int t = blockIdx.y, chunk = blockIdx.x;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int gl = lane % GW, gidx = lane / GW, group = warp * GPW + gidx;
    const uint32_t gmask = GW == 32 ? 0xffffffffu : ((1u << GW) - 1u);
    const TableDesc& T = tabs[tb + t];
    const int64_t S = T.num_sets;
    const int64_t* __restrict__ tags = T.tags;
    const int64_t* tid = ids + (int64_t)t * ld_ids;
    int32_t* tsl = slots + (int64_t)t * ld_slots;
    const int base = chunk * FWD_CHUNK + group * IPG;
    uint32_t missmask = 0;

    constexpr int U = 4;  // ids in flight per group
#pragma unroll 1
    for (int i0 = 0; i0 < IPG; i0 += U) {
        int64_t id[U], s[U], tag[U];
        bool valid[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            int j = base + i0 + u;
            valid[u] = (i0 + u) < IPG && j < n_idx;
            id[u] = valid[u] ? __ldg(tid + j) : 0;
            s[u] = set_index(id[u], S);
        }
        int way[U];
#pragma unroll
        for (int u = 0; u < U; ++u) way[u] = -1;
        for (int w0 = 0; w0 < ways; w0 += GW) {
            const int w = w0 + gl;
#pragma unroll
            for (int u = 0; u < U; ++u) tag[u] = (valid[u] && w < ways) ? tags[s[u] * ways + w] : 0;
#pragma unroll
            for (int u = 0; u < U; ++u) {
                bool m = valid[u] && w < ways && tag[u] == id[u];
                uint32_t gb = (__ballot_sync(0xffffffffu, m) >> (gidx * GW)) & gmask;
                if (gb && way[u] < 0) way[u] = w0 + __ffs(gb) - 1;
            }
        }
        if (gl == 0) {
#pragma unroll
            for (int u = 0; u < U; ++u) {
                if (!valid[u]) continue;
                if (way[u] >= 0) tsl[base + i0 + u] = (int32_t)(S * way[u] + s[u]);
                else missmask |= 1u << (i0 + u);
            }
        }
    }
    __shared__ int s_cnt[NG];
    if (gl == 0) s_cnt[group] = __popc(missmask);
    __syncthreads();
    if (gl == 0 && missmask) {
        int b = 0;
        for (int g = 0; g < group; ++g) b += s_cnt[g];
        while (missmask) {
            int i = __ffs(missmask) - 1;
            missmask &= missmask - 1;
            tsl[base + i] = -(1 + b++);
        }
    }
    if (threadIdx.x == 0) {
        int tot = 0;
        for (int g = 0; g < NG; ++g) tot += s_cnt[g];
        miss_cnt[t * chunks + chunk] = tot;
    }
}

// ------------------------------------------------------------------------------
// K2: resolve misses + gather (+ copy-out when every bag holds exactly one id).
// A group of G lanes moves one row with 16-byte accesses (G*VEC*4 bytes per pass,
// 512 B for dim 128); U rows are in flight per group.  Misses take aux slot
// num_sets*ways + ordinal (batch order, model_no_ddp.py:177), read the master row
// (zero-copy from pinned host memory) and park it in the aux row (:179).
// POOL_P1: also write out[j] = row (EmbeddingBag sum with one id per bag, :202).
// ------------------------------------------------------------------------------
template <int VEC, bool POOL_P1>
__global__ void __launch_bounds__(256) gather_kernel(const TableDesc* __restrict__ tabs, int tb,
                                                     const int64_t* __restrict__ ids, int64_t ld_ids, int n_idx,
                                                     int32_t* __restrict__ slots, int64_t ld_slots,
                                                     const int32_t* __restrict__ miss_cnt, int chunks,
                                                     float* __restrict__ out, int64_t ld_out,
                                                     int32_t* __restrict__ n_miss, uint32_t* __restrict__ flags,
                                                     int dim, int ways, int64_t aux_rows, int G) {
    using V = typename VecT<VEC>::type;
    const int t = blockIdx.y, chunk = blockIdx.x;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    __shared__ int s_prefix;
    if (warp == 0) {
        int acc = 0;
        for (int c = lane; c < chunk; c += 32) acc += miss_cnt[t * chunks + c];
        acc = warp_sum(acc);
        if (lane == 0) {
            s_prefix = acc;
            if (chunk == chunks - 1) n_miss[t] = acc + miss_cnt[t * chunks + chunk];
        }
    }
    __syncthreads();
    const int prefix = s_prefix;
    const TableDesc& T = tabs[tb + t];
    const int64_t S = T.num_sets;
    float* __restrict__ weight = T.weight;
    const float* __restrict__ master = T.master;
    const int64_t* tid = ids + (int64_t)t * ld_ids;
    int32_t* tsl = slots + (int64_t)t * ld_slots;
    float* tout = POOL_P1 ? out + (int64_t)t * ld_out : nullptr;
    const int cpr = dim / VEC;             // vector chunks per row
    const int gl = threadIdx.x % G, group = threadIdx.x / G, NG = 256 / G;
    const int64_t aux_base = S * ways;

    constexpr int U = 4;
    if (cpr <= G) {
        for (int r0 = group; r0 < FWD_CHUNK; r0 += NG * U) {
            const float* src[U];
            int64_t dst_aux[U];
            int j[U];
            bool valid[U];
#pragma unroll
            for (int u = 0; u < U; ++u) {
                int r = r0 + u * NG;
                j[u] = chunk * FWD_CHUNK + r;
                valid[u] = r < FWD_CHUNK && j[u] < n_idx;
                dst_aux[u] = -1;
                src[u] = nullptr;
                if (valid[u]) {
                    int32_t sl = tsl[j[u]];
                    if (sl < 0) {
                        int64_t ord = (int64_t)prefix + (-(int64_t)sl - 1);
                        if (ord >= aux_rows) {        // IndexError in the reference
                            if (gl == 0) atomicOr(flags, 1u);
                            valid[u] = false;
                        } else {
                            dst_aux[u] = aux_base + ord;
                            src[u] = master + __ldg(tid + j[u]) * dim;
                        }
                    } else if (POOL_P1) {
                        src[u] = weight + (int64_t)sl * dim;
                    } else {
                        valid[u] = false;            // hits need no work without copy-out
                    }
                }
            }
            V v[U];
#pragma unroll
            for (int u = 0; u < U; ++u)
                if (valid[u] && gl < cpr) v[u] = reinterpret_cast<const V*>(src[u])[gl];
#pragma unroll
            for (int u = 0; u < U; ++u) {
                if (!valid[u]) continue;
                if (gl < cpr) {
                    if (POOL_P1) reinterpret_cast<V*>(tout + (int64_t)j[u] * dim)[gl] = v[u];
                    if (dst_aux[u] >= 0) reinterpret_cast<V*>(weight + dst_aux[u] * dim)[gl] = v[u];
                }
                if (dst_aux[u] >= 0 && gl == 0) tsl[j[u]] = (int32_t)dst_aux[u];
            }
        }
    } else {  // wide rows: several passes per row
        for (int r = group; r < FWD_CHUNK; r += NG) {
            int j = chunk * FWD_CHUNK + r;
            if (j >= n_idx) break;
            int32_t sl = tsl[j];
            const float* src;
            int64_t dst_aux = -1;
            if (sl < 0) {
                int64_t ord = (int64_t)prefix + (-(int64_t)sl - 1);
                if (ord >= aux_rows) {
                    if (gl == 0) atomicOr(flags, 1u);
                    continue;
                }
                dst_aux = aux_base + ord;
                src = master + __ldg(tid + j) * dim;
            } else if (POOL_P1) {
                src = weight + (int64_t)sl * dim;
            } else {
                continue;
            }
            for (int c = gl; c < cpr; c += G) {
                V v = reinterpret_cast<const V*>(src)[c];
                if (POOL_P1) reinterpret_cast<V*>(tout + (int64_t)j * dim)[c] = v;
                if (dst_aux >= 0) reinterpret_cast<V*>(weight + dst_aux * dim)[c] = v;
            }
            if (dst_aux >= 0 && gl == 0) tsl[j] = (int32_t)dst_aux;
        }
    }
}

// ------------------------------------------------------------------------------
// K2 (fast path, float4 rows of <= 128 floats): every warp owns 32 consecutive ids of
// the chunk.  Lane l first loads slot/id of row l (one coalesced load instead of one
// dependent load per row), resolves the miss ordinal, and the row descriptors are then
// broadcast by shuffle while groups of G lanes move U rows at a time (U*NGW rows, i.e.
// up to 4 KB, in flight per warp).
// ------------------------------------------------------------------------------
template <int G, bool POOL_P1>
__global__ void __launch_bounds__(256) gather_rows_kernel(const TableDesc* __restrict__ tabs, int tb,
                                                          const int64_t* __restrict__ ids, int64_t ld_ids, int n_idx,
                                                          int32_t* __restrict__ slots, int64_t ld_slots,
                                                          const int32_t* __restrict__ miss_cnt, int chunks,
                                                          float* __restrict__ out, int64_t ld_out,
                                                          int32_t* __restrict__ n_miss, uint32_t* __restrict__ flags,
                                                          int dim, int ways, int64_t aux_rows) {
    constexpr int NGW = 32 / G;               // rows moved per warp instruction
    constexpr int ITERS = 32 / NGW;           // == G
    constexpr int U = ITERS < 8 ? ITERS : 8;  // row batches in flight
    const int t = blockIdx.y, chunk = blockIdx.x;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    __shared__ int s_prefix;
    if (warp == 0) {
        int acc = 0;
        for (int c = lane; c < chunk; c += 32) acc += miss_cnt[t * chunks + c];
        acc = warp_sum(acc);
        if (lane == 0) {
            s_prefix = acc;
            if (chunk == chunks - 1) n_miss[t] = acc + miss_cnt[t * chunks + chunk];
        }
    }
    const TableDesc& T = tabs[tb + t];
    float* __restrict__ weight = T.weight;
    const float* __restrict__ master = T.master;
    int32_t* tsl = slots + (int64_t)t * ld_slots;
    float* tout = POOL_P1 ? out + (int64_t)t * ld_out : nullptr;
    const int64_t aux_base = T.num_sets * ways;
    const int cpr = dim >> 2;
    const int gl = lane % G, g = lane / G;
    // lane-parallel descriptor of row (j0 + lane)
    const int j0 = chunk * FWD_CHUNK + warp * 32;
    const int jl = j0 + lane;
    const bool in_range = jl < n_idx;
    const int32_t sl = in_range ? tsl[jl] : 0;
    const int64_t id = (in_range && sl < 0) ? __ldg(ids + (int64_t)t * ld_ids + jl) : 0;
    __syncthreads();
    const int prefix = s_prefix;
    const float* src_l = nullptr;     // null: nothing to move for this row
    int64_t aux_l = -1;
    if (in_range) {
        if (sl < 0) {
            const int64_t ord = (int64_t)prefix + (-(int64_t)sl - 1);
            if (ord >= aux_rows) {
                atomicOr(flags, 1u);               // IndexError in the reference
            } else {
                aux_l = aux_base + ord;
                src_l = master + id * dim;
                tsl[jl] = (int32_t)aux_l;
            }
        } else if (POOL_P1) {
            src_l = weight + (int64_t)sl * dim;
        }
    }
    const unsigned long long src_bits = (unsigned long long)src_l;
#pragma unroll 1
    for (int it0 = 0; it0 < ITERS; it0 += U) {
        float4 v[U];
        unsigned long long sp[U];
        long long ax[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int rr = (it0 + u) * NGW + g;                 // row (within the warp's 32) of this group
            sp[u] = __shfl_sync(0xffffffffu, src_bits, rr);
            ax[u] = __shfl_sync(0xffffffffu, (long long)aux_l, rr);
        }
#pragma unroll
        for (int u = 0; u < U; ++u)
            if (sp[u] && gl < cpr) v[u] = reinterpret_cast<const float4*>(sp[u])[gl];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            if (!sp[u] || gl >= cpr) continue;
            const int rr = (it0 + u) * NGW + g;
            if (POOL_P1) reinterpret_cast<float4*>(tout + (int64_t)(j0 + rr) * dim)[gl] = v[u];
            if (ax[u] >= 0) reinterpret_cast<float4*>(weight + ax[u] * dim)[gl] = v[u];
        }
    }
}

// ------------------------------------------------------------------------------
// K3: EmbeddingBag(mode="sum") over final slots for general offsets
// (model_no_ddp.py:191,200-202).  One group of G lanes per bag.
// ------------------------------------------------------------------------------
template <int VEC>
__global__ void __launch_bounds__(256) pool_kernel(const TableDesc* __restrict__ tabs, int tb,
                                                   const int32_t* __restrict__ slots, int64_t ld_slots,
                                                   const int64_t* __restrict__ offsets, int64_t ld_off,
                                                   int n_idx, int n_bags, float* __restrict__ out, int64_t ld_out,
                                                   int32_t* __restrict__ bag_ids, int64_t ld_bag, int dim, int G) {
    using V = typename VecT<VEC>::type;
    const int t = blockIdx.y;
    const int gl = threadIdx.x % G, group = threadIdx.x / G, NG = 256 / G;
    const int b = blockIdx.x * NG + group;
    if (b >= n_bags) return;
    const float* __restrict__ weight = tabs[tb + t].weight;
    const int64_t* off = offsets + (int64_t)t * ld_off;
    const int32_t* tsl = slots + (int64_t)t * ld_slots;
    int64_t lo = off[b], hi = (b + 1 < n_bags) ? off[b + 1] : n_idx;
    const int cpr = dim / VEC;
    if (bag_ids && gl == 0)
        for (int64_t j = lo; j < hi; ++j) bag_ids[(int64_t)t * ld_bag + j] = b;
    for (int c = gl; c < cpr; c += G) {
        V acc;
        vzero(acc);
        for (int64_t j = lo; j < hi; ++j)
            acc = vadd(acc, reinterpret_cast<const V*>(weight + (int64_t)tsl[j] * dim)[c]);
        reinterpret_cast<V*>(out + (int64_t)t * ld_out + (int64_t)b * dim)[c] = acc;
    }
}

// ------------------------------------------------------------------------------
// Backward plan: one 1024-thread CTA per table radix-sorts (slot, position) pairs
// in shared memory (stable LSD, 8-bit digits, ranks from __match_any_sync), then
// cuts the sorted run into chunks of <= CH equal slots.  A chunk that covers its
// whole segment is applied with a plain read-modify-write, the (rare) others with
// vector reds.  This is the "deduplicated indices instead of atomics" step.
// ------------------------------------------------------------------------------
constexpr int PLAN_NT = 1024;
constexpr int PLAN_MAX_ROUNDS = CDLRM_SORT_MAX / 1024;  // 16 keys per thread at most

struct PlanView {
    int32_t* sorted_pos;    // [tc][n_idx] absolute position j, grouped by slot (stable)
    int4* chunks;           // [tc][n_idx] {slot, start (index into sorted_pos, relative to the table),
                            //              len | first<<30 | excl<<31, position of the first element}
    int32_t* n_chunks;      // [tc][nsub][2] {singles (listed from the front), multis (from the back)}
};

__host__ __device__ inline int plan_nsub(int n_idx) { return (n_idx + CDLRM_SORT_MAX - 1) / CDLRM_SORT_MAX; }

__host__ inline PlanView plan_view(void* plan, int tc, int n_idx) {
    PlanView v;
    char* p = (char*)plan;
    size_t a = (size_t)tc * n_idx * sizeof(int32_t);
    a = (a + 255) & ~(size_t)255;
    v.sorted_pos = (int32_t*)p; p += a;
    v.chunks = (int4*)p; p += 4 * a;
    v.n_chunks = (int32_t*)p;
    return v;
}

__device__ __forceinline__ int block_excl_scan_1024(int v, int* s_warp, int& total) {
    // exclusive sum scan over 1024 threads; s_warp: 32 ints of shared scratch
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        int n = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += n;
    }
    if (lane == 31) s_warp[warp] = inc;
    __syncthreads();
    if (warp == 0) {
        int w = s_warp[lane];
        int winc = w;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            int n = __shfl_up_sync(0xffffffffu, winc, o);
            if (lane >= o) winc += n;
        }
        s_warp[lane] = winc - w;  // exclusive warp offsets
        if (lane == 31) s_warp[32] = winc;
    }
    __syncthreads();
    int res = s_warp[warp] + inc - v;
    total = s_warp[32];
    __syncthreads();
    return res;
}

__device__ __forceinline__ int block_excl_maxscan_1024(int v, int* s_warp) {
    // exclusive max scan (identity -1) over 1024 threads
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        int n = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc = max(inc, n);
    }
    if (lane == 31) s_warp[warp] = inc;
    __syncthreads();
    if (warp == 0) {
        int w = s_warp[lane];
        int winc = w;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            int n = __shfl_up_sync(0xffffffffu, winc, o);
            if (lane >= o) winc = max(winc, n);
        }
        int wex = __shfl_up_sync(0xffffffffu, winc, 1);
        s_warp[lane] = lane == 0 ? -1 : wex;
    }
    __syncthreads();
    int ex = __shfl_up_sync(0xffffffffu, inc, 1);
    if (lane == 0) ex = -1;
    int res = max(s_warp[warp], ex);
    __syncthreads();
    return res;
}

__global__ void __launch_bounds__(PLAN_NT, 1) bwd_plan_kernel(const TableDesc* __restrict__ tabs, int tb,
                                                              const int32_t* __restrict__ slots, int64_t ld_slots,
                                                              int n_idx, int j0, int n, int sub, int nsub,
                                                              PlanView pv) {
    extern __shared__ __align__(16) unsigned char smem[];
    // layout: keyA[n] keyB[n] (u32) | valA[n] valB[n] (u16) | cnt[8192] (u16) | scratch[40] (int)
    const int npad = (n + 7) & ~7;
    uint32_t* keyA = reinterpret_cast<uint32_t*>(smem);
    uint32_t* keyB = keyA + npad;
    uint16_t* valA = reinterpret_cast<uint16_t*>(keyB + npad);
    uint16_t* valB = valA + npad;
    uint16_t* cnt = valB + npad;                       // [256 digits][32 warps]
    int* s_scr = reinterpret_cast<int*>(cnt + 8192);

    const int t = blockIdx.x;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int32_t* tsl = slots + (int64_t)t * ld_slots + j0;
    for (int i = tid; i < n; i += PLAN_NT) {
        keyA[i] = (uint32_t)tsl[i];
        valA[i] = (uint16_t)i;
    }
    const int64_t rows = tabs[tb + t].cache_rows;
    const int bits = 64 - __clzll((unsigned long long)(rows > 1 ? rows - 1 : 1));
    const int passes = (bits + 7) / 8;
    const int per_warp = (((n + 31) / 32) + 31) & ~31;  // keys per warp, multiple of 32
    const int rounds = per_warp / 32;                   // <= PLAN_MAX_ROUNDS
    const uint32_t lt = (1u << lane) - 1u;
    __syncthreads();

    for (int pass = 0; pass < passes; ++pass) {
        const int shift = pass * 8;
        for (int e = tid; e < 4096; e += PLAN_NT) reinterpret_cast<uint32_t*>(cnt)[e] = 0;
        __syncthreads();
        uint16_t rk[PLAN_MAX_ROUNDS];
#pragma unroll
        for (int r = 0; r < PLAN_MAX_ROUNDS; ++r) {
            if (r < rounds) {
                int i = warp * per_warp + r * 32 + lane;
                bool valid = i < n;
                uint32_t dg = valid ? ((keyA[i] >> shift) & 255u) : 256u;
                uint32_t peers = __match_any_sync(0xffffffffu, dg);
                int leader = __ffs(peers) - 1;
                int old = 0;
                if (valid && lane == leader) {
                    old = cnt[dg * 32 + warp];
                    cnt[dg * 32 + warp] = (uint16_t)(old + __popc(peers));
                }
                old = __shfl_sync(0xffffffffu, old, leader);
                rk[r] = (uint16_t)(old + __popc(peers & lt));
                __syncwarp();
            }
        }
        __syncthreads();
        {   // exclusive scan of cnt in (digit, warp) order: 8 entries per thread
            uint4 raw = reinterpret_cast<uint4*>(cnt)[tid];
            uint16_t c[8];
            c[0] = raw.x & 0xffff; c[1] = raw.x >> 16; c[2] = raw.y & 0xffff; c[3] = raw.y >> 16;
            c[4] = raw.z & 0xffff; c[5] = raw.z >> 16; c[6] = raw.w & 0xffff; c[7] = raw.w >> 16;
            int sum = 0;
#pragma unroll
            for (int q = 0; q < 8; ++q) sum += c[q];
            int total;
            int ex = block_excl_scan_1024(sum, s_scr, total);
#pragma unroll
            for (int q = 0; q < 8; ++q) {
                int v = c[q];
                c[q] = (uint16_t)ex;
                ex += v;
            }
            raw.x = c[0] | ((uint32_t)c[1] << 16); raw.y = c[2] | ((uint32_t)c[3] << 16);
            raw.z = c[4] | ((uint32_t)c[5] << 16); raw.w = c[6] | ((uint32_t)c[7] << 16);
            reinterpret_cast<uint4*>(cnt)[tid] = raw;
        }
        __syncthreads();
#pragma unroll
        for (int r = 0; r < PLAN_MAX_ROUNDS; ++r) {
            if (r < rounds) {
                int i = warp * per_warp + r * 32 + lane;
                if (i < n) {
                    uint32_t k = keyA[i];
                    uint32_t dg = (k >> shift) & 255u;
                    int dst = cnt[dg * 32 + warp] + rk[r];
                    keyB[dst] = k;
                    valB[dst] = valA[i];
                }
            }
        }
        __syncthreads();
        uint32_t* tk = keyA; keyA = keyB; keyB = tk;
        uint16_t* tv = valA; valA = valB; valB = tv;
    }

    // ---- cut into chunks ---------------------------------------------------------
    const int items = (n + PLAN_NT - 1) / PLAN_NT;
    const int i0 = tid * items, i1 = min(n, i0 + items);
    int last_head = -1;
    for (int i = i0; i < i1; ++i)
        if (i == 0 || keyA[i] != keyA[i - 1]) last_head = i;
    int carry = block_excl_maxscan_1024(last_head, s_scr);
    // singles (one gradient row, sole owner of its slot: the common case) are listed from the
    // front of the chunk array, multi-row / shared-slot chunks from the back, so that each kind
    // gets its own specialised apply kernel
    int cur = carry, n_single = 0, n_multi = 0;
    for (int i = i0; i < i1; ++i) {
        const bool head = (i == 0 || keyA[i] != keyA[i - 1]);
        if (head) cur = i;
        if (((i - cur) % CH) == 0) {
            const bool single = head && (i + 1 == n || keyA[i + 1] != keyA[i]);
            if (single) ++n_single; else ++n_multi;
        }
    }
    int tot_single, tot_multi;
    int sidx = block_excl_scan_1024(n_single, s_scr, tot_single);
    int midx = block_excl_scan_1024(n_multi, s_scr, tot_multi);
    const int64_t obase = (int64_t)t * n_idx + j0;
    cur = carry;
    for (int i = i0; i < i1; ++i) {
        const bool head = (i == 0 || keyA[i] != keyA[i - 1]);
        if (head) cur = i;
        if (((i - cur) % CH) == 0) {
            const uint32_t k = keyA[i];
            int len = 1;
            while (len < CH && i + len < n && keyA[i + len] == k) ++len;
            const bool last = (i + len == n) || keyA[i + len] != k;
            const uint32_t desc = (uint32_t)len | (head ? (1u << 30) : 0u) | ((head && last) ? (1u << 31) : 0u);
            const int4 rec = make_int4((int)k, j0 + i, (int)desc, j0 + (int)valA[i]);
            if (head && last && len == 1) pv.chunks[obase + sidx++] = rec;
            else pv.chunks[obase + n - 1 - midx++] = rec;
        }
        pv.sorted_pos[obase + i] = j0 + (int)valA[i];
    }
    if (tid == 0) {
        pv.n_chunks[(t * nsub + sub) * 2 + 0] = tot_single;
        pv.n_chunks[(t * nsub + sub) * 2 + 1] = tot_multi;
    }
}

// ------------------------------------------------------------------------------
// Backward apply.  A chunk = up to CH gradient rows that hit the same slot; its record
// {slot, start, len|flags, first position} is one 16-byte load.  acc = sum of the chunk's
// upstream gradient rows (ascending position: deterministic), then
// weight[slot] += -lr * acc -- a plain read-modify-write when the chunk owns the slot,
// a vector red otherwise (only slots with more than CH contributions).
// Fast path (float4 rows, dim <= 128): a warp takes 32 consecutive chunk records with one
// coalesced load and groups of G lanes work on U chunks at a time, so that U gradient
// rows and U weight rows are in flight per group.
// ------------------------------------------------------------------------------
template <int G>
__global__ void __launch_bounds__(256, 4) bwd_sgd_single_kernel(const TableDesc* __restrict__ tabs, int tb,
                                                                PlanView pv, int n_idx, int j0, int sub, int nsub,
                                                                const int32_t* __restrict__ bag_ids, int64_t ld_bag,
                                                                const float* __restrict__ d_out, int64_t ld_dout,
                                                                int64_t row_stride, float lr, int dim) {
    // singles: weight[slot] = fma(-lr, d_out[pos], weight[slot]) -- a pure streaming RMW
    // a warp takes RPW records: few enough that even a short singles list spreads over all SMs
    constexpr int NGW = 32 / G;
    constexpr int RPW = NGW > 8 ? NGW : 8;
    constexpr int ITERS = RPW / NGW;
    constexpr int U = ITERS < 4 ? ITERS : 4;
    const int t = blockIdx.y;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int nch = pv.n_chunks[(t * nsub + sub) * 2];
    const int c0 = (blockIdx.x * 8 + warp) * RPW;
    if (c0 >= nch) return;
    const int64_t obase = (int64_t)t * n_idx + j0;
    const int gl = lane % G, g = lane / G;
    const int cpr = dim >> 2;
    const TableDesc& T = tabs[tb + t];
    float* __restrict__ weight = T.weight;
    const float* gbase = d_out + (int64_t)t * ld_dout;
    int slot_l = -1, pos_l = 0;
    if (lane < RPW && c0 + lane < nch) {
        const int4 rec = pv.chunks[obase + c0 + lane];
        slot_l = rec.x;
        pos_l = bag_ids ? bag_ids[(int64_t)t * ld_bag + rec.w] : rec.w;
        if (T.dirty) atomicOr(T.dirty + (slot_l >> 5), 1u << (slot_l & 31));
    }
    const bool act = gl < cpr;
#pragma unroll 1
    for (int it0 = 0; it0 < ITERS; it0 += U) {
        int slot[U], pos[U];
        float4 gv[U], wv[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int cc = (it0 + u) * NGW + g;
            slot[u] = __shfl_sync(0xffffffffu, slot_l, cc);
            pos[u] = __shfl_sync(0xffffffffu, pos_l, cc);
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            if (slot[u] >= 0 && act) {
                gv[u] = reinterpret_cast<const float4*>(gbase + (int64_t)pos[u] * row_stride)[gl];
                wv[u] = reinterpret_cast<const float4*>(weight + (int64_t)slot[u] * dim)[gl];
            }
        }
#pragma unroll
        for (int u = 0; u < U; ++u)
            if (slot[u] >= 0 && act)
                reinterpret_cast<float4*>(weight + (int64_t)slot[u] * dim)[gl] = vfma(-lr, gv[u], wv[u]);
    }
}

template <int G>
__global__ void __launch_bounds__(256) bwd_sgd_multi_kernel(const TableDesc* __restrict__ tabs, int tb,
                                                            PlanView pv, int n_idx, int j0, int n, int sub, int nsub,
                                                            const int32_t* __restrict__ bag_ids, int64_t ld_bag,
                                                            const float* __restrict__ d_out, int64_t ld_dout,
                                                            int64_t row_stride, float lr, int dim) {
    // multis: up to CH gradient rows of one slot are loaded together, summed in ascending
    // position and applied with a plain RMW when the chunk is the slot's only chunk, else a red
    constexpr int NG = 256 / G;
    const int t = blockIdx.y;
    const int gl = threadIdx.x % G, group = threadIdx.x / G;
    const int m = blockIdx.x * NG + group;
    if (m >= pv.n_chunks[(t * nsub + sub) * 2 + 1]) return;
    const int64_t obase = (int64_t)t * n_idx + j0;
    const int64_t tbase = (int64_t)t * n_idx;
    const int4 rec = pv.chunks[obase + n - 1 - m];
    const int len = rec.z & 0xff;
    const int32_t* bag = bag_ids ? bag_ids + (int64_t)t * ld_bag : nullptr;
    const float* gbase = d_out + (int64_t)t * ld_dout;
    const TableDesc& T = tabs[tb + t];
    if (((rec.z >> 30) & 1) && gl == 0 && T.dirty) atomicOr(T.dirty + (rec.x >> 5), 1u << (rec.x & 31));
    if (gl >= (dim >> 2)) return;
    int pp[CH];
    float4 ga[CH];
#pragma unroll
    for (int i = 0; i < CH; ++i) {
        pp[i] = i < len ? pv.sorted_pos[tbase + rec.y + i] : 0;
        if (bag && i < len) pp[i] = bag[pp[i]];
    }
#pragma unroll
    for (int i = 0; i < CH; ++i)
        if (i < len) ga[i] = reinterpret_cast<const float4*>(gbase + (int64_t)pp[i] * row_stride)[gl];
    float4* wp = reinterpret_cast<float4*>(T.weight + (int64_t)rec.x * dim) + gl;
    float4 wv;
    if (rec.z < 0) wv = *wp;
    float4 acc = ga[0];
#pragma unroll
    for (int i = 1; i < CH; ++i)
        if (i < len) acc = vadd(acc, ga[i]);
    if (rec.z < 0) *wp = vfma(-lr, acc, wv);
    else red_add(wp, vscale(-lr, acc));
}

// generic fallback (any dim / alignment): one group of G lanes per chunk
template <int VEC>
__global__ void __launch_bounds__(256) bwd_sgd_kernel(const TableDesc* __restrict__ tabs, int tb,
                                                      PlanView pv, int n_idx, int j0, int n, int sub, int nsub,
                                                      const int32_t* __restrict__ bag_ids, int64_t ld_bag,
                                                      const float* __restrict__ d_out, int64_t ld_dout,
                                                      int64_t row_stride, float lr, int dim, int G) {
    using V = typename VecT<VEC>::type;
    const int t = blockIdx.y;
    const int gl = threadIdx.x % G, group = threadIdx.x / G, NG = 256 / G;
    const int c = blockIdx.x * NG + group;
    const int ns = pv.n_chunks[(t * nsub + sub) * 2], nm = pv.n_chunks[(t * nsub + sub) * 2 + 1];
    if (c >= ns + nm) return;
    const int64_t obase = (int64_t)t * n_idx + j0;
    const int4 rec = c < ns ? pv.chunks[obase + c] : pv.chunks[obase + n - 1 - (c - ns)];
    const int len = rec.z & 0xff;
    const bool first = (rec.z >> 30) & 1, excl = rec.z < 0;
    const int32_t slot = rec.x;
    const int32_t* pos = pv.sorted_pos + (int64_t)t * n_idx + rec.y;
    const int32_t* bag = bag_ids ? bag_ids + (int64_t)t * ld_bag : nullptr;
    const float* g = d_out + (int64_t)t * ld_dout;
    const TableDesc& T = tabs[tb + t];
    float* wrow = T.weight + (int64_t)slot * dim;
    const int cpr = dim / VEC;
    for (int cc = gl; cc < cpr; cc += G) {
        V acc;
        vzero(acc);
        for (int i = 0; i < len; ++i) {
            int p0 = pos[i];
            if (bag) p0 = bag[p0];
            acc = vadd(acc, reinterpret_cast<const V*>(g + (int64_t)p0 * row_stride)[cc]);
        }
        V* wp = reinterpret_cast<V*>(wrow) + cc;
        if (excl) *wp = vfma(-lr, acc, *wp);
        else red_add(wp, vscale(-lr, acc));
    }
    if (first && gl == 0 && T.dirty) atomicOr(T.dirty + (slot >> 5), 1u << (slot & 31));
}

inline int pow2_ceil(int v) {
    int p = 1;
    while (p < v) p <<= 1;
    return p;
}

inline int pick_vec(int dim, const void* a, const void* b, int64_t s1, int64_t s2) {
    auto al = [](const void* p, int n) { return p == nullptr || ((uintptr_t)p % n) == 0; };
    if (dim % 4 == 0 && al(a, 16) && al(b, 16) && s1 % 4 == 0 && s2 % 4 == 0) return 4;
    if (dim % 2 == 0 && al(a, 8) && al(b, 8) && s1 % 2 == 0 && s2 % 2 == 0) return 2;
    return 1;
}

}  // namespace

static int ensure_scratch(cdlrm_ctx* c, int64_t n_idx) {
    if (n_idx <= c->scratch_max_idx && c->d_miss_cnt) return CDLRM_OK;
    return cdlrm_ctx_reserve(c, n_idx > 8192 ? n_idx : 8192);
}

extern "C" int cdlrm_embed_fwd(cdlrm_ctx* c, int tb, int tc, const int64_t* ids, int64_t ld_ids,
                               const int64_t* offsets, int64_t ld_off, int32_t n_idx, int32_t n_bags,
                               float* out, int64_t ld_out, int32_t* slots, int64_t ld_slots, int32_t* n_miss,
                               int32_t* bag_ids, int64_t ld_bag, cdlrm_stream stream) {
    ARG_CHECK(c && ids && out && slots && n_miss);
    ARG_CHECK(tb >= 0 && tc > 0 && tb + tc <= c->T);
    ARG_CHECK(n_idx >= 0 && n_bags >= 0);
    ARG_CHECK(offsets != nullptr || n_idx == n_bags);
    cudaStream_t s = (cudaStream_t)stream;
    CU_CHECK(cudaSetDevice(c->device));
    for (int k = tb; k < tb + tc; ++k) {
        if (!c->tabs[k].weight || !c->tabs[k].tags || !c->tabs[k].master) {
            cdlrm_set_error("table %d: cache or master not bound", k);
            return CDLRM_ERR_STATE;
        }
    }
    int rc = cdlrm_sync_tabs(c, s);
    if (rc) return rc;
    if (n_idx == 0) {
        CU_CHECK(cudaMemsetAsync(n_miss, 0, sizeof(int32_t) * tc, s));
        if (n_bags > 0)
            for (int t = 0; t < tc; ++t)
                CU_CHECK(cudaMemsetAsync(out + (int64_t)t * ld_out, 0, sizeof(float) * (int64_t)n_bags * c->dim, s));
        return CDLRM_OK;
    }
    rc = ensure_scratch(c, n_idx);
    if (rc) return rc;
    const int chunks = (n_idx + FWD_CHUNK - 1) / FWD_CHUNK;
    dim3 grid(chunks, tc);
    const int gw = pow2_ceil(c->ways) > 32 ? 32 : pow2_ceil(c->ways);
#define LAUNCH_PROBE(GW) LAUNCH(K_PROBE, s, probe_kernel<GW><<<grid, ProbeCfg<GW>::NT, 0, s>>>(c->d_tabs, tb, ids, ld_ids, n_idx, slots, ld_slots, c->d_miss_cnt, chunks, c->ways))
    switch (gw) {
        case 1: LAUNCH_PROBE(1); break;
        case 2: LAUNCH_PROBE(2); break;
        case 4: LAUNCH_PROBE(4); break;
        case 8: LAUNCH_PROBE(8); break;
        case 16: LAUNCH_PROBE(16); break;
        default: LAUNCH_PROBE(32); break;
    }
#undef LAUNCH_PROBE
    CU_CHECK(cudaGetLastError());
    const bool p1 = offsets == nullptr;
    bool al16 = true, al8 = true;
    for (int k = tb; k < tb + tc; ++k) {
        al16 = al16 && ((uintptr_t)c->tabs[k].master % 16 == 0);
        al8 = al8 && ((uintptr_t)c->tabs[k].master % 8 == 0);
    }
    int vec = pick_vec(c->dim, out, nullptr, ld_out, c->dim);
    if (vec == 4 && !al16) vec = 2;
    if (vec == 2 && !al8) vec = 1;
    const int cpr = c->dim / vec;
    const int G = pow2_ceil(cpr) > 32 ? 32 : pow2_ceil(cpr);
#define LAUNCH_GATHER(VEC, P1) LAUNCH(K_GATHER, s, (gather_kernel<VEC, P1><<<grid, 256, 0, s>>>(c->d_tabs, tb, ids, ld_ids, n_idx, slots, ld_slots, c->d_miss_cnt, chunks, out, ld_out, n_miss, c->d_flags, c->dim, c->ways, c->aux, G)))
#define LAUNCH_GROWS(GG, P1) LAUNCH(K_GATHER, s, (gather_rows_kernel<GG, P1><<<grid, 256, 0, s>>>(c->d_tabs, tb, ids, ld_ids, n_idx, slots, ld_slots, c->d_miss_cnt, chunks, out, ld_out, n_miss, c->d_flags, c->dim, c->ways, c->aux)))
#define GROWS_SWITCH(P1)                                  \
    switch (G) {                                          \
        case 1: LAUNCH_GROWS(1, P1); break;               \
        case 2: LAUNCH_GROWS(2, P1); break;               \
        case 4: LAUNCH_GROWS(4, P1); break;               \
        case 8: LAUNCH_GROWS(8, P1); break;               \
        case 16: LAUNCH_GROWS(16, P1); break;             \
        default: LAUNCH_GROWS(32, P1); break;             \
    }
    if (vec == 4 && cpr <= 32) {
        if (p1) { GROWS_SWITCH(true) } else { GROWS_SWITCH(false) }
    } else if (p1) {
        if (vec == 4) LAUNCH_GATHER(4, true); else if (vec == 2) LAUNCH_GATHER(2, true); else LAUNCH_GATHER(1, true);
    } else {
        if (vec == 4) LAUNCH_GATHER(4, false); else if (vec == 2) LAUNCH_GATHER(2, false); else LAUNCH_GATHER(1, false);
    }
#undef GROWS_SWITCH
#undef LAUNCH_GROWS
#undef LAUNCH_GATHER
    CU_CHECK(cudaGetLastError());
    if (!p1) {
        if (n_bags > 0) {
            const int NG = 256 / G;
            dim3 pg((n_bags + NG - 1) / NG, tc);
            if (vec == 4) LAUNCH(K_POOL, s, pool_kernel<4><<<pg, 256, 0, s>>>(c->d_tabs, tb, slots, ld_slots, offsets, ld_off, n_idx, n_bags, out, ld_out, bag_ids, ld_bag, c->dim, G));
            else if (vec == 2) LAUNCH(K_POOL, s, pool_kernel<2><<<pg, 256, 0, s>>>(c->d_tabs, tb, slots, ld_slots, offsets, ld_off, n_idx, n_bags, out, ld_out, bag_ids, ld_bag, c->dim, G));
            else LAUNCH(K_POOL, s, pool_kernel<1><<<pg, 256, 0, s>>>(c->d_tabs, tb, slots, ld_slots, offsets, ld_off, n_idx, n_bags, out, ld_out, bag_ids, ld_bag, c->dim, G));
            CU_CHECK(cudaGetLastError());
        }
    }
    return CDLRM_OK;
}

extern "C" int64_t cdlrm_embed_bwd_plan_bytes(int tc, int32_t n_idx) {
    if (tc <= 0 || n_idx < 0) return -1;
    size_t a = (size_t)tc * (size_t)(n_idx > 0 ? n_idx : 1) * sizeof(int32_t);
    a = (a + 255) & ~(size_t)255;
    size_t nsub = (size_t)plan_nsub(n_idx > 0 ? n_idx : 1);
    return (int64_t)(5 * a + ((((size_t)tc * nsub * 2 * sizeof(int32_t)) + 255) & ~(size_t)255));
}

extern "C" int cdlrm_embed_bwd_plan(cdlrm_ctx* c, int tb, int tc, const int32_t* slots, int64_t ld_slots,
                                    int32_t n_idx, void* plan, cdlrm_stream stream) {
    ARG_CHECK(c && slots && plan);
    ARG_CHECK(tb >= 0 && tc > 0 && tb + tc <= c->T && n_idx >= 0);
    ARG_CHECK(((uintptr_t)plan & 255) == 0);
    if (n_idx == 0) return CDLRM_OK;
    cudaStream_t s = (cudaStream_t)stream;
    CU_CHECK(cudaSetDevice(c->device));
    int rc = cdlrm_sync_tabs(c, s);
    if (rc) return rc;
    PlanView pv = plan_view(plan, tc, n_idx);
    const int nsub = plan_nsub(n_idx);
    static bool attr_done = false;
    const int max_smem = CDLRM_SORT_MAX * 12 + 8192 * 2 + 64 * 4;
    if (!attr_done) {
        CU_CHECK(cudaFuncSetAttribute(bwd_plan_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem));
        attr_done = true;
    }
    for (int sub = 0; sub < nsub; ++sub) {
        const int j0 = sub * CDLRM_SORT_MAX;
        const int n = n_idx - j0 < CDLRM_SORT_MAX ? n_idx - j0 : CDLRM_SORT_MAX;
        const int npad = (n + 7) & ~7;
        const int smem = npad * 12 + 8192 * 2 + 64 * 4;
        LAUNCH(K_BWD_PLAN, s, bwd_plan_kernel<<<tc, PLAN_NT, smem, s>>>(c->d_tabs, tb, slots, ld_slots, n_idx, j0, n, sub, nsub, pv));
    }
    CU_CHECK(cudaGetLastError());
    return CDLRM_OK;
}

extern "C" int cdlrm_embed_bwd_sgd(cdlrm_ctx* c, int tb, int tc, const void* plan, int32_t n_idx,
                                   const int32_t* bag_ids, int64_t ld_bag, const float* d_out, int64_t ld_dout,
                                   int64_t row_stride, float lr, cdlrm_stream stream) {
    ARG_CHECK(c && plan && d_out);
    ARG_CHECK(tb >= 0 && tc > 0 && tb + tc <= c->T && n_idx >= 0);
    if (n_idx == 0) return CDLRM_OK;
    cudaStream_t s = (cudaStream_t)stream;
    CU_CHECK(cudaSetDevice(c->device));
    int rc = cdlrm_sync_tabs(c, s);
    if (rc) return rc;
    PlanView pv = plan_view(const_cast<void*>(plan), tc, n_idx);
    const int nsub = plan_nsub(n_idx);
    const int vec = pick_vec(c->dim, d_out, nullptr, ld_dout, row_stride);
    const int cpr = c->dim / vec;
    const int G = pow2_ceil(cpr) > 32 ? 32 : pow2_ceil(cpr);
    const int NG = 256 / G;
    // sub-batches are applied one after the other: the update is additive in the
    // gradients, so the split only changes fp32 rounding order.
    for (int sub = 0; sub < nsub; ++sub) {
        const int j0 = sub * CDLRM_SORT_MAX;
        const int n = n_idx - j0 < CDLRM_SORT_MAX ? n_idx - j0 : CDLRM_SORT_MAX;
        if (vec == 4 && cpr <= 32) {
            const int rpw = (32 / G) > 8 ? (32 / G) : 8;
            dim3 grid((n + 8 * rpw - 1) / (8 * rpw), tc);       // singles: 8 warps x RPW records per CTA
            dim3 gridm((n / 2 + NG - 1) / NG + 1, tc);          // multis: at most n/2 chunks, one group each
#define LAUNCH_SGD(GG)                                                                                         \
    LAUNCH(K_BWD_SGD, s, (bwd_sgd_single_kernel<GG><<<grid, 256, 0, s>>>(c->d_tabs, tb, pv, n_idx, j0, sub, nsub, bag_ids, ld_bag, d_out, ld_dout, row_stride, lr, c->dim))); \
    LAUNCH(K_BWD_SGD_MULTI, s, (bwd_sgd_multi_kernel<GG><<<gridm, 256, 0, s>>>(c->d_tabs, tb, pv, n_idx, j0, n, sub, nsub, bag_ids, ld_bag, d_out, ld_dout, row_stride, lr, c->dim)))
            switch (G) {
                case 1: LAUNCH_SGD(1); break;
                case 2: LAUNCH_SGD(2); break;
                case 4: LAUNCH_SGD(4); break;
                case 8: LAUNCH_SGD(8); break;
                case 16: LAUNCH_SGD(16); break;
                default: LAUNCH_SGD(32); break;
            }
#undef LAUNCH_SGD
        } else {
            dim3 grid((n + NG - 1) / NG, tc);
            if (vec == 4) LAUNCH(K_BWD_SGD, s, bwd_sgd_kernel<4><<<grid, 256, 0, s>>>(c->d_tabs, tb, pv, n_idx, j0, n, sub, nsub, bag_ids, ld_bag, d_out, ld_dout, row_stride, lr, c->dim, G));
            else if (vec == 2) LAUNCH(K_BWD_SGD, s, bwd_sgd_kernel<2><<<grid, 256, 0, s>>>(c->d_tabs, tb, pv, n_idx, j0, n, sub, nsub, bag_ids, ld_bag, d_out, ld_dout, row_stride, lr, c->dim, G));
            else LAUNCH(K_BWD_SGD, s, bwd_sgd_kernel<1><<<grid, 256, 0, s>>>(c->d_tabs, tb, pv, n_idx, j0, n, sub, nsub, bag_ids, ld_bag, d_out, ld_dout, row_stride, lr, c->dim, G));
        }
    }
    CU_CHECK(cudaGetLastError());
    return CDLRM_OK;
}
