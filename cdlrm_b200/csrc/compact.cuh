// compact.cuh -- block scan + bitmap compaction kernels shared by the window planner
// (plan.cu) and the aggregation collect (move.cu).
#pragma once
#include "common.cuh"

namespace {

constexpr int TILE = 1024;  // items per CTA in the compaction kernels (256 threads x 4)

template <int NT>
__device__ __forceinline__ int block_excl_scan(int v, int* s_warp, int& total) {
    constexpr int NW = NT / 32;
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
        int w = lane < NW ? s_warp[lane] : 0;
        int winc = w;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            int n = __shfl_up_sync(0xffffffffu, winc, o);
            if (lane >= o) winc += n;
        }
        if (lane < NW) s_warp[lane] = winc - w;
        if (lane == 31) s_warp[32] = winc;
    }
    __syncthreads();
    int res = s_warp[warp] + inc - v;
    total = s_warp[32];
    __syncthreads();
    return res;
}

// ---- A1: set one bit per id ---------------------------------------------------------
__global__ void __launch_bounds__(256) bitmap_set_kernel(const int64_t* __restrict__ ids, int64_t n,
                                                         uint32_t* __restrict__ bitmap, int64_t n_rows,
                                                         uint32_t* __restrict__ flags) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (; i < n; i += stride) {
        const int64_t id = __ldg(ids + i);
        if ((uint64_t)id >= (uint64_t)n_rows) {
            atomicOr(flags, 2u);  // id outside the table: IndexError in the reference
            continue;
        }
        uint32_t* w = bitmap + (id >> 5);
        const uint32_t bit = 1u << (id & 31);
        // a stale L1 line can only under-report set bits -> at worst a redundant atomic
        if (!(*w & bit)) atomicOr(w, bit);
    }
}

// ---- A2: popcount per tile of 1024 words ----------------------------------------------
__global__ void __launch_bounds__(256) bitmap_count_kernel(const uint32_t* __restrict__ bitmap, int64_t nwords,
                                                           int32_t* __restrict__ blocksum) {
    __shared__ int s_w[33];
    const int64_t w0 = (int64_t)blockIdx.x * TILE + threadIdx.x * 4;
    int c = 0;
#pragma unroll
    for (int q = 0; q < 4; ++q)
        if (w0 + q < nwords) c += __popc(bitmap[w0 + q]);
    int total;
    block_excl_scan<256>(c, s_w, total);
    if (threadIdx.x == 0) blocksum[blockIdx.x] = total;
}

// ---- exclusive scan of the per-tile counts (single CTA), total -> *total_out -------------
__global__ void __launch_bounds__(1024) scan_tiles_kernel(int32_t* __restrict__ blocksum, int nblk,
                                                          unsigned long long* __restrict__ total_out) {
    __shared__ int s_w[33];
    int carry = 0;
    for (int base = 0; base < nblk; base += 1024) {
        const int i = base + threadIdx.x;
        const int v = i < nblk ? blocksum[i] : 0;
        int total;
        const int ex = block_excl_scan<1024>(v, s_w, total);
        if (i < nblk) blocksum[i] = carry + ex;
        carry += total;
    }
    if (threadIdx.x == 0) *total_out = (unsigned long long)carry;
}

// ---- A4: emit ascending set-bit indices from the bitmap; CLEAR: zero it for the next use ------
template <typename OutT, bool CLEAR>
__global__ void __launch_bounds__(256) bitmap_emit_kernel(uint32_t* __restrict__ bitmap, int64_t nwords,
                                                          const int32_t* __restrict__ blocksum,
                                                          OutT* __restrict__ uniq) {
    __shared__ int s_w[33];
    const int64_t w0 = (int64_t)blockIdx.x * TILE + threadIdx.x * 4;
    uint32_t w[4];
    int c = 0;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        w[q] = (w0 + q < nwords) ? bitmap[w0 + q] : 0u;
        c += __popc(w[q]);
    }
    int total;
    int off = blocksum[blockIdx.x] + block_excl_scan<256>(c, s_w, total);
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        uint32_t m = w[q];
        if (CLEAR && m) bitmap[w0 + q] = 0u;
        while (m) {
            const int b = __ffs(m) - 1;
            m &= m - 1;
            uniq[off++] = (OutT)((w0 + q) * 32 + b);
        }
    }
}


}  // namespace
