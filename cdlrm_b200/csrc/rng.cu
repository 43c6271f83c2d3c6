// rng.cu -- host-side victim-way random stream, bit-compatible with the torch CPU
// generator the reference consumes through torch.distributions.Categorical(...).sample()
// (main_no_ddp.py:183-185): torch.multinomial's fast path draws
// q = empty(rows, ways).exponential_(1) and takes argmax(probs / q).
//   exponential_(1) on float32 == float32(-log1p(-u)),  u = (r64 & (2^53-1)) * 2^-53,
//   r64 = (mt32() << 32) | mt32(),  mt19937 seeded with init_genrand(seed).
// mt19937 is sequential, so the raw stream is produced by the calling thread while a
// small pool of threads applies the (expensive) log1p transform block by block.
#include <math.h>

#include <random>
#include <thread>

#include "common.cuh"

struct cdlrm_rng {
    std::mt19937 gen;
    uint64_t draws = 0;
    std::vector<uint64_t> raw[2];
};

static inline float transform(uint64_t r) {
    const double u = (double)(r & ((1ull << 53) - 1ull)) * (1.0 / 9007199254740992.0);
    return (float)(-log1p(-u));
}

extern "C" int cdlrm_rng_create(cdlrm_rng** out, uint64_t seed) {
    ARG_CHECK(out);
    cdlrm_rng* r = new cdlrm_rng();
    r->gen.seed((uint32_t)(seed & 0xffffffffull));  // torch: mt19937(seed) truncates to 32 bits
    *out = r;
    return CDLRM_OK;
}

extern "C" int cdlrm_rng_destroy(cdlrm_rng* r) {
    delete r;
    return CDLRM_OK;
}

extern "C" uint64_t cdlrm_rng_draws(const cdlrm_rng* r) { return r ? r->draws : 0; }

extern "C" int cdlrm_rng_exponential(cdlrm_rng* r, float* out, int64_t n, int threads) {
    ARG_CHECK(r && n >= 0 && (out || n == 0));
    if (threads < 1) threads = 1;
    constexpr int64_t BLOCK = 1 << 20;
    if (n < (1 << 16) || threads == 1) {
        for (int64_t i = 0; i < n; ++i) {
            uint64_t hi = r->gen(), lo = r->gen();
            out[i] = transform((hi << 32) | lo);
        }
        r->draws += (uint64_t)n;
        return CDLRM_OK;
    }
    r->raw[0].resize(BLOCK);
    r->raw[1].resize(BLOCK);
    auto fill = [&](int buf, int64_t cnt) {
        uint64_t* p = r->raw[buf].data();
        for (int64_t i = 0; i < cnt; ++i) {
            uint64_t hi = r->gen(), lo = r->gen();
            p[i] = (hi << 32) | lo;
        }
    };
    const int64_t nblk = (n + BLOCK - 1) / BLOCK;
    fill(0, n < BLOCK ? n : BLOCK);
    for (int64_t b = 0; b < nblk; ++b) {
        const int64_t base = b * BLOCK;
        const int64_t cnt = n - base < BLOCK ? n - base : BLOCK;
        const uint64_t* src = r->raw[b & 1].data();
        float* dst = out + base;
        std::vector<std::thread> pool;
        const int nt = threads - 1 > 0 ? threads - 1 : 1;
        const int64_t per = (cnt + nt - 1) / nt;
        for (int t = 0; t < nt; ++t) {
            const int64_t lo = t * per, hi = lo + per < cnt ? lo + per : cnt;
            if (lo >= hi) break;
            pool.emplace_back([=]() {
                for (int64_t i = lo; i < hi; ++i) dst[i] = transform(src[i]);
            });
        }
        if (b + 1 < nblk) {  // overlap: produce the next raw block while the pool transforms this one
            const int64_t nb = (b + 1) * BLOCK;
            fill((b + 1) & 1, n - nb < BLOCK ? n - nb : BLOCK);
        }
        for (auto& th : pool) th.join();
    }
    r->draws += (uint64_t)n;
    return CDLRM_OK;
}

// =================================================================================================
// Device-resident victim stream.  The same mt19937 sequence, generated ON the GPU by one CTA:
// the recurrence x[k+624] = x[k+397] ^ f(x[k], x[k+1]) leaves 227 / 227 / 170 independent
// elements per regeneration, so a 256-thread CTA refreshes the 624-word state in three
// barrier-separated phases (~4 G words/s) -- an order of magnitude faster than a host core and
// with no PCIe traffic for the draws.  The exponential transform (expdraw.cuh) is applied by
// whoever consumes the raw words (the planner's select kernel, or the kernel below).
// =================================================================================================
#include "expdraw.cuh"

struct cdlrm_rngdev {
    int device = 0;
    uint32_t* d_state = nullptr;  // x[624], pos
    uint64_t draws = 0;
};

namespace {

constexpr int MT_N = 624, MT_M = 397;

__device__ __forceinline__ uint32_t mt_twist(uint32_t a, uint32_t b, uint32_t c) {
    const uint32_t y = (a & 0x80000000u) | (b & 0x7fffffffu);
    return c ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
}

__device__ __forceinline__ uint32_t mt_temper(uint32_t y) {
    y ^= y >> 11;
    y ^= (y << 7) & 0x9d2c5680u;
    y ^= (y << 15) & 0xefc60000u;
    y ^= y >> 18;
    return y;
}

__global__ void __launch_bounds__(256) mt_generate_kernel(uint32_t* __restrict__ state, uint32_t* __restrict__ out,
                                                          long long n_words) {
    __shared__ uint32_t x[MT_N];
    const int tid = threadIdx.x;
    for (int i = tid; i < MT_N; i += 256) x[i] = state[i];
    int pos = (int)state[MT_N];
    __syncthreads();
    long long done = 0;
    while (done < n_words) {
        if (pos == MT_N) {  // uniform across the CTA
            uint32_t v = 0;
            if (tid < 227) v = mt_twist(x[tid], x[tid + 1], x[tid + MT_M]);
            __syncthreads();
            if (tid < 227) x[tid] = v;
            __syncthreads();
            if (tid < 227) v = mt_twist(x[227 + tid], x[228 + tid], x[tid]);
            __syncthreads();
            if (tid < 227) x[227 + tid] = v;
            __syncthreads();
            if (tid < 170) v = mt_twist(x[454 + tid], x[tid == 169 ? 0 : 455 + tid], x[227 + tid]);
            __syncthreads();
            if (tid < 170) x[454 + tid] = v;
            __syncthreads();
            pos = 0;
        }
        const long long left = n_words - done;
        const int cnt = (MT_N - pos) < left ? (MT_N - pos) : (int)left;
        for (int i = tid; i < cnt; i += 256) out[done + i] = mt_temper(x[pos + i]);
        done += cnt;
        pos += cnt;
        __syncthreads();  // emit reads of x[] complete before the next regeneration writes
    }
    for (int i = tid; i < MT_N; i += 256) state[i] = x[i];
    if (tid == 0) state[MT_N] = (uint32_t)pos;
}

__global__ void exp_from_raw_kernel(const uint2* __restrict__ raw, float* __restrict__ out, long long n) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (; i < n; i += stride) out[i] = exp_draw_from_raw(raw[i]);
}

}  // namespace

extern "C" int cdlrm_rngdev_create(cdlrm_rngdev** out, int device, uint64_t seed) {
    ARG_CHECK(out);
    CU_CHECK(cudaSetDevice(device));
    cdlrm_rngdev* r = new cdlrm_rngdev();
    r->device = device;
    uint32_t h[MT_N + 1];
    h[0] = (uint32_t)(seed & 0xffffffffull);  // init_genrand, as torch.manual_seed(seed) does
    for (int i = 1; i < MT_N; ++i) h[i] = 1812433253u * (h[i - 1] ^ (h[i - 1] >> 30)) + (uint32_t)i;
    h[MT_N] = MT_N;  // first use regenerates
    CU_CHECK(cudaMalloc(&r->d_state, sizeof(h)));
    CU_CHECK(cudaMemcpy(r->d_state, h, sizeof(h), cudaMemcpyHostToDevice));
    *out = r;
    return CDLRM_OK;
}

extern "C" int cdlrm_rngdev_destroy(cdlrm_rngdev* r) {
    if (!r) return CDLRM_OK;
    cudaSetDevice(r->device);
    cudaFree(r->d_state);
    delete r;
    return CDLRM_OK;
}

extern "C" uint64_t cdlrm_rngdev_draws(const cdlrm_rngdev* r) { return r ? r->draws : 0; }

extern "C" int cdlrm_rngdev_raw(cdlrm_rngdev* r, uint32_t* d_out, int64_t n_draws, cdlrm_stream stream) {
    ARG_CHECK(r && n_draws >= 0);
    if (n_draws == 0) return CDLRM_OK;
    ARG_CHECK(d_out && ((uintptr_t)d_out & 7) == 0);
    CU_CHECK(cudaSetDevice(r->device));
    cudaStream_t s = (cudaStream_t)stream;
    LAUNCH(K_RNG_MT, s, mt_generate_kernel<<<1, 256, 0, s>>>(r->d_state, d_out, (long long)n_draws * 2));
    CU_CHECK(cudaGetLastError());
    r->draws += (uint64_t)n_draws;
    return CDLRM_OK;
}

extern "C" int cdlrm_exp_from_raw(const uint32_t* d_raw, float* d_out, int64_t n, cdlrm_stream stream) {
    ARG_CHECK(n >= 0);
    if (n == 0) return CDLRM_OK;
    ARG_CHECK(d_raw && d_out && ((uintptr_t)d_raw & 7) == 0);
    cudaStream_t s = (cudaStream_t)stream;
    int64_t blocks = (n + 255) / 256;
    if (blocks > 148 * 16) blocks = 148 * 16;
    LAUNCH(K_RNG_EXP, s, exp_from_raw_kernel<<<(int)blocks, 256, 0, s>>>(reinterpret_cast<const uint2*>(d_raw), d_out, (long long)n));
    CU_CHECK(cudaGetLastError());
    return CDLRM_OK;
}

extern "C" int cdlrm_rngdev_exponential(cdlrm_rngdev* r, float* d_out, int64_t n, uint32_t* d_raw_scratch,
                                        cdlrm_stream stream) {
    ARG_CHECK(r && n >= 0);
    if (n == 0) return CDLRM_OK;
    ARG_CHECK(d_out && d_raw_scratch);
    int rc = cdlrm_rngdev_raw(r, d_raw_scratch, n, stream);
    if (rc) return rc;
    return cdlrm_exp_from_raw(d_raw_scratch, d_out, n, stream);
}
