// synth.cu -- Criteo-shaped synthetic sparse ids generated on the device (the synthetic loader and its
// cache_ld twin; the reference's own random mode, main_no_ddp.py:539-547, builds no cache loader and cannot run).
//
// One counter-based stream per table: the id of (table k, global step s, sample b of the global batch) is a pure
// function of (seed, k, s, b).  Any rank can therefore generate any slice of any window -- its own training batches,
// or chunks of the GLOBAL window for the look-ahead planner's bitmap scan -- without ever holding the whole
// [T, lookahead x global batch] int64 window in memory (41 GB at 8 GPUs), and the train and look-ahead views of
// the stream agree by construction.
//   u  = 53 random bits of splitmix64(seed, k, s * Bg + b) / 2^53
//   r  = uniform: floor(u n);  power law (exponent a): floor((((n+1)^(1-a) - 1) u + 1)^(1/(1-a))) - 1
//   id = (r * 2654435761 + 40503 k) mod n          (scrambles the ranks over the id space)
#include "common.cuh"

namespace {

struct SynthTables {
    int64_t n[64];
    double c0[64];      // (n+1)^(1-a) - 1
};

__device__ __forceinline__ uint64_t splitmix64(uint64_t x) {
    x += 0x9e3779b97f4a7c15ull;
    x = (x ^ (x >> 30)) * 0xbf58476d1ce4e5b9ull;
    x = (x ^ (x >> 27)) * 0x94d049bb133111ebull;
    return x ^ (x >> 31);
}

// (r * 2654435761 + c) mod n without a 64-bit division: 32-bit arithmetic for small n, a double-precision quotient
// estimate (off by at most one: x < 2^58, n >= 2^16) with an exact integer remainder otherwise
__device__ __forceinline__ int64_t scramble_mod(uint64_t r, uint64_t c, uint64_t n, double inv_n) {
    if (n < 65536ull) {
        const uint32_t n32 = (uint32_t)n;
        return (int64_t)(((uint32_t)r * (uint32_t)(2654435761ull % n) + (uint32_t)(c % n)) % n32);
    }
    const uint64_t x = r * 2654435761ull + c;
    const uint64_t q = (uint64_t)((double)x * inv_n);
    int64_t rem = (int64_t)(x - q * n);
    if (rem < 0) rem += (int64_t)n;
    else if (rem >= (int64_t)n) rem -= (int64_t)n;
    return rem;
}

// One thread per (step, sample) element and SHORT CTAs (no grid-stride loop): the stream is generated on a side stream
// beside training, and a low-priority grid of long-running CTAs keeps the SM slots it once got (a training kernel's
// CTAs then wait for them), whereas short CTAs hand the slots back within microseconds.  blockIdx.y = step chunk,
// blockIdx.z = table: no 64-bit division per element.
constexpr int SYNTH_STEPS_PER_CTA = 4;
__global__ void __launch_bounds__(256) synth_ids_kernel(SynthTables tb, int t0, uint64_t seed, int64_t batch_global,
                                                        int64_t step0, int n_steps, int64_t b0, int nb, int uniform,
                                                        double inv_e, int64_t* __restrict__ out, int64_t ld) {
    const int k = t0 + blockIdx.z;
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= nb) return;
    const int64_t n = tb.n[k];
    const double c0 = tb.c0[k], inv_n = 1.0 / (double)n;
    const uint64_t key = splitmix64(seed ^ (0x51ed270b1f2d3a4full * (uint64_t)(k + 1)));
    const int s_lo = blockIdx.y * SYNTH_STEPS_PER_CTA;
    const int s_hi = min(n_steps, s_lo + SYNTH_STEPS_PER_CTA);
    for (int s = s_lo; s < s_hi; ++s) {
        const uint64_t ctr = (uint64_t)((step0 + s) * batch_global + b0 + b);
        const double u = (double)(splitmix64(key + ctr) >> 11) * (1.0 / 9007199254740992.0);
        int64_t r;
        if (uniform || n == 1) {
            r = (int64_t)(u * (double)n);
        } else {
            r = (int64_t)exp2(inv_e * log2(c0 * u + 1.0)) - 1;      // x^(1/(1-a)); ~half the cost of pow(), error << 1 rank
        }
        r = r < 0 ? 0 : (r >= n ? n - 1 : r);
        out[(int64_t)blockIdx.z * ld + (int64_t)s * nb + b] = scramble_mod((uint64_t)r, 40503ull * (uint64_t)k, (uint64_t)n, inv_n);
    }
}

}  // namespace

extern "C" int cdlrm_synth_ids(int device, int table_begin, int table_count, const int64_t* h_n_rows, uint64_t seed,
                               int64_t batch_global, int64_t step0, int32_t n_steps, int64_t b0, int32_t nb,
                               int uniform, double zipf_a, int64_t* out, int64_t ld, cdlrm_stream stream) {
    ARG_CHECK(h_n_rows && out && table_begin >= 0 && table_count > 0 && table_begin + table_count <= 64);
    ARG_CHECK(batch_global > 0 && step0 >= 0 && n_steps >= 0 && b0 >= 0 && nb >= 0 && b0 + nb <= batch_global);
    ARG_CHECK(ld >= (int64_t)n_steps * nb);
    ARG_CHECK(uniform || zipf_a != 1.0);
    if (n_steps == 0 || nb == 0) return CDLRM_OK;
    CU_CHECK(cudaSetDevice(device));
    SynthTables tb;
    const double e = 1.0 - zipf_a;
    for (int k = table_begin; k < table_begin + table_count; ++k) {
        ARG_CHECK(h_n_rows[k - table_begin] > 0);
        tb.n[k] = h_n_rows[k - table_begin];
        tb.c0[k] = uniform ? 0.0 : pow((double)tb.n[k] + 1.0, e) - 1.0;
    }
    cudaStream_t s = (cudaStream_t)stream;
    const int chunks = (n_steps + SYNTH_STEPS_PER_CTA - 1) / SYNTH_STEPS_PER_CTA;
    for (int c0 = 0; c0 < chunks; c0 += 65535) {          // gridDim.y limit
        const int cy = chunks - c0 < 65535 ? chunks - c0 : 65535;
        const int s0 = c0 * SYNTH_STEPS_PER_CTA;
        const int ns = n_steps - s0 < cy * SYNTH_STEPS_PER_CTA ? n_steps - s0 : cy * SYNTH_STEPS_PER_CTA;
        LAUNCH(K_MISC, s, synth_ids_kernel<<<dim3((unsigned)((nb + 255) / 256), (unsigned)cy, (unsigned)table_count), 256, 0, s>>>(
            tb, table_begin, seed, batch_global, step0 + s0, ns, b0, nb, uniform, uniform ? 0.0 : 1.0 / e,
            out + (int64_t)s0 * nb, ld));
    }
    CU_CHECK(cudaGetLastError());
    return CDLRM_OK;
}
