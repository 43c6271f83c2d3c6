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

__global__ void __launch_bounds__(256) synth_ids_kernel(SynthTables tb, int t0, uint64_t seed, int64_t batch_global,
                                                        int64_t step0, int n_steps, int64_t b0, int nb, int uniform,
                                                        double inv_e, int64_t* __restrict__ out, int64_t ld) {
    const int k = t0 + blockIdx.y;
    const int64_t total = (int64_t)n_steps * nb;
    const int64_t n = tb.n[k];
    const double c0 = tb.c0[k];
    const uint64_t key = splitmix64(seed ^ (0x51ed270b1f2d3a4full * (uint64_t)(k + 1)));
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t s = i / nb, b = i - s * nb;
        const uint64_t ctr = (uint64_t)((step0 + s) * batch_global + b0 + b);
        const double u = (double)(splitmix64(key + ctr) >> 11) * (1.0 / 9007199254740992.0);
        int64_t r;
        if (uniform || n == 1) {
            r = (int64_t)(u * (double)n);
        } else {
            r = (int64_t)pow(c0 * u + 1.0, inv_e) - 1;
        }
        r = r < 0 ? 0 : (r >= n ? n - 1 : r);
        out[(int64_t)blockIdx.y * ld + i] = (int64_t)(((uint64_t)r * 2654435761ull + 40503ull * (uint64_t)k) % (uint64_t)n);
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
    const int64_t total = (int64_t)n_steps * nb;
    int64_t gx = (total + 255) / 256;
    if (gx > 148 * 8) gx = 148 * 8;
    cudaStream_t s = (cudaStream_t)stream;
    LAUNCH(K_MISC, s, synth_ids_kernel<<<dim3((unsigned)gx, table_count), 256, 0, s>>>(
        tb, table_begin, seed, batch_global, step0, n_steps, b0, nb, uniform, uniform ? 0.0 : 1.0 / e, out, ld));
    CU_CHECK(cudaGetLastError());
    return CDLRM_OK;
}
