// interact.cu -- pairwise-dot feature interaction, forward and backward
// (DLRM_Net.interact_features "dot", model_no_ddp.py:272-293 of the reference).
//
// FP32 on CUDA cores: the loss-parity bar is 1e-5 relative, which TF32 tensor-core
// products (10-bit mantissa) cannot meet, and the op sits at ~12 flop/byte -- below
// the FP32 ridge -- so it is HBM-bound once the dot products stay in registers.
//
// Forward (register-resident, no shared memory): LPS = dim/4 lanes own one sample;
// lane c holds the float4 column slice c of all F feature rows (F*4 registers).  The
// strict lower triangle is walked in the reference's row-major pair order; every
// block of LPS pairs is reduced across the LPS lanes with a halving butterfly
// (LPS-1 shuffles per LPS pairs), after which lane c owns pair (block*LPS + c) and the
// block is written with one coalesced store.
//
// Backward: lanes are independent along dim (no reduction), so TPS = dim/VEC threads
// own one sample with VEC = 2 to keep 2*F*VEC accumulators + operands in registers;
// the dZ coefficients are loaded coalesced (one per lane) and broadcast by shuffle.
#include <type_traits>
#include <utility>

#include "common.cuh"

namespace {

constexpr int MAX_FEAT = 64;

struct FeatPtrs {
    const float* p[MAX_FEAT];
};

template <int F, bool ITSELF>
struct Pairs {
    static constexpr int N = ITSELF ? F * (F + 1) / 2 : F * (F - 1) / 2;
};

__device__ __forceinline__ float dot4(const float4& a, const float4& b) {
    return fmaf(a.w, b.w, fmaf(a.z, b.z, fmaf(a.y, b.y, a.x * b.x)));
}

// compile-time loops: every index into the register arrays below must be a constant,
// otherwise the arrays fall into local memory (that cost 12x in the first version)
template <int... Is, class Fn>
__device__ __forceinline__ void static_for_impl(std::integer_sequence<int, Is...>, Fn&& f) {
    (f(std::integral_constant<int, Is>{}), ...);
}
template <int N, class Fn>
__device__ __forceinline__ void static_for(Fn&& f) {
    static_for_impl(std::make_integer_sequence<int, N>{}, static_cast<Fn&&>(f));
}

template <int OFF, int LPS>
__device__ __forceinline__ void butterfly_level(float (&v)[LPS], int sub) {
    if constexpr (OFF >= 1) {
        const bool hi = (sub & OFF) != 0;
        static_for<OFF>([&](auto Q) {
            constexpr int q = decltype(Q)::value;
            const float send = hi ? v[q] : v[q + OFF];
            const float keep = hi ? v[q + OFF] : v[q];
            v[q] = keep + __shfl_xor_sync(0xffffffffu, send, OFF);
        });
        butterfly_level<OFF / 2, LPS>(v, sub);
    }
}

// after the call lane `sub` holds sum over the LPS lanes of v[sub]
template <int LPS>
__device__ __forceinline__ float butterfly(float (&v)[LPS], int sub) {
    butterfly_level<LPS / 2, LPS>(v, sub);
    return v[0];
}

template <int F, int LPS, bool ITSELF>
__global__ void __launch_bounds__(128) interact_fwd_kernel(FeatPtrs fp, int64_t row_stride, int B,
                                                           float* __restrict__ out, int64_t ld_out) {
    constexpr int DIM = LPS * 4;
    constexpr int SPW = 32 / LPS;  // samples per warp
    constexpr int NP = Pairs<F, ITSELF>::N;
    const int lane = threadIdx.x & 31;
    const int sub = lane % LPS;
    const int gwarp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int b = gwarp * SPW + lane / LPS;
    const bool live = b < B;
    float4 t[F];
    static_for<F>([&](auto I) {
        constexpr int i = decltype(I)::value;
        t[i] = live ? reinterpret_cast<const float4*>(fp.p[i] + (int64_t)b * row_stride)[sub]
                    : make_float4(0.f, 0.f, 0.f, 0.f);
    });
    float* orow = out + (int64_t)b * ld_out;
    if (live) {  // dense features pass through (model_no_ddp.py:293)
        orow[sub * 4 + 0] = t[0].x; orow[sub * 4 + 1] = t[0].y;
        orow[sub * 4 + 2] = t[0].z; orow[sub * 4 + 3] = t[0].w;
    }
    float v[LPS];
    static_for<F>([&](auto I) {
        constexpr int i = decltype(I)::value;
        constexpr int nj = ITSELF ? i + 1 : i;
        constexpr int p0 = ITSELF ? i * (i + 1) / 2 : i * (i - 1) / 2;   // row-major triangle offset
        static_for<nj>([&](auto J) {
            constexpr int j = decltype(J)::value;
            constexpr int p = p0 + j;
            v[p % LPS] = dot4(t[i], t[j]);
            if constexpr ((p + 1) % LPS == 0 || p + 1 == NP) {
                constexpr int blk = p / LPS;
                constexpr int cnt = p + 1 - blk * LPS;
                static_for<LPS>([&](auto Q) {
                    constexpr int q = decltype(Q)::value;
                    if constexpr (q >= cnt) v[q] = 0.f;
                });
                const float r = butterfly<LPS>(v, sub);
                if (live && sub < cnt) orow[DIM + blk * LPS + sub] = r;
            }
        });
    });
}

template <int F, int TPS, bool ITSELF>
__global__ void __launch_bounds__(128) interact_bwd_kernel(FeatPtrs fp, int64_t row_stride, int B,
                                                           const float* __restrict__ d_out, int64_t ld_dout,
                                                           float* __restrict__ d_feat, int64_t ld_dfeat) {
    // TPS threads per sample, each owning a float2 column slice: dim = 2*TPS.  The NP pair
    // coefficients of a sample are staged once in shared memory (coalesced) and read back as
    // broadcast float4 loads -- one LDS.128 per four pairs instead of one shuffle per pair.
    constexpr int DIM = TPS * 2;
    constexpr int SPB = 128 / TPS;                  // samples per block
    constexpr int NP = Pairs<F, ITSELF>::N;
    constexpr int NPP = (NP + 3) & ~3;
    __shared__ __align__(16) float s_c[SPB][NPP];
    const int ls = threadIdx.x / TPS;               // sample within the block
    const int col = threadIdx.x % TPS;
    const int b = blockIdx.x * SPB + ls;
    const bool live = b < B;
    for (int e = threadIdx.x; e < SPB * NPP; e += 128) {
        const int sm = e / NPP, p = e - sm * NPP;
        const int bb = blockIdx.x * SPB + sm;
        s_c[sm][p] = (bb < B && p < NP) ? d_out[(int64_t)bb * ld_dout + DIM + p] : 0.f;
    }
    float2 t[F], g[F];
    static_for<F>([&](auto I) {
        constexpr int i = decltype(I)::value;
        t[i] = live ? reinterpret_cast<const float2*>(fp.p[i] + (int64_t)b * row_stride)[col] : make_float2(0.f, 0.f);
        g[i] = make_float2(0.f, 0.f);
    });
    const float* drow = d_out + (int64_t)b * ld_dout;
    if (live) g[0] = make_float2(drow[col * 2], drow[col * 2 + 1]);  // d/dx of the pass-through (rows may be odd-sized)
    __syncthreads();
    const float4* c4p = reinterpret_cast<const float4*>(&s_c[ls][0]);
    float4 cq = make_float4(0.f, 0.f, 0.f, 0.f);
    static_for<F>([&](auto I) {
        constexpr int i = decltype(I)::value;
        constexpr int nj = ITSELF ? i + 1 : i;
        constexpr int p0 = ITSELF ? i * (i + 1) / 2 : i * (i - 1) / 2;
        static_for<nj>([&](auto J) {
            constexpr int j = decltype(J)::value;
            constexpr int p = p0 + j;
            if constexpr (p % 4 == 0) cq = c4p[p / 4];
            const float c = (p % 4 == 0) ? cq.x : (p % 4 == 1) ? cq.y : (p % 4 == 2) ? cq.z : cq.w;
            if constexpr (i == j) {  // d(T_i.T_i) = 2 T_i
                g[i].x = fmaf(2.f * c, t[i].x, g[i].x);
                g[i].y = fmaf(2.f * c, t[i].y, g[i].y);
            } else {
                g[i].x = fmaf(c, t[j].x, g[i].x);
                g[i].y = fmaf(c, t[j].y, g[i].y);
                g[j].x = fmaf(c, t[i].x, g[j].x);
                g[j].y = fmaf(c, t[i].y, g[j].y);
            }
        });
    });
    if (live) {
        static_for<F>([&](auto I) {
            constexpr int i = decltype(I)::value;
            reinterpret_cast<float2*>(d_feat + (int64_t)i * ld_dfeat + (int64_t)b * DIM)[col] = g[i];
        });
    }
}

// ---- generic fallbacks (any F <= MAX_FEAT, any dim): correctness path ---------------
__global__ void interact_fwd_generic(FeatPtrs fp, int F, int64_t row_stride, int B, int dim, int itself,
                                     float* __restrict__ out, int64_t ld_out) {
    const int b = blockIdx.x;
    const int np = itself ? F * (F + 1) / 2 : F * (F - 1) / 2;
    float* orow = out + (int64_t)b * ld_out;
    for (int c = threadIdx.x; c < dim; c += blockDim.x) orow[c] = fp.p[0][(int64_t)b * row_stride + c];
    for (int p = threadIdx.x; p < np; p += blockDim.x) {
        // invert the row-major triangle index
        int i = itself ? 0 : 1, acc = 0;
        while (true) {
            int len = itself ? i + 1 : i;
            if (p < acc + len) break;
            acc += len;
            ++i;
        }
        int j = p - acc;
        const float* a = fp.p[i] + (int64_t)b * row_stride;
        const float* bb = fp.p[j] + (int64_t)b * row_stride;
        float s = 0.f;
        for (int c = 0; c < dim; ++c) s = fmaf(a[c], bb[c], s);
        orow[dim + p] = s;
    }
}

__global__ void interact_bwd_generic(FeatPtrs fp, int F, int64_t row_stride, int B, int dim, int itself,
                                     const float* __restrict__ d_out, int64_t ld_dout,
                                     float* __restrict__ d_feat, int64_t ld_dfeat) {
    const int b = blockIdx.x;
    const float* drow = d_out + (int64_t)b * ld_dout;
    for (int e = threadIdx.x; e < F * dim; e += blockDim.x) {
        const int i = e / dim, c = e % dim;
        float s = (i == 0) ? drow[c] : 0.f;
        for (int j = 0; j < F; ++j) {
            float coef;
            if (j < i) coef = drow[dim + (itself ? i * (i + 1) / 2 : i * (i - 1) / 2) + j];
            else if (j > i) coef = drow[dim + (itself ? j * (j + 1) / 2 : j * (j - 1) / 2) + i];
            else if (itself) coef = 2.f * drow[dim + i * (i + 1) / 2 + i];
            else continue;
            s = fmaf(coef, fp.p[j][(int64_t)b * row_stride + c], s);
        }
        d_feat[(int64_t)i * ld_dfeat + (int64_t)b * dim + c] = s;
    }
}

template <int F, int LPS, bool ITSELF>
void launch_fwd(const FeatPtrs& fp, int64_t rs, int B, float* out, int64_t ld_out, cudaStream_t s) {
    constexpr int SPW = 32 / LPS;
    const int warps = (B + SPW - 1) / SPW;
    const int blocks = (warps + 3) / 4;
    LAUNCH(K_INT_FWD, s, (interact_fwd_kernel<F, LPS, ITSELF><<<blocks, 128, 0, s>>>(fp, rs, B, out, ld_out)));
}

template <int F, int TPS, bool ITSELF>
void launch_bwd(const FeatPtrs& fp, int64_t rs, int B, const float* d_out, int64_t ld_dout, float* d_feat,
                int64_t ld_dfeat, cudaStream_t s) {
    const int64_t threads = (int64_t)B * TPS;
    const int blocks = (int)((threads + 127) / 128);
    LAUNCH(K_INT_BWD, s, (interact_bwd_kernel<F, TPS, ITSELF><<<blocks, 128, 0, s>>>(fp, rs, B, d_out, ld_dout, d_feat, ld_dfeat)));
}

bool aligned_for(const FeatPtrs& fp, int n, int64_t rs, int bytes) {
    for (int i = 0; i < n; ++i)
        if ((uintptr_t)fp.p[i] % bytes) return false;
    return (rs * 4) % bytes == 0;
}

}  // namespace

#define FWD_CASE(F_, D_)                                                                  \
    if (n_feat == F_ && dim == D_ && !itself && fast16) {                                 \
        launch_fwd<F_, D_ / 4, false>(fp, rs, batch, out, ld_out, s);                      \
        done = true;                                                                      \
    }
#define BWD_CASE(F_, D_)                                                                  \
    if (n_feat == F_ && dim == D_ && !itself && fast8) {                                  \
        launch_bwd<F_, D_ / 2, false>(fp, rs, batch, d_out, ld_dout, d_feat, ld_dfeat, s); \
        done = true;                                                                      \
    }

extern "C" int cdlrm_interact_fwd(int device, const float* const* h_feat, int n_feat, int64_t rs, int32_t batch,
                                  int dim, int itself, float* out, int64_t ld_out, cdlrm_stream stream) {
    ARG_CHECK(h_feat && out);
    ARG_CHECK(n_feat >= 1 && n_feat <= MAX_FEAT && dim >= 1 && batch >= 0);
    const int np = itself ? n_feat * (n_feat + 1) / 2 : n_feat * (n_feat - 1) / 2;
    ARG_CHECK(ld_out >= dim + np && rs >= dim);
    if (batch == 0) return CDLRM_OK;
    cudaStream_t s = (cudaStream_t)stream;
    CU_CHECK(cudaSetDevice(device));
    FeatPtrs fp;
    for (int i = 0; i < n_feat; ++i) {
        ARG_CHECK(h_feat[i]);
        fp.p[i] = h_feat[i];
    }
    const bool fast16 = aligned_for(fp, n_feat, rs, 16);
    bool done = false;
    FWD_CASE(27, 128) FWD_CASE(27, 64) FWD_CASE(27, 32) FWD_CASE(27, 16)
    FWD_CASE(9, 128) FWD_CASE(9, 64) FWD_CASE(9, 32) FWD_CASE(9, 16)
    if (!done) LAUNCH(K_INT_FWD, s, interact_fwd_generic<<<batch, 128, 0, s>>>(fp, n_feat, rs, batch, dim, itself, out, ld_out));
    CU_CHECK(cudaGetLastError());
    return CDLRM_OK;
}

extern "C" int cdlrm_interact_bwd(int device, const float* const* h_feat, int n_feat, int64_t rs, int32_t batch,
                                  int dim, int itself, const float* d_out, int64_t ld_dout, float* d_feat,
                                  int64_t ld_dfeat, cdlrm_stream stream) {
    ARG_CHECK(h_feat && d_out && d_feat);
    ARG_CHECK(n_feat >= 1 && n_feat <= MAX_FEAT && dim >= 1 && batch >= 0);
    const int np = itself ? n_feat * (n_feat + 1) / 2 : n_feat * (n_feat - 1) / 2;
    ARG_CHECK(ld_dout >= dim + np && rs >= dim && ld_dfeat >= (int64_t)batch * dim);
    if (batch == 0) return CDLRM_OK;
    cudaStream_t s = (cudaStream_t)stream;
    CU_CHECK(cudaSetDevice(device));
    FeatPtrs fp;
    for (int i = 0; i < n_feat; ++i) {
        ARG_CHECK(h_feat[i]);
        fp.p[i] = h_feat[i];
    }
    const bool fast8 = aligned_for(fp, n_feat, rs, 8) && ((uintptr_t)d_feat % 8 == 0) && (ld_dfeat % 2 == 0);
    bool done = false;
    BWD_CASE(27, 128) BWD_CASE(27, 64) BWD_CASE(27, 32) BWD_CASE(27, 16)
    BWD_CASE(9, 128) BWD_CASE(9, 64) BWD_CASE(9, 32) BWD_CASE(9, 16)
    if (!done)
        LAUNCH(K_INT_BWD, s, interact_bwd_generic<<<batch, 128, 0, s>>>(fp, n_feat, rs, batch, dim, itself, d_out, ld_dout, d_feat, ld_dfeat));
    CU_CHECK(cudaGetLastError());
    return CDLRM_OK;
}
