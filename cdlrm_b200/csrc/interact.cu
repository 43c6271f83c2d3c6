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
#include <cstdlib>
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

// packed FP32 (sm_100 FFMA2 / FMUL2): one issue slot for two lanes of work -- these kernels are issue-bound
__device__ __forceinline__ float dot4_packed(const float4& a, const float4& b) {
    float2 p = __fmul2_rn(make_float2(a.x, a.y), make_float2(b.x, b.y));
    p = __ffma2_rn(make_float2(a.z, a.w), make_float2(b.z, b.w), p);
    return p.x + p.y;
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
    pdl_enter();
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
            v[p % LPS] = dot4_packed(t[i], t[j]);
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

// Forward for dim = 128 (32 lanes own a sample, lane c the float4 column slice c).  Same walk over the
// triangle as above, but a block of 32 partial dot products per lane is reduced across the lanes through
// shared memory: every lane stores its 32 partials as one row (8 STS.128, row pitch 36 words: conflict
// free), then lane c adds up column c (32 LDS + 31 FADD).  71 instructions per 32 pairs instead of the
// butterfly's 31 x (2 SEL + SHFL + FADD) = 124; the rows are double-buffered so that one __syncwarp per
// block is enough.
template <int F, bool ITSELF>
__global__ void __launch_bounds__(128, 4) interact_fwd_tr_kernel(FeatPtrs fp, int64_t row_stride, int B,
                                                              float* __restrict__ out, int64_t ld_out) {
    pdl_enter();
    constexpr int DIM = 128, PITCH = 36;
    constexpr int NP = Pairs<F, ITSELF>::N;
    __shared__ __align__(16) float s_part[4][2][32 * PITCH];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const int b = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (b >= B) return;                                   // warp-uniform
    float4 t[F];
    static_for<F>([&](auto I) {
        constexpr int i = decltype(I)::value;
        t[i] = __ldg(reinterpret_cast<const float4*>(fp.p[i] + (int64_t)b * row_stride) + lane);
    });
    float* orow = out + (int64_t)b * ld_out;
    orow[lane * 4 + 0] = t[0].x; orow[lane * 4 + 1] = t[0].y;      // dense features pass through (model_no_ddp.py:293)
    orow[lane * 4 + 2] = t[0].z; orow[lane * 4 + 3] = t[0].w;
    float v[32];
    static_for<F>([&](auto I) {
        constexpr int i = decltype(I)::value;
        constexpr int nj = ITSELF ? i + 1 : i;
        constexpr int p0 = ITSELF ? i * (i + 1) / 2 : i * (i - 1) / 2;   // row-major triangle offset
        static_for<nj>([&](auto J) {
            constexpr int j = decltype(J)::value;
            constexpr int p = p0 + j;
            v[p % 32] = dot4_packed(t[i], t[j]);
            if constexpr ((p + 1) % 32 == 0 || p + 1 == NP) {
                constexpr int blk = p / 32;
                constexpr int cnt = p + 1 - blk * 32;
                float* buf = s_part[wib][blk & 1];
                float4* wr = reinterpret_cast<float4*>(buf + lane * PITCH);
                static_for<(cnt + 3) / 4>([&](auto Q) {
                    constexpr int q = decltype(Q)::value;
                    wr[q] = make_float4(v[4 * q], 4 * q + 1 < cnt ? v[4 * q + 1] : 0.f, 4 * q + 2 < cnt ? v[4 * q + 2] : 0.f,
                                        4 * q + 3 < cnt ? v[4 * q + 3] : 0.f);
                });
                __syncwarp();
                if (lane < cnt) {
                    float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
#pragma unroll
                    for (int r = 0; r < 32; r += 4) {
                        a0 += buf[(r + 0) * PITCH + lane];
                        a1 += buf[(r + 1) * PITCH + lane];
                        a2 += buf[(r + 2) * PITCH + lane];
                        a3 += buf[(r + 3) * PITCH + lane];
                    }
                    orow[DIM + blk * 32 + lane] = (a0 + a1) + (a2 + a3);
                }
            }
        });
    });
}

// Forward for dim = 128, HALF a warp per sample: lane h of a half-warp holds the float4 column slices h and h + 16
// of the 27 rows (8 floats per row, 216 registers), so a partial dot product covers 8 columns (5 packed
// instructions) and only 16 partials per pair have to be added up -- against interact_fwd_tr_kernel (a warp per
// sample, 4 columns per lane, 32 partials per pair) that is 0.83x the multiply instructions and half of the
// shared-memory transpose and of the adds per sample: ~1 150 instead of ~2 050 warp instructions per sample.
// Blocks of 32 pairs: every lane stores its 32 partials as one row of a [32 lanes][36] tile (STS.128, conflict
// free), then lane L adds up columns 2(L%16), 2(L%16)+1 over the 16 rows of its own half-warp (LDS.64); the tile is
// double-buffered so that one __syncwarp per block is enough.  Two CTAs of four warps per SM (255 registers).
// The arithmetic of a half-warp on the rows it holds in registers (ta / tb: column slices h and h + 16 of the F rows of
// its sample).  PB transpose tiles per warp: 2 = one __syncwarp per block of 32 pairs, 1 = two (half the shared memory).
template <int F, int PB>
__device__ __forceinline__ void fwd_h_math(const float4 (&ta)[F], const float4 (&tb)[F], float* __restrict__ orow, bool live,
                                           float* __restrict__ s_part) {
    constexpr int DIM = 128, PITCH = 36;
    constexpr int NP = Pairs<F, false>::N;
    const int lane = threadIdx.x & 31;
    const int h = lane & 15, half = lane >> 4;
    if (live) {                                           // dense features pass through (model_no_ddp.py:293)
        orow[h * 4 + 0] = ta[0].x; orow[h * 4 + 1] = ta[0].y; orow[h * 4 + 2] = ta[0].z; orow[h * 4 + 3] = ta[0].w;
        orow[64 + h * 4 + 0] = tb[0].x; orow[64 + h * 4 + 1] = tb[0].y; orow[64 + h * 4 + 2] = tb[0].z; orow[64 + h * 4 + 3] = tb[0].w;
    }
    float v[8];
    static_for<F>([&](auto I) {
        constexpr int i = decltype(I)::value;
        constexpr int p0 = i * (i - 1) / 2;               // row-major offset into the strict lower triangle
        static_for<i>([&](auto J) {
            constexpr int j = decltype(J)::value;
            constexpr int p = p0 + j;
            {
                float2 q = __fmul2_rn(make_float2(ta[i].x, ta[i].y), make_float2(ta[j].x, ta[j].y));
                q = __ffma2_rn(make_float2(ta[i].z, ta[i].w), make_float2(ta[j].z, ta[j].w), q);
                q = __ffma2_rn(make_float2(tb[i].x, tb[i].y), make_float2(tb[j].x, tb[j].y), q);
                q = __ffma2_rn(make_float2(tb[i].z, tb[i].w), make_float2(tb[j].z, tb[j].w), q);
                v[p % 8] = q.x + q.y;
            }
            constexpr int blk = p / 32;
            if constexpr ((p + 1) % 8 == 0 || p + 1 == NP) {      // 8 partials (or the tail) -> two STS.128
                constexpr int c8 = (p % 32) / 8;                  // chunk of 8 within the block
                constexpr int n8 = p % 8 + 1;                     // valid partials in this chunk
                float4* wr = reinterpret_cast<float4*>(s_part + (PB == 2 ? (blk & 1) : 0) * 32 * PITCH + lane * PITCH + c8 * 8);
                wr[0] = make_float4(v[0], n8 > 1 ? v[1] : 0.f, n8 > 2 ? v[2] : 0.f, n8 > 3 ? v[3] : 0.f);
                if constexpr (n8 > 4) wr[1] = make_float4(v[4], n8 > 5 ? v[5] : 0.f, n8 > 6 ? v[6] : 0.f, n8 > 7 ? v[7] : 0.f);
            }
            if constexpr ((p + 1) % 32 == 0 || p + 1 == NP) {
                constexpr int cnt = p + 1 - blk * 32;
                const float* buf = s_part + (PB == 2 ? (blk & 1) : 0) * 32 * PITCH;
                __syncwarp();
                const int c0 = 2 * h;
                if (c0 < cnt) {
                    const float* col = buf + (half * 16) * PITCH + c0;
                    float2 a0 = make_float2(0.f, 0.f), a1 = make_float2(0.f, 0.f);
#pragma unroll
                    for (int r = 0; r < 16; r += 2) {
                        a0 = __fadd2_rn(a0, *reinterpret_cast<const float2*>(col + r * PITCH));
                        a1 = __fadd2_rn(a1, *reinterpret_cast<const float2*>(col + (r + 1) * PITCH));
                    }
                    a0 = __fadd2_rn(a0, a1);
                    if (live) {
                        orow[DIM + blk * 32 + c0] = a0.x;
                        if (c0 + 1 < cnt) orow[DIM + blk * 32 + c0 + 1] = a0.y;
                    }
                }
                if constexpr (PB == 1) __syncwarp();              // single tile: read before the next block's partials land
            }
        });
    });
}

template <int F>
__device__ __forceinline__ void fwd_h_pair(const FeatPtrs& fp, int64_t row_stride, int B, float* __restrict__ out,
                                           int64_t ld_out, int pair, float (*s_part)[32 * 36]) {
    const int lane = threadIdx.x & 31;
    const int h = lane & 15, half = lane >> 4;
    const int b = pair * 2 + half;
    const bool live = b < B;
    const int bb = live ? b : B - 1;                      // a dead half-warp (odd B) recomputes the last sample, stores nothing
    float4 ta[F], tb[F];
    static_for<F>([&](auto I) {
        constexpr int i = decltype(I)::value;
        const float4* row = reinterpret_cast<const float4*>(fp.p[i] + (int64_t)bb * row_stride);
        ta[i] = __ldg(row + h);
        tb[i] = __ldg(row + 16 + h);
    });
    fwd_h_math<F, 2>(ta, tb, out + (int64_t)bb * ld_out, live, &s_part[0][0]);
}

template <int F>
__global__ void __launch_bounds__(128, 2) interact_fwd_h_kernel(FeatPtrs fp, int64_t row_stride, int B,
                                                                float* __restrict__ out, int64_t ld_out) {
    pdl_enter();
    __shared__ __align__(16) float s_part[4][2][32 * 36];
    fwd_h_pair<F>(fp, row_stride, B, out, ld_out, (int)((blockIdx.x * blockDim.x + threadIdx.x) >> 5), s_part[threadIdx.x >> 5]);
}

// The same, persistent: a warp walks sample pairs gw, gw + W, ... (W = warps in the grid: two CTAs per SM).  A grid
// of one-shot CTAs runs in lock-step waves -- every warp of a wave loads its 27 KB at once (HBM-bound), then every warp
// computes (issue-bound) -- so neither resource is busy for more than half of the time.  Here the warps of an SM are
// started `stagger_ns` apart, so that some load while the others compute.
template <int F>
__global__ void __launch_bounds__(128, 2) interact_fwd_hp_kernel(FeatPtrs fp, int64_t row_stride, int B,
                                                                 float* __restrict__ out, int64_t ld_out, int n_pairs,
                                                                 unsigned stagger_ns) {
    pdl_enter();
    __shared__ __align__(16) float s_part[4][2][32 * 36];
    const int wib = threadIdx.x >> 5;
    const int gw = (int)((blockIdx.x * blockDim.x + threadIdx.x) >> 5), W = (int)((gridDim.x * blockDim.x) >> 5);
    // warps of the two CTAs that share an SM interleave: slot = (CTA parity, warp) in [0, 8)
    const unsigned slot = (blockIdx.x & 1u) * 4u + (unsigned)wib;
    if (stagger_ns && slot) __nanosleep(slot * stagger_ns);
    for (int pair = gw; pair < n_pairs; pair += W) {
        fwd_h_pair<F>(fp, row_stride, B, out, ld_out, pair, s_part[wib]);
        __syncwarp();                                     // the transpose tiles are reused by the next pair
    }
}

template <int F, int TPS, bool ITSELF>
__global__ void __launch_bounds__(128, 4) interact_bwd_kernel(FeatPtrs fp, int64_t row_stride, int B,
                                                           const float* __restrict__ d_out, int64_t ld_dout,
                                                           float* __restrict__ d_feat, int64_t ld_dfeat) {
    pdl_enter();
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
                g[i] = __ffma2_rn(make_float2(2.f * c, 2.f * c), t[i], g[i]);
            } else {                 // packed FMAs (FFMA2): one issue slot per float2
                const float2 cc = make_float2(c, c);
                g[i] = __ffma2_rn(cc, t[j], g[i]);
                g[j] = __ffma2_rn(cc, t[i], g[j]);
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

// ---- backward, dim 128, software-pipelined through shared memory -----------------------------------
// ncu on interact_bwd_kernel (B200, 27 x 128, B = 8192): 60 us, issue slots 33 % busy, DRAM 40 %, the top stall
// is long_scoreboard (4.9 stall cycles per issued instruction): with 128 registers only four 128-thread CTAs fit on an
// SM, every CTA loads its rows, waits a full HBM round trip, computes, stores -- and the four of them do that
// nearly in phase, so neither the memory system nor the issue slots stay busy.
// Here a CTA is persistent and owns a two-stage shared-memory ring: while the 128 threads work on the two samples
// of stage s, the 27 feature rows and the gradient rows of the next pair are already in flight into stage s^1 as
// bulk async copies (cp.async.bulk, completion on an mbarrier: no registers, no LSU slots held).  The arithmetic is
// the same as interact_bwd_kernel's (same order, same FFMA2s): results are bit-identical.
namespace pipe {
constexpr int SPB = 2;                       // samples per iteration (64 threads each)
__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void bar_init(uint64_t* bar, uint32_t n) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(s32(bar)), "r"(n) : "memory");
}
__device__ __forceinline__ void bar_expect(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bar_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok = 0, spins = 0;
    while (true) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(ok) : "r"(s32(bar)), "r"(parity) : "memory");
        if (ok) return;
        if (++spins > (1u << 26)) __trap();       // a protocol error becomes a launch failure, never a hung GPU
    }
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(s32(dst)), "l"(src), "r"(bytes), "r"(s32(bar)) : "memory");
}
}  // namespace pipe

template <int F>
struct BwdPipe {
    static constexpr int DIM = 128, TPS = 64;
    static constexpr int NP = Pairs<F, false>::N;
    static constexpr int ROWF = (DIM + NP + 3) & ~3;                   // floats copied per gradient row
    static constexpr int T_BYTES = F * pipe::SPB * DIM * 4;            // [F][SPB][DIM]
    static constexpr int C_BYTES = pipe::SPB * ROWF * 4;               // [SPB][ROWF]
    static constexpr int STAGE_BYTES = T_BYTES + C_BYTES;
    static constexpr int SMEM_BYTES = 2 * STAGE_BYTES + 16;
};

template <int F>
__global__ void __launch_bounds__(128, 3) interact_bwd_pipe_kernel(FeatPtrs fp, int64_t row_stride, int B,
                                                                   const float* __restrict__ d_out, int64_t ld_dout,
                                                                   float* __restrict__ d_feat, int64_t ld_dfeat) {
    using P = BwdPipe<F>;
    constexpr int DIM = P::DIM, TPS = P::TPS, SPB = pipe::SPB, NP = P::NP, ROWF = P::ROWF;
    extern __shared__ __align__(128) unsigned char pipe_smem[];
    uint64_t* bars = reinterpret_cast<uint64_t*>(pipe_smem + 2 * P::STAGE_BYTES);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int ls = tid / TPS, col = tid % TPS;
    const int items = (B + SPB - 1) / SPB;
    if (tid == 0) {
        pipe::bar_init(&bars[0], 1);
        pipe::bar_init(&bars[1], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    pdl_enter();

    // warp 0 issues the copies of one item: lane i < F moves feature i (both samples in one copy when the rows
    // are dense), lanes F .. F+SPB-1 the gradient rows
    auto issue = [&](int stage, int item) {
        unsigned char* base = pipe_smem + stage * P::STAGE_BYTES;
        const int b0 = item * SPB;
        const int ns = min(SPB, B - b0);
        if (lane == 0) pipe::bar_expect(&bars[stage], (uint32_t)(ns * (F * DIM * 4 + ROWF * 4)));
        __syncwarp();
        if (lane < F) {
            float* dst = reinterpret_cast<float*>(base) + lane * SPB * DIM;
            const float* src = fp.p[lane] + (int64_t)b0 * row_stride;
            if (row_stride == DIM) {
                pipe::bulk_g2s(dst, src, (uint32_t)(ns * DIM * 4), &bars[stage]);
            } else {
                for (int q = 0; q < ns; ++q) pipe::bulk_g2s(dst + q * DIM, src + (int64_t)q * row_stride, DIM * 4, &bars[stage]);
            }
        } else if (lane < F + ns) {
            const int q = lane - F;
            pipe::bulk_g2s(base + P::T_BYTES + q * ROWF * 4, d_out + (int64_t)(b0 + q) * ld_dout, ROWF * 4, &bars[stage]);
        }
    };
    if (warp == 0) {
        if ((int)blockIdx.x < items) issue(0, blockIdx.x);
        if ((int)(blockIdx.x + gridDim.x) < items) issue(1, blockIdx.x + gridDim.x);
    }

    int it = 0;
    for (int item = blockIdx.x; item < items; item += gridDim.x, ++it) {
        const int st = it & 1;
        pipe::bar_wait(&bars[st], (uint32_t)((it >> 1) & 1));
        const unsigned char* base = pipe_smem + st * P::STAGE_BYTES;
        const float* T = reinterpret_cast<const float*>(base);                         // [F][SPB][DIM]
        const float* crow = reinterpret_cast<const float*>(base + P::T_BYTES) + ls * ROWF;
        const int b = item * SPB + ls;
        const bool live = b < B;
        float2 t[F], g[F];
        static_for<F>([&](auto I) {
            constexpr int i = decltype(I)::value;
            t[i] = live ? reinterpret_cast<const float2*>(T + (i * SPB + ls) * DIM)[col] : make_float2(0.f, 0.f);
            g[i] = make_float2(0.f, 0.f);
        });
        if (live) g[0] = reinterpret_cast<const float2*>(crow)[col];     // d/dx of the pass-through
        const float4* c4p = reinterpret_cast<const float4*>(crow + DIM);
        float4 cq = make_float4(0.f, 0.f, 0.f, 0.f);
        static_for<F>([&](auto I) {
            constexpr int i = decltype(I)::value;
            constexpr int p0 = i * (i - 1) / 2;
            static_for<i>([&](auto J) {
                constexpr int j = decltype(J)::value;
                constexpr int p = p0 + j;
                if constexpr (p % 4 == 0) cq = c4p[p / 4];      // (a dead tail sample reads stale shared memory: never stored)
                const float c = (p % 4 == 0) ? cq.x : (p % 4 == 1) ? cq.y : (p % 4 == 2) ? cq.z : cq.w;
                const float2 cc = make_float2(c, c);
                g[i] = __ffma2_rn(cc, t[j], g[i]);
                g[j] = __ffma2_rn(cc, t[i], g[j]);
            });
        });
        if (live) {
            static_for<F>([&](auto I) {
                constexpr int i = decltype(I)::value;
                reinterpret_cast<float2*>(d_feat + (int64_t)i * ld_dfeat + (int64_t)b * DIM)[col] = g[i];
            });
        }
        __syncthreads();                                     // every thread is done with this stage
        const int nxt = item + 2 * (int)gridDim.x;
        if (warp == 0 && nxt < items) issue(st, nxt);
    }
}

// ---- forward, dim 128, software-pipelined through shared memory ------------------------------------
// Same diagnosis as the backward (ncu on interact_fwd_tr_kernel: 40 us, issue slots 44 % busy, DRAM 37 %, 16 resident
// warps per SM that all wait for their 27 rows at the same time).  Every warp is persistent and owns a private
// two-stage ring [2][F][128] floats + two mbarriers: lane i issues the bulk async copy of feature row i of the sample
// after next as soon as the current sample's rows sit in registers, so a full sample of work (about 2 000 warp
// instructions) hides each load.  Arithmetic and reduction order are interact_fwd_tr_kernel's: results are
// bit-identical.  Two warps per CTA, three CTAs per SM (2 x (27.6 KB ring + 9.2 KB transpose buffers) each).
// STAGES ring stages per warp (2: the sample after next is in flight; 1: the next sample, issued as soon as the
// current one sits in registers -- a sample of compute still hides the load), WARPS per CTA, PB transpose buffers
// per warp (2: one __syncwarp per block of 32 pairs, 1: two).  <2, 2, 2>: 36.9 KB per warp, 6 warps per SM;
// <1, 4, 1>: 18.4 KB per warp, 12 warps per SM.
template <int F, int STAGES, int WARPS_, int PB>
struct FwdPipe {
    static constexpr int DIM = 128, PITCH = 36, WARPS = WARPS_;
    static constexpr int RING_BYTES = STAGES * F * DIM * 4;            // per warp
    static constexpr int PART_BYTES = PB * 32 * PITCH * 4;             // per warp
    static constexpr int SMEM_BYTES = WARPS * (RING_BYTES + PART_BYTES) + WARPS * STAGES * 8;
    static constexpr int CTAS_PER_SM = 3;
};

template <int F, int STAGES, int WARPS, int PB>
__global__ void __launch_bounds__(32 * WARPS, 3) interact_fwd_pipe_kernel(FeatPtrs fp, int64_t row_stride, int B,
                                                                          float* __restrict__ out, int64_t ld_out) {
    using P = FwdPipe<F, STAGES, WARPS, PB>;
    constexpr int DIM = P::DIM, PITCH = P::PITCH;
    constexpr int NP = Pairs<F, false>::N;
    extern __shared__ __align__(128) unsigned char pipe_smem[];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    float* ring = reinterpret_cast<float*>(pipe_smem + wib * P::RING_BYTES);                                  // [STAGES][F][DIM]
    float* part = reinterpret_cast<float*>(pipe_smem + WARPS * P::RING_BYTES + wib * P::PART_BYTES);           // [PB][32 * PITCH]
    uint64_t* bars = reinterpret_cast<uint64_t*>(pipe_smem + WARPS * (P::RING_BYTES + P::PART_BYTES)) + wib * STAGES;
    if (lane == 0) {
#pragma unroll
        for (int q = 0; q < STAGES; ++q) pipe::bar_init(&bars[q], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();
    pdl_enter();
    const int gw = blockIdx.x * WARPS + wib, nw = gridDim.x * WARPS;
    auto issue = [&](int stage, int b) {
        if (lane == 0) pipe::bar_expect(&bars[stage], (uint32_t)(F * DIM * 4));
        __syncwarp();
        if (lane < F) pipe::bulk_g2s(ring + (stage * F + lane) * DIM, fp.p[lane] + (int64_t)b * row_stride, DIM * 4, &bars[stage]);
    };
#pragma unroll
    for (int q = 0; q < STAGES; ++q)
        if (gw + q * nw < B) issue(q, gw + q * nw);
    int it = 0;
    for (int b = gw; b < B; b += nw, ++it) {
        const int st = it % STAGES;
        pipe::bar_wait(&bars[st], (uint32_t)((it / STAGES) & 1));
        float4 t[F];
        static_for<F>([&](auto I) {
            constexpr int i = decltype(I)::value;
            t[i] = reinterpret_cast<const float4*>(ring + (st * F + i) * DIM)[lane];
        });
        __syncwarp();                                          // the stage is in registers: refill it right away
        if (b + STAGES * nw < B) issue(st, b + STAGES * nw);
        float* orow = out + (int64_t)b * ld_out;
        orow[lane * 4 + 0] = t[0].x; orow[lane * 4 + 1] = t[0].y;      // dense features pass through (model_no_ddp.py:293)
        orow[lane * 4 + 2] = t[0].z; orow[lane * 4 + 3] = t[0].w;
        float v[32];
        static_for<F>([&](auto I) {
            constexpr int i = decltype(I)::value;
            constexpr int p0 = i * (i - 1) / 2;                        // row-major triangle offset
            static_for<i>([&](auto J) {
                constexpr int j = decltype(J)::value;
                constexpr int p = p0 + j;
                v[p % 32] = dot4_packed(t[i], t[j]);
                if constexpr ((p + 1) % 32 == 0 || p + 1 == NP) {
                    constexpr int blk = p / 32;
                    constexpr int cnt = p + 1 - blk * 32;
                    float* buf = part + (PB == 2 ? (blk & 1) : 0) * 32 * PITCH;
                    float4* wr = reinterpret_cast<float4*>(buf + lane * PITCH);
                    static_for<(cnt + 3) / 4>([&](auto Q) {
                        constexpr int q = decltype(Q)::value;
                        wr[q] = make_float4(v[4 * q], 4 * q + 1 < cnt ? v[4 * q + 1] : 0.f, 4 * q + 2 < cnt ? v[4 * q + 2] : 0.f,
                                            4 * q + 3 < cnt ? v[4 * q + 3] : 0.f);
                    });
                    __syncwarp();
                    if (lane < cnt) {
                        float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
#pragma unroll
                        for (int r = 0; r < 32; r += 4) {
                            a0 += buf[(r + 0) * PITCH + lane];
                            a1 += buf[(r + 1) * PITCH + lane];
                            a2 += buf[(r + 2) * PITCH + lane];
                            a3 += buf[(r + 3) * PITCH + lane];
                        }
                        orow[DIM + blk * 32 + lane] = (a0 + a1) + (a2 + a3);
                    }
                    if constexpr (PB == 1) __syncwarp();       // single buffer: read before the next block's rows land
                }
            });
        });
        __syncwarp();     // the last blocks' transpose buffers are read before the next sample overwrites them
    }
}

// ---- forward, dim 128: half-warp arithmetic fed through shared memory ---------------------------------
// interact_fwd_h_kernel holds a sample pair in 216 registers per lane, so only 8 warps fit on an SM and all of them
// load (HBM-bound, issue slots idle), then compute (issue-bound, HBM idle), in lock-step waves: 30.6 us where the
// loads alone take 20 us and the arithmetic 9-16 us.  Here every warp is persistent and owns a one-item staging
// buffer [F][2 samples][128] in shared memory plus an mbarrier: as soon as the current pair of samples sits in
// registers (54 LDS.128 per lane), lanes 0 .. F-1 issue the bulk async copies of the warp's NEXT pair into the same
// buffer, and the 2 300 instructions of arithmetic on the registers hide that load.  One CTA per SM; WARPS warps of
// 27 KB staging + PB transpose tiles (4.6 KB each).  Arithmetic: fwd_h_math, bit-identical to interact_fwd_h_kernel.
template <int F, int WARPS_, int PB_>
struct FwdHs {
    static constexpr int DIM = 128, WARPS = WARPS_, PB = PB_;
    static constexpr int STAGE_BYTES = F * 2 * DIM * 4;                // per warp: [F][2][DIM]
    static constexpr int PART_BYTES = PB * 32 * 36 * 4;                // per warp
    static constexpr int SMEM_BYTES = WARPS * (STAGE_BYTES + PART_BYTES) + WARPS * 8;
};

template <int F, int WARPS, int PB>
__global__ void __launch_bounds__(32 * WARPS, 1) interact_fwd_hs_kernel(FeatPtrs fp, int64_t row_stride, int B,
                                                                        float* __restrict__ out, int64_t ld_out) {
    using P = FwdHs<F, WARPS, PB>;
    constexpr int DIM = P::DIM;
    extern __shared__ __align__(128) unsigned char pipe_smem[];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const int h = lane & 15, half = lane >> 4;
    float* stage = reinterpret_cast<float*>(pipe_smem + wib * P::STAGE_BYTES);                                  // [F][2][DIM]
    float* part = reinterpret_cast<float*>(pipe_smem + WARPS * P::STAGE_BYTES + wib * P::PART_BYTES);           // [PB][32 * 36]
    uint64_t* bar = reinterpret_cast<uint64_t*>(pipe_smem + WARPS * (P::STAGE_BYTES + P::PART_BYTES)) + wib;
    if (lane == 0) {
        pipe::bar_init(bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();
    pdl_enter();
    const int items = (B + 1) / 2;
    const int gw = blockIdx.x * WARPS + wib, nw = gridDim.x * WARPS;
    // lane i < F moves feature i of the item's samples (one copy when the rows of consecutive samples are adjacent)
    auto issue = [&](int item) {
        const int b0 = item * 2;
        const int ns = min(2, B - b0);
        if (lane == 0) pipe::bar_expect(bar, (uint32_t)(ns * F * DIM * 4));
        __syncwarp();
        if (lane < F) {
            float* dst = stage + lane * 2 * DIM;
            const float* src = fp.p[lane] + (int64_t)b0 * row_stride;
            if (row_stride == DIM) {
                pipe::bulk_g2s(dst, src, (uint32_t)(ns * DIM * 4), bar);
            } else {
                for (int q = 0; q < ns; ++q) pipe::bulk_g2s(dst + q * DIM, src + (int64_t)q * row_stride, DIM * 4, bar);
            }
        }
    };
    if (gw < items) issue(gw);
    int it = 0;
    for (int item = gw; item < items; item += nw, ++it) {
        pipe::bar_wait(bar, (uint32_t)(it & 1));
        const int b = item * 2 + half;
        const bool live = b < B;            // (the dead half-warp of an odd B works on stale shared memory, stores nothing)
        float4 ta[F], tb[F];
        static_for<F>([&](auto I) {
            constexpr int i = decltype(I)::value;
            const float4* row = reinterpret_cast<const float4*>(stage + (i * 2 + half) * DIM);
            ta[i] = row[h];
            tb[i] = row[16 + h];
        });
        __syncwarp();                                          // the item is in registers: refill the buffer right away
        if (item + nw < items) issue(item + nw);
        fwd_h_math<F, PB>(ta, tb, out + (int64_t)(live ? b : 0) * ld_out, live, part);
        __syncwarp();                                          // the transpose tiles are reused by the next item
    }
}

// ---- tensor-core path (F <= 32, dim % 32 == 0, no self-interaction) -----------------------------
// ncu on the CUDA-core kernels above: issue-bound (3048 / 3864 warp instructions per sample, FMA pipe
// 36 %, DRAM 35 %), because the K reduction costs one shuffle + add per pair on top of the FMAs.
// mma.sync.m16n8k8 does that reduction inside the instruction.  FP32 accuracy is kept with the 3xTF32
// split (hi = rna.tf32(x), lo = rna.tf32(x - hi); hi*hi + hi*lo + lo*hi, the cross terms in their own
// accumulator); one warp owns one sample and the fragments are loaded straight from global memory
// with 16-byte loads: the MMA's k index (forward) and n index (backward) are free permutations, so
// lane (g, t) reads whole float4s and no shared-memory transpose is needed.
__device__ __forceinline__ uint32_t tf32_rna(float x) {
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
    return r;
}
__device__ __forceinline__ void split_tf32(float x, uint32_t& hi, uint32_t& lo) {
    hi = tf32_rna(x);
    lo = tf32_rna(x - __uint_as_float(hi));
}
__device__ __forceinline__ void mma_tf32(float (&d)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
                 : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}

// Forward: Z = T T^T, T = [F, DIM].  Row groups of 8 rows: group r holds rows 8r + g.  Lane (g, t)
// loads, per 16-column chunk j, the float4 at column 16j + 4t of its row in every group; (x, y) feed
// one k-step and (z, w) the next (virtual k = t <-> x / z, k = t + 4 <-> y / w: the same permutation
// on the A and the B side, so the dot products are unchanged).  The A fragment of m-tile i is groups
// (2i, 2i + 1), the B fragment of n-tile n is group n: the same registers.
template <int F, int DIM>
__global__ void __launch_bounds__(128, 3) interact_fwd_mma_kernel(FeatPtrs fp, int64_t row_stride, int B,
                                                                  float* __restrict__ out, int64_t ld_out) {
    constexpr int MT = (F + 15) / 16, RG = 2 * MT;       // m-tiles, row groups
    constexpr int NCH = DIM / 16;                        // 16-column chunks
    // lower-triangle tiles (m-tile i, n-tile n <= 2i + 1)
    constexpr int NTILE = MT * (MT + 1);                 // sum over i of (2i + 2)
    const int lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
    const int warps = (gridDim.x * blockDim.x) >> 5;
    for (int b = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; b < B; b += warps) {
        const float* rowp[RG];
#pragma unroll
        for (int r = 0; r < RG; ++r) rowp[r] = (8 * r + g < F) ? fp.p[8 * r + g] + (int64_t)b * row_stride + 4 * t : nullptr;
        float big[NTILE][4], small[NTILE][4];
#pragma unroll
        for (int q = 0; q < NTILE; ++q)
#pragma unroll
            for (int e = 0; e < 4; ++e) big[q][e] = small[q][e] = 0.f;
#pragma unroll 2
        for (int j = 0; j < NCH; ++j) {
            float4 v[RG];
#pragma unroll
            for (int r = 0; r < RG; ++r)
                v[r] = rowp[r] ? __ldg(reinterpret_cast<const float4*>(rowp[r] + 16 * j)) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                uint32_t hi0[RG], lo0[RG], hi1[RG], lo1[RG];
#pragma unroll
                for (int r = 0; r < RG; ++r) {
                    split_tf32(h ? v[r].z : v[r].x, hi0[r], lo0[r]);
                    split_tf32(h ? v[r].w : v[r].y, hi1[r], lo1[r]);
                }
                int q = 0;
#pragma unroll
                for (int i = 0; i < MT; ++i) {
#pragma unroll
                    for (int n = 0; n <= 2 * i + 1; ++n, ++q) {
                        mma_tf32(small[q], lo0[2 * i], lo0[2 * i + 1], lo1[2 * i], lo1[2 * i + 1], hi0[n], hi1[n]);
                        mma_tf32(small[q], hi0[2 * i], hi0[2 * i + 1], hi1[2 * i], hi1[2 * i + 1], lo0[n], lo1[n]);
                        mma_tf32(big[q], hi0[2 * i], hi0[2 * i + 1], hi1[2 * i], hi1[2 * i + 1], hi0[n], hi1[n]);
                    }
                }
            }
        }
        float* orow = out + (int64_t)b * ld_out;
        // dense features pass through (model_no_ddp.py:293); rows of `out` are not 16-byte aligned
        const float* x = fp.p[0] + (int64_t)b * row_stride;
#pragma unroll
        for (int c = lane; c < DIM; c += 32) orow[c] = __ldg(x + c);
        // C fragment: c0 (g, 2t), c1 (g, 2t + 1), c2 (g + 8, 2t), c3 (g + 8, 2t + 1) of the tile
        int q = 0;
#pragma unroll
        for (int i = 0; i < MT; ++i) {
#pragma unroll
            for (int n = 0; n <= 2 * i + 1; ++n, ++q) {
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const int I = 16 * i + g + 8 * (e >> 1), J = 8 * n + 2 * t + (e & 1);
                    if (J < I && I < F) orow[DIM + I * (I - 1) / 2 + J] = big[q][e] + small[q][e];
                }
            }
        }
    }
}

// Backward: dT = S T with S the symmetric [F, F] matrix of the pair gradients (zero diagonal), plus the
// pass-through gradient on row 0.  A = S (m-tile i, k-step ks), B = T (k-step ks = rows 8ks + t, + 4;
// n-tile n = columns), D = dT.  The n index is permuted: virtual column c of n-tile n is the real column
// (DIM/8) * c + n, so lane (g, t) needs DIM/8 consecutive floats of its rows and owns DIM/8
// consecutive floats of its output rows.
template <int F, int DIM>
__global__ void __launch_bounds__(128, 3) interact_bwd_mma_kernel(FeatPtrs fp, int64_t row_stride, int B,
                                                                  const float* __restrict__ d_out, int64_t ld_dout,
                                                                  float* __restrict__ d_feat, int64_t ld_dfeat) {
    constexpr int MT = (F + 15) / 16, KS = (F + 7) / 8;  // m-tiles, k-steps
    constexpr int NTL = DIM / 8;                         // n-tiles = floats per lane per row
    constexpr int NG = NTL / 4;                          // groups of four n-tiles (one float4)
    const int lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
    const int warps = (gridDim.x * blockDim.x) >> 5;
    for (int b = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; b < B; b += warps) {
        const float* drow = d_out + (int64_t)b * ld_dout;
        // A fragments: a0 (I = 16i + g, J = 8ks + t), a1 (I + 8, J), a2 (I, J + 4), a3 (I + 8, J + 4)
        uint32_t ahi[MT][KS][4], alo[MT][KS][4];
#pragma unroll
        for (int i = 0; i < MT; ++i)
#pragma unroll
            for (int ks = 0; ks < KS; ++ks)
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const int I = 16 * i + g + 8 * (e & 1), J = 8 * ks + t + 4 * (e >> 1);
                    float c = 0.f;
                    if (I != J && I < F && J < F) {
                        const int hi = max(I, J), lo = min(I, J);
                        c = __ldg(drow + DIM + hi * (hi - 1) / 2 + lo);
                    }
                    split_tf32(c, ahi[i][ks][e], alo[i][ks][e]);
                }
        const float* rowp[2 * KS];
#pragma unroll
        for (int r = 0; r < 2 * KS; ++r) rowp[r] = (4 * r + t < F) ? fp.p[4 * r + t] + (int64_t)b * row_stride + NTL * g : nullptr;
#pragma unroll 1
        for (int c4 = 0; c4 < NG; ++c4) {
            float4 v[2 * KS];
#pragma unroll
            for (int r = 0; r < 2 * KS; ++r)
                v[r] = rowp[r] ? __ldg(reinterpret_cast<const float4*>(rowp[r]) + c4) : make_float4(0.f, 0.f, 0.f, 0.f);
            float acc[MT][4][4];
#pragma unroll
            for (int i = 0; i < MT; ++i)
#pragma unroll
                for (int n = 0; n < 4; ++n)
#pragma unroll
                    for (int e = 0; e < 4; ++e) acc[i][n][e] = 0.f;
#pragma unroll
            for (int ks = 0; ks < KS; ++ks) {
#pragma unroll
                for (int n = 0; n < 4; ++n) {
                    const float e0 = n == 0 ? v[2 * ks].x : n == 1 ? v[2 * ks].y : n == 2 ? v[2 * ks].z : v[2 * ks].w;
                    const float e1 = n == 0 ? v[2 * ks + 1].x : n == 1 ? v[2 * ks + 1].y : n == 2 ? v[2 * ks + 1].z : v[2 * ks + 1].w;
                    uint32_t bh0, bl0, bh1, bl1;
                    split_tf32(e0, bh0, bl0);
                    split_tf32(e1, bh1, bl1);
#pragma unroll
                    for (int i = 0; i < MT; ++i) {
                        mma_tf32(acc[i][n], alo[i][ks][0], alo[i][ks][1], alo[i][ks][2], alo[i][ks][3], bh0, bh1);
                        mma_tf32(acc[i][n], ahi[i][ks][0], ahi[i][ks][1], ahi[i][ks][2], ahi[i][ks][3], bl0, bl1);
                        mma_tf32(acc[i][n], ahi[i][ks][0], ahi[i][ks][1], ahi[i][ks][2], ahi[i][ks][3], bh0, bh1);
                    }
                }
            }
            // D fragment of n-tile n = 4 c4 + e: c0 (row g, real column NTL * 2t + n), c1 (NTL * (2t + 1) + n),
            // c2 / c3 the same for row g + 8: four consecutive floats per (row, column half)
#pragma unroll
            for (int i = 0; i < MT; ++i)
#pragma unroll
                for (int rh = 0; rh < 2; ++rh) {
                    const int I = 16 * i + g + 8 * rh;
                    if (I < F) {
#pragma unroll
                        for (int ch = 0; ch < 2; ++ch) {
                            const int col = NTL * (2 * t + ch) + 4 * c4;
                            float4 o = make_float4(acc[i][0][2 * rh + ch], acc[i][1][2 * rh + ch], acc[i][2][2 * rh + ch], acc[i][3][2 * rh + ch]);
                            if (I == 0) {       // d/dx of the pass-through
                                o.x += __ldg(drow + col); o.y += __ldg(drow + col + 1);
                                o.z += __ldg(drow + col + 2); o.w += __ldg(drow + col + 3);
                            }
                            *reinterpret_cast<float4*>(d_feat + (int64_t)I * ld_dfeat + (int64_t)b * DIM + col) = o;
                        }
                    }
                }
        }
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
    LAUNCH_PDL(K_INT_FWD, s, (interact_fwd_kernel<F, LPS, ITSELF>), blocks, 128, 0, fp, rs, B, out, ld_out);
}

template <int F, bool ITSELF>
void launch_fwd_tr(const FeatPtrs& fp, int64_t rs, int B, float* out, int64_t ld_out, cudaStream_t s) {
    const int blocks = (B + 3) / 4;                       // one warp per sample
    LAUNCH_PDL(K_INT_FWD, s, (interact_fwd_tr_kernel<F, ITSELF>), blocks, 128, 0, fp, rs, B, out, ld_out);
}

int g_fwd_stagger_ns = [] {
    const char* e = getenv("CDLRM_INTERACT_STAGGER_NS");
    return e ? atoi(e) : 800;
}();

template <int F>
void launch_fwd_h(const FeatPtrs& fp, int64_t rs, int B, float* out, int64_t ld_out, cudaStream_t s, bool persistent) {
    const int blocks = (B + 7) / 8;                       // half a warp per sample, 8 samples per block
    static int sms = 0;
    if (!sms) {
        int dev = 0;
        cudaGetDevice(&dev);
        if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms <= 0) sms = 148;
    }
    if (persistent && blocks > 2 * sms) {
        LAUNCH_PDL(K_INT_FWD, s, (interact_fwd_hp_kernel<F>), 2 * sms, 128, 0, fp, rs, B, out, ld_out, (B + 1) / 2,
                   (unsigned)g_fwd_stagger_ns);
    } else {
        LAUNCH_PDL(K_INT_FWD, s, (interact_fwd_h_kernel<F>), blocks, 128, 0, fp, rs, B, out, ld_out);
    }
}

template <int F, int TPS, bool ITSELF>
void launch_bwd(const FeatPtrs& fp, int64_t rs, int B, const float* d_out, int64_t ld_dout, float* d_feat,
                int64_t ld_dfeat, cudaStream_t s) {
    const int64_t threads = (int64_t)B * TPS;
    const int blocks = (int)((threads + 127) / 128);
    LAUNCH_PDL(K_INT_BWD, s, (interact_bwd_kernel<F, TPS, ITSELF>), blocks, 128, 0, fp, rs, B, d_out, ld_dout, d_feat, ld_dfeat);
}

int g_bwd_pipe = 1;      // cdlrm_interact_set_option(1, .): software-pipelined backward for dim 128
// cdlrm_interact_set_option(2, .): software-pipelined forward for dim 128 (1 = two-stage rings, 6 warps per SM;
// 2 = one-stage rings + single transpose buffer, 12 warps per SM).  Both bit-identical to interact_fwd_tr_kernel and
// neither faster on B200 (tools/interact_pipe_time.py, event-timed: 54 / 47-50 us against 46-48 us): hiding the row
// loads does not help a kernel whose 2 050 instructions per sample are mostly dependent FADD / LDS chains of the
// cross-lane reduction; the backward (independent FFMA2s, 2x the bytes per sample) gains 25 % from the same ring.
// 3 = interact_fwd_h_kernel (half a warp per sample, not bit-identical to the others: 8-column partials); 4 = the same,
// persistent with staggered warps; 5 / 6 = interact_fwd_hs_kernel (the half-warp arithmetic fed through shared memory by
// bulk async copies: 7 warps + single transpose tile / 6 warps + two tiles; bit-identical to 3); -1 = default
constexpr int FWD_DEFAULT = 3;
int g_fwd_pipe = [] {
    const char* e = getenv("CDLRM_INTERACT_FWD");       // A/B switch for whole-step measurements
    return (e && e[0] >= '0' && e[0] <= '6') ? e[0] - '0' : FWD_DEFAULT;
}();
int g_variant = 0;       // cdlrm_interact_set_option(0, .): 0 CUDA cores (default), 1 mma.sync 3xTF32, 2 first CUDA-core version
bool use_simt_only() { return g_variant != 1; }

int interact_grid(int B) {
    static int sms = 0;
    if (!sms) {
        int dev = 0;
        cudaGetDevice(&dev);
        if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms <= 0) sms = 148;
    }
    const int need = (B + 3) / 4;                 // one warp per sample, 4 warps per block
    const int cap = sms * 3 * 4;                  // 3 resident blocks per SM, a few samples per warp
    return need < cap ? need : cap;
}

template <int F, int STAGES, int WARPS, int PB>
bool launch_fwd_pipe(const FeatPtrs& fp, int64_t rs, int B, float* out, int64_t ld_out, cudaStream_t s) {
    using P = FwdPipe<F, STAGES, WARPS, PB>;
    static int sms = 0;
    if (!sms) {
        int dev = 0;
        cudaGetDevice(&dev);
        if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms <= 0) sms = 148;
    }
    if (cdlrm_smem_optin((const void*)interact_fwd_pipe_kernel<F, STAGES, WARPS, PB>, P::SMEM_BYTES) != cudaSuccess) {
        cudaGetLastError();
        return false;
    }
    const int need = (B + WARPS - 1) / WARPS;
    const int grid = need < sms * P::CTAS_PER_SM ? need : sms * P::CTAS_PER_SM;      // persistent warps
    LAUNCH_PDL(K_INT_FWD, s, (interact_fwd_pipe_kernel<F, STAGES, WARPS, PB>), grid, 32 * WARPS, P::SMEM_BYTES, fp, rs, B, out, ld_out);
    return true;
}

template <int F, int WARPS, int PB>
bool launch_fwd_hs(const FeatPtrs& fp, int64_t rs, int B, float* out, int64_t ld_out, cudaStream_t s) {
    using P = FwdHs<F, WARPS, PB>;
    static int sms = 0;
    if (!sms) {
        int dev = 0;
        cudaGetDevice(&dev);
        if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms <= 0) sms = 148;
    }
    if (cdlrm_smem_optin((const void*)interact_fwd_hs_kernel<F, WARPS, PB>, P::SMEM_BYTES) != cudaSuccess) {
        cudaGetLastError();
        return false;
    }
    const int items = (B + 1) / 2;
    const int need = (items + WARPS - 1) / WARPS;
    const int grid = need < sms ? need : sms;                 // persistent: one CTA per SM
    LAUNCH_PDL(K_INT_FWD, s, (interact_fwd_hs_kernel<F, WARPS, PB>), grid, 32 * WARPS, P::SMEM_BYTES, fp, rs, B, out, ld_out);
    return true;
}

template <int F>
bool launch_bwd_pipe(const FeatPtrs& fp, int64_t rs, int B, const float* d_out, int64_t ld_dout, float* d_feat,
                     int64_t ld_dfeat, cudaStream_t s) {
    using P = BwdPipe<F>;
    static int sms = 0;
    if (!sms) {
        int dev = 0;
        cudaGetDevice(&dev);
        if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms <= 0) sms = 148;
    }
    if (cdlrm_smem_optin((const void*)interact_bwd_pipe_kernel<F>, P::SMEM_BYTES) != cudaSuccess) {
        cudaGetLastError();
        return false;
    }
    const int items = (B + pipe::SPB - 1) / pipe::SPB;
    const int grid = items < sms * 3 ? items : sms * 3;          // persistent: three CTAs per SM (63 KB of smem each)
    LAUNCH_PDL(K_INT_BWD, s, (interact_bwd_pipe_kernel<F>), grid, 128, P::SMEM_BYTES, fp, rs, B, d_out, ld_dout, d_feat, ld_dfeat);
    return true;
}

template <int F, int DIM>
void launch_fwd_mma(const FeatPtrs& fp, int64_t rs, int B, float* out, int64_t ld_out, cudaStream_t s) {
    LAUNCH(K_INT_FWD, s, (interact_fwd_mma_kernel<F, DIM><<<interact_grid(B), 128, 0, s>>>(fp, rs, B, out, ld_out)));
}

template <int F, int DIM>
void launch_bwd_mma(const FeatPtrs& fp, int64_t rs, int B, const float* d_out, int64_t ld_dout, float* d_feat,
                    int64_t ld_dfeat, cudaStream_t s) {
    LAUNCH(K_INT_BWD, s, (interact_bwd_mma_kernel<F, DIM><<<interact_grid(B), 128, 0, s>>>(fp, rs, B, d_out, ld_dout, d_feat, ld_dfeat)));
}

bool aligned_for(const FeatPtrs& fp, int n, int64_t rs, int bytes) {
    for (int i = 0; i < n; ++i)
        if ((uintptr_t)fp.p[i] % bytes) return false;
    return (rs * 4) % bytes == 0;
}

}  // namespace

#define FWD_MMA_CASE(F_, D_)                                                              \
    if (!done && n_feat == F_ && dim == D_ && !itself && fast16 && !use_simt_only()) {    \
        launch_fwd_mma<F_, D_>(fp, rs, batch, out, ld_out, s);                            \
        done = true;                                                                      \
    }
#define BWD_MMA_CASE(F_, D_)                                                              \
    if (!done && n_feat == F_ && dim == D_ && !itself && fast16 && !use_simt_only()) {    \
        launch_bwd_mma<F_, D_>(fp, rs, batch, d_out, ld_dout, d_feat, ld_dfeat, s);        \
        done = true;                                                                      \
    }
#define FWD_TR_CASE(F_)                                                                  \
    if (!done && n_feat == F_ && dim == 128 && !itself && fast16 && g_variant == 0) {     \
        launch_fwd_tr<F_, false>(fp, rs, batch, out, ld_out, s);                          \
        done = true;                                                                      \
    }
#define FWD_PIPE_CASE(F_)                                                                                   \
    if (!done && n_feat == F_ && dim == 128 && !itself && fast16 && g_fwd_pipe >= 5 && g_variant == 0) {     \
        done = g_fwd_pipe == 5 ? launch_fwd_hs<F_, 7, 1>(fp, rs, batch, out, ld_out, s)                      \
                               : launch_fwd_hs<F_, 6, 2>(fp, rs, batch, out, ld_out, s);                     \
    }                                                                                                        \
    if (!done && n_feat == F_ && dim == 128 && !itself && fast16 && g_fwd_pipe >= 3 && g_variant == 0) {     \
        launch_fwd_h<F_>(fp, rs, batch, out, ld_out, s, g_fwd_pipe == 4);                                    \
        done = true;                                                                                         \
    }                                                                                                        \
    if (!done && n_feat == F_ && dim == 128 && !itself && fast16 && g_fwd_pipe && g_variant == 0) {          \
        done = g_fwd_pipe == 2 ? launch_fwd_pipe<F_, 1, 4, 1>(fp, rs, batch, out, ld_out, s)                 \
                               : launch_fwd_pipe<F_, 2, 2, 2>(fp, rs, batch, out, ld_out, s);                \
    }
#define FWD_CASE(F_, D_)                                                                  \
    if (!done && n_feat == F_ && dim == D_ && !itself && fast16) {                        \
        launch_fwd<F_, D_ / 4, false>(fp, rs, batch, out, ld_out, s);                      \
        done = true;                                                                      \
    }
// pipelined backward: dim 128, 16-byte aligned rows everywhere, gradient rows padded to a multiple of 4 floats
#define BWD_PIPE_CASE(F_)                                                                                   \
    if (!done && n_feat == F_ && dim == 128 && !itself && fast16 && g_bwd_pipe && g_variant == 0 &&         \
        ((uintptr_t)d_out % 16 == 0) && (ld_dout % 4 == 0) && ld_dout >= BwdPipe<F_>::ROWF) {                 \
        done = launch_bwd_pipe<F_>(fp, rs, batch, d_out, ld_dout, d_feat, ld_dfeat, s);                      \
    }
#define BWD_CASE(F_, D_)                                                                  \
    if (!done && n_feat == F_ && dim == D_ && !itself && fast8) {                         \
        launch_bwd<F_, D_ / 2, false>(fp, rs, batch, d_out, ld_dout, d_feat, ld_dfeat, s); \
        done = true;                                                                      \
    }

// key 0: kernel variant (0 = CUDA cores with the shared-memory transpose reduction, the default and the fastest
// measured; 1 = mma.sync.m16n8k8 3xTF32, kept as the evidence for "tensor cores do not pay off here":
// 78 / 93 us against 51 / 73 us forward / backward at B = 8192, 27 x 128 on B200; 2 = the butterfly version)
// key 1: software-pipelined backward for dim 128 (interact_bwd_pipe_kernel; default 1), 0 = interact_bwd_kernel
// key 2: software-pipelined forward (interact_fwd_pipe_kernel): 0 = interact_fwd_tr_kernel (default), 1 = two-stage
// rings (6 warps per SM: measured slower), 2 = one-stage rings + single transpose buffer (12 warps per SM)
extern "C" int cdlrm_interact_set_option(int key, int value) {
    if (key == 1) {
        g_bwd_pipe = value ? 1 : 0;
        return CDLRM_OK;
    }
    if (key == 2) {
        ARG_CHECK(value >= -1 && value <= 6);
        g_fwd_pipe = value < 0 ? FWD_DEFAULT : value;
        return CDLRM_OK;
    }
    if (key == 3) {                 // start stagger (ns) of the persistent forward's warps (variant 4)
        ARG_CHECK(value >= 0 && value <= 100000);
        g_fwd_stagger_ns = value;
        return CDLRM_OK;
    }
    ARG_CHECK(key == 0 && value >= 0 && value <= 2);
    g_variant = value;
    return CDLRM_OK;
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
    FWD_MMA_CASE(27, 128) FWD_MMA_CASE(27, 64) FWD_MMA_CASE(27, 32) FWD_MMA_CASE(9, 128) FWD_MMA_CASE(9, 64) FWD_MMA_CASE(9, 32)
    FWD_PIPE_CASE(27) FWD_PIPE_CASE(9)
    FWD_TR_CASE(27) FWD_TR_CASE(9)
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
    const bool fast16 = aligned_for(fp, n_feat, rs, 16) && ((uintptr_t)d_feat % 16 == 0) && (ld_dfeat % 4 == 0);
    bool done = false;
    BWD_MMA_CASE(27, 128) BWD_MMA_CASE(27, 64) BWD_MMA_CASE(27, 32) BWD_MMA_CASE(9, 128) BWD_MMA_CASE(9, 64) BWD_MMA_CASE(9, 32)
    BWD_PIPE_CASE(27) BWD_PIPE_CASE(9)
    BWD_CASE(27, 128) BWD_CASE(27, 64) BWD_CASE(27, 32) BWD_CASE(27, 16)
    BWD_CASE(9, 128) BWD_CASE(9, 64) BWD_CASE(9, 32) BWD_CASE(9, 16)
    if (!done)
        LAUNCH(K_INT_BWD, s, interact_bwd_generic<<<batch, 128, 0, s>>>(fp, n_feat, rs, batch, dim, itself, d_out, ld_dout, d_feat, ld_dfeat));
    CU_CHECK(cudaGetLastError());
    return CDLRM_OK;
}
