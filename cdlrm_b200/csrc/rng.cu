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
