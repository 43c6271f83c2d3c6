// expdraw.cuh -- device restatement of one draw of torch.empty(n).exponential_(1) on the
// CPU generator, as consumed by torch.distributions.Categorical(...).sample()
// (main_no_ddp.py:183-185 of the reference):
//   r64 = (mt32() << 32) | mt32();  u = (r64 & (2^53-1)) * 2^-53;  q = float32(-log1p(-u))
// log1p is glibc 2.39's x86_64 __log1p_fma (sysdeps/ieee754/dbl-64/s_log1p.c built with FMA
// contraction), restated operation by operation with explicit IEEE intrinsics so that the
// double result -- and therefore the float32 draw -- is bit-identical to the host's
// (checked on the CPU against libm over 3e8 arguments incl. u -> 0 and u -> 1, and on the
// GPU against cdlrm_rng_exponential by tests/test_gpu_parity.py).
#pragma once
#include <stdint.h>

__device__ __forceinline__ double glibc_log1p_neg(double x) {  // valid for -1 < x <= 0
    const double ln2_hi = 6.93147180369123816490e-01, ln2_lo = 1.90821492927058770002e-10;
    const double Lp1 = 6.666666666666735130e-01, Lp2 = 3.999999999940941908e-01, Lp3 = 2.857142874366239149e-01,
                 Lp4 = 2.222219843214978396e-01, Lp5 = 1.818357216161805012e-01, Lp6 = 1.531383769920937332e-01,
                 Lp7 = 1.479819860511658591e-01;
    const int hx = __double2hiint(x);
    const int ax = hx & 0x7fffffff;
    double f = 0.0, c = 0.0, u;
    int k = 1, hu = 0;
    if (ax >= 0x3ff00000) return -__longlong_as_double(0x7ff0000000000000ll);  // x == -1 (cannot happen: u < 1)
    if (ax < 0x3e200000) {                                                      // |x| < 2^-29
        if (ax < 0x3c900000) return x;                                          // |x| < 2^-54
        return __fma_rn(-__dmul_rn(x, x), 0.5, x);
    }
    if (hx > 0 || hx <= (int)0xbfd2bec3) {  // -0.2929 < x
        k = 0;
        f = x;
        hu = 1;
    }
    if (k != 0) {
        u = __dadd_rn(1.0, x);
        hu = __double2hiint(u);
        k = (hu >> 20) - 1023;
        c = (k > 0) ? __dsub_rn(1.0, __dsub_rn(u, x)) : __dsub_rn(x, __dsub_rn(u, 1.0));
        c = __ddiv_rn(c, u);
        hu &= 0x000fffff;
        if (hu < 0x6a09e) {
            u = __hiloint2double(hu | 0x3ff00000, __double2loint(u));
        } else {
            k += 1;
            u = __hiloint2double(hu | 0x3fe00000, __double2loint(u));
            hu = (0x00100000 - hu) >> 2;
        }
        f = __dsub_rn(u, 1.0);
    }
    const double hfsq = __dmul_rn(__dmul_rn(0.5, f), f);
    const double kd = (double)k;
    if (hu == 0) {  // |f| < 2^-20
        if (f == 0.0) {
            if (k == 0) return 0.0;
            c = __fma_rn(kd, ln2_lo, c);
            return __fma_rn(kd, ln2_hi, c);
        }
        const double R = __dmul_rn(hfsq, __fma_rn(-0.66666666666666666, f, 1.0));
        if (k == 0) return __dsub_rn(f, R);
        return __fma_rn(kd, ln2_hi, -__dsub_rn(__dsub_rn(R, __fma_rn(kd, ln2_lo, c)), f));
    }
    const double s = __ddiv_rn(f, __dadd_rn(2.0, f));
    const double z = __dmul_rn(s, s);
    const double R2 = __fma_rn(z, Lp3, Lp2), R3 = __fma_rn(z, Lp5, Lp4), R4 = __fma_rn(z, Lp7, Lp6);
    const double z2 = __dmul_rn(z, z), z4 = __dmul_rn(z2, z2), z6 = __dmul_rn(z4, z2);
    const double R = __fma_rn(z6, R4, __fma_rn(z4, R3, __fma_rn(z, Lp1, __dmul_rn(z2, R2))));
    const double t3 = __dmul_rn(s, __dadd_rn(hfsq, R));
    if (k == 0) return __dsub_rn(f, __dsub_rn(hfsq, t3));
    return __fma_rn(kd, ln2_hi, -__dsub_rn(__dsub_rn(hfsq, __dadd_rn(t3, __fma_rn(kd, ln2_lo, c))), f));
}

// raw: two consecutive mt19937 outputs {hi, lo} of one draw
__device__ __forceinline__ float exp_draw_from_raw(uint2 raw) {
    const unsigned long long r = ((unsigned long long)raw.x << 32) | raw.y;
    const double u = __dmul_rn((double)(long long)(r & ((1ull << 53) - 1ull)), 1.0 / 9007199254740992.0);
    return __double2float_rn(-glibc_log1p_neg(-u));
}
