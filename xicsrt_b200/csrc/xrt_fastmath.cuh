// xrt_fastmath.cuh -- the few FP64 transcendentals the ray code needs, written for
// the argument ranges that occur there.
//
// Why not the CUDA math library: its sincos/log/exp/asin/acos materialise every
// polynomial coefficient with two UMOV instructions (12.7 % of all issued
// instructions in the first version of the fused kernel, profiles/r01_*), carry
// range reduction and special-case paths for arguments that cannot occur here, and
// sincos(2*pi*u) pays a Payne-Hanek stack frame.  Here the coefficients sit in
// __constant__ tables (one LDCU.128 fetches two of them into uniform registers),
// and the reductions are exact for the ranges in use.
//
// Accuracy (checked on the host against long double, tests/test_fastmath.py):
// a few 1e-16 relative, far inside the 1e-9 parity tolerance.
//
// Every function is __host__ __device__ so the same source is exercised on the CPU.
#pragma once
#include <stdint.h>
#include <string.h>
#include <math.h>

#ifdef __CUDACC__
#define XRT_HD __host__ __device__ __forceinline__
#define XRT_ALIGN16 __align__(16)     // a table starts on a 16-byte boundary: LDCU.128 fetches coefficient pairs
#else
#define XRT_HD inline
#define XRT_ALIGN16
#define __constant__
#endif

#ifdef __CUDA_ARCH__
#define XRT_TAB(name) name##_d
#else
#define XRT_TAB(name) name##_h
#endif
#define XRT_DEFINE_TABLE(name, n, ...)                       \
    static __constant__ XRT_ALIGN16 double name##_d[n] = {__VA_ARGS__};  \
    static const double name##_h[n] = {__VA_ARGS__};

namespace xrt {

// sin(x) = x + x^3 (S[0] + z S[1] + ... + z^5 S[5]),  |x| <= pi/4   (fdlibm __kernel_sin)
XRT_DEFINE_TABLE(kSin, 6,
    -1.66666666666666324348e-01, 8.33333333332248946124e-03, -1.98412698298579493134e-04,
    2.75573137070700676789e-06, -2.50507602534068634195e-08, 1.58969099521155010221e-10)
// cos(x) = 1 - z/2 + z^2 (C[0] + z C[1] + ... + z^5 C[5]),  |x| <= pi/4   (fdlibm __kernel_cos)
XRT_DEFINE_TABLE(kCos, 6,
    4.16666666666666019037e-02, -1.38888888888741095749e-03, 2.48015872894767294178e-05,
    -2.75573143513906633035e-07, 2.08757232129817482790e-09, -1.13596475577881948265e-11)
// log(1+f) = f - f^2/2 + s (f^2/2 + R(z)), s = f/(2+f), z = s^2, R = z Lg[0] + ... + z^7 Lg[6]  (fdlibm __ieee754_log)
XRT_DEFINE_TABLE(kLog, 7,
    6.666666666666735130e-01, 3.999999999940941908e-01, 2.857142874366239149e-01, 2.222219843214978396e-01,
    1.818357216161805012e-01, 1.531383769920937332e-01, 1.479819860511658591e-01)
// exp(r) = sum r^k / k!, k = 0..13, |r| <= ln2/2  (truncation 4e-18)
XRT_DEFINE_TABLE(kExp, 12,
    1.0 / 6227020800.0, 1.0 / 479001600.0, 1.0 / 39916800.0, 1.0 / 3628800.0, 1.0 / 362880.0, 1.0 / 40320.0,
    1.0 / 5040.0, 1.0 / 720.0, 1.0 / 120.0, 1.0 / 24.0, 1.0 / 6.0, 0.5)
// Inverse normal CDF through erfinv(x) / x as a polynomial in w = -log(1 - x^2) (centre) or sqrt(w) (tails), the form
// of M. Giles' single-precision erfinv; coefficients: Chebyshev interpolation in 60-digit arithmetic (mpmath) converted
// to powers of the centred variable, truncation below 4e-18 relative.
// erfinv(x) / x = sum c_i (w - 3.125)^i,  w = -log(1 - x^2) in [0, 6.25]   (|x| <= 0.99903: 99.9 % of the uniforms)
XRT_DEFINE_TABLE(kNinvC, 25,
    1.6536545626831027, 0.24015818242558834, -0.006033670871426851, -0.0007407025341546431,
    0.00018673420801981186, -1.3882523393957483e-05, -1.3654691758785656e-06, 4.23478816822246e-07,
    -2.907039127564132e-08, -4.1126604371632185e-09, 1.051223377050429e-09, -5.414303283919504e-11,
    -1.2978805369932565e-11, 2.6305268312595183e-12, -8.07192593899004e-14, -4.0020031087558496e-14,
    6.521333511502239e-15, -3.94018812230432e-17, -1.2215637192404172e-16, 1.5510787009902526e-17,
    6.075050702072414e-19, -3.4734793888538036e-19, 1.999259988861535e-20, 3.194015548136271e-21,
    -3.5932028927020693e-22)
// erfinv(x) / x = sum c_i (sqrt(w) - 3.25)^i,  w in [6.25, 16]
XRT_DEFINE_TABLE(kNinvT1, 21,
    3.0838856104922208, 1.0052589676941655, 0.005370914553555033, -0.0037512085082247342,
    0.00249144209795696, -0.0016882755354488555, 0.0009532893415794137, -0.0003550378137852452,
    2.4031512865758357e-05, 6.828711739251955e-05, -4.732068066544697e-05, 1.2465028224217455e-05,
    2.93257845538535e-06, -3.985705945648173e-06, 1.4815977874022539e-06, -2.761716223816241e-08,
    -2.4549818963339823e-07, 1.3158663683192966e-07, -2.091453334773126e-08, -1.5305397964152548e-08,
    7.680200479077053e-09)
// erfinv(x) / x = sum c_i (sqrt(w) - 5.01)^i,  w in [16, 36.3]   (u down to 2^-54)
XRT_DEFINE_TABLE(kNinvT2, 20,
    4.860009391971025, 1.010297626272141, -0.00014512482094814632, -0.00021200990027194828,
    7.501794401970369e-05, -1.941228065183597e-05, 4.457176465510443e-06, -9.749411285981377e-07,
    2.2309945268031453e-07, -6.475270571992455e-08, 2.739547529297803e-08, -1.4302618643547702e-08,
    7.374469177986523e-09, -3.3530766811338047e-09, 1.258897391154837e-09, -3.702974831012114e-10,
    7.228405264876682e-11, 1.1638608468654179e-11, -1.940432963946844e-11, 5.767787767112359e-12)
// constants
XRT_DEFINE_TABLE(kMisc, 6,
    1.57079632679489661923,        // pi/2
    6.93147180369123816490e-01,    // ln2_hi
    1.90821492927058770002e-10,    // ln2_lo
    1.44269504088896338700,        // log2(e)
    6.28318530717958647692,        // 2 pi
    0.0)

XRT_HD int32_t hi_word(double x) {
#ifdef __CUDA_ARCH__
    return __double2hiint(x);
#else
    int64_t b; memcpy(&b, &x, 8); return (int32_t)(b >> 32);
#endif
}
XRT_HD int32_t lo_word(double x) {
#ifdef __CUDA_ARCH__
    return __double2loint(x);
#else
    int64_t b; memcpy(&b, &x, 8); return (int32_t)(b & 0xffffffff);
#endif
}
XRT_HD double from_words(int32_t hi, int32_t lo) {
#ifdef __CUDA_ARCH__
    return __hiloint2double(hi, lo);
#else
    int64_t b = ((int64_t)hi << 32) | (uint32_t)lo; double x; memcpy(&x, &b, 8); return x;
#endif
}
XRT_HD double fm(double a, double b, double c) {
#ifdef __CUDA_ARCH__
    return fma(a, b, c);
#else
    return __builtin_fma(a, b, c);
#endif
}
XRT_HD double round_even(double x) {
#ifdef __CUDA_ARCH__
    return rint(x);
#else
    return __builtin_rint(x);
#endif
}

// 1 / x for normal-range x: hardware seed plus two Newton steps on the device (no special-case
// path), a plain division on the host.
XRT_HD double recip(double x) {
#ifdef __CUDA_ARCH__
    double y;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    double e = fma(-x, y, 1.0);
    y = fma(y, fma(e, e, e), y);
    e = fma(-x, y, 1.0);
    return fma(y, e, y);
#else
    return 1.0 / x;
#endif
}

// sin(2 pi u), cos(2 pi u) for u in [0, 1].  Reduction: t = 4u, q = rint(t), r = t - q is
// exact, x = r pi/2 in [-pi/4, pi/4]; quadrant from q mod 4.
XRT_HD void sincos_2pi(double u, double &s, double &c) {
    const double t = 4.0 * u;
    const double q = round_even(t);
    const double x = (t - q) * XRT_TAB(kMisc)[0];
    const double z = x * x;
    double ps = XRT_TAB(kSin)[5];
    ps = fm(ps, z, XRT_TAB(kSin)[4]); ps = fm(ps, z, XRT_TAB(kSin)[3]); ps = fm(ps, z, XRT_TAB(kSin)[2]);
    ps = fm(ps, z, XRT_TAB(kSin)[1]); ps = fm(ps, z, XRT_TAB(kSin)[0]);
    const double sk = fm(x * z, ps, x);
    double pc = XRT_TAB(kCos)[5];
    pc = fm(pc, z, XRT_TAB(kCos)[4]); pc = fm(pc, z, XRT_TAB(kCos)[3]); pc = fm(pc, z, XRT_TAB(kCos)[2]);
    pc = fm(pc, z, XRT_TAB(kCos)[1]); pc = fm(pc, z, XRT_TAB(kCos)[0]);
    const double ck = fm(z * z, pc, fm(z, -0.5, 1.0));
    const int iq = (int)q;
    const double a = (iq & 1) ? ck : sk;      // |sin|
    const double b = (iq & 1) ? sk : ck;      // |cos|
    s = (iq & 2) ? -a : a;
    c = ((iq + 1) & 2) ? -b : b;
}

// Table version for the fused kernel: (cos, sin)(2 pi k / 256) from a 4 kB shared-memory table
// (filled by the block with sincos_2pi, exact arguments) and a rotation by the remainder
// |r| <= pi / 256, for which sin needs terms to r^5 and cos to r^6 (next terms 8e-18, 1e-20).
// 16 FP64 instructions, one 16-byte table read and no quadrant selects, against 28 and a dozen
// selects; error 3e-16 (table) + 1e-16.  tab[2k] = cos, tab[2k + 1] = sin; k = 256 wraps to 0.
constexpr int kSincosTable = 256;
XRT_HD void sincos_2pi_tab(double u, const double *tab, double &s, double &c) {
    const double q = u * (double)kSincosTable;
    const double magic = 6755399441055744.0;             // 1.5 * 2^52: q + magic holds rint(q) in its low word
    const double tq = q + magic;
    const int k = lo_word(tq) & (kSincosTable - 1);
    const double r = (q - (tq - magic)) * (6.283185307179586476925286766559 / kSincosTable);
    const double ck = tab[2 * k], sk = tab[2 * k + 1];
    const double r2 = r * r;
    const double sr = fm(r * r2, fm(r2, 1.0 / 120.0, -1.0 / 6.0), r);
    const double cr = fm(r2, fm(r2, fm(r2, -1.0 / 720.0, 1.0 / 24.0), -0.5), 1.0);
    c = fm(-sk, sr, ck * cr);
    s = fm(ck, sr, sk * cr);
}

// cos(2 pi u) alone: the quadrant parity decides which of the two kernels is needed, and both
// have the same Horner shape, so one chain runs on coefficients picked per lane -- six fused
// multiply-adds instead of twelve.  Same reduction and same values as sincos_2pi.
XRT_HD double cos_2pi(double u) {
    const double t = 4.0 * u;
    const double q = round_even(t);
    const double x = (t - q) * XRT_TAB(kMisc)[0];
    const double z = x * x;
    const int iq = (int)q;
    const bool odd = (iq & 1) != 0;                 // odd quadrant: |cos(2 pi u)| = |sin x|
    double p = odd ? XRT_TAB(kSin)[5] : XRT_TAB(kCos)[5];
    p = fm(p, z, odd ? XRT_TAB(kSin)[4] : XRT_TAB(kCos)[4]);
    p = fm(p, z, odd ? XRT_TAB(kSin)[3] : XRT_TAB(kCos)[3]);
    p = fm(p, z, odd ? XRT_TAB(kSin)[2] : XRT_TAB(kCos)[2]);
    p = fm(p, z, odd ? XRT_TAB(kSin)[1] : XRT_TAB(kCos)[1]);
    p = fm(p, z, odd ? XRT_TAB(kSin)[0] : XRT_TAB(kCos)[0]);
    const double sk = fm(x * z, p, x);                       // sin x   (odd)
    const double ck = fm(z * z, p, fm(z, -0.5, 1.0));        // cos x   (even)
    const double b = odd ? sk : ck;
    return ((iq + 1) & 2) ? -b : b;
}

// natural logarithm for normal, finite v > 0 (here v in [2^-53, 1])
XRT_HD double log_pos(double v) {
    int32_t hx = hi_word(v);
    int32_t k = (hx >> 20) - 1023;
    hx &= 0x000fffff;
    const int32_t i = (hx + 0x95f64) & 0x100000;   // mantissa >= sqrt(2): use m/2, k+1
    const double m = from_words(hx | (i ^ 0x3ff00000), lo_word(v));
    k += (i >> 20);
    const double f = m - 1.0;
    const double s = f * recip(2.0 + f);
    const double dk = (double)k;
    const double z = s * s;
    // two interleaved Horner chains in w = z^2 (odd and even coefficients, as fdlibm's t1 / t2):
    // half the dependent-issue latency of one seven-term chain
    const double w = z * z;
    double t1 = XRT_TAB(kLog)[5];
    double t2 = XRT_TAB(kLog)[6];
    t1 = fm(t1, w, XRT_TAB(kLog)[3]); t2 = fm(t2, w, XRT_TAB(kLog)[4]);
    t1 = fm(t1, w, XRT_TAB(kLog)[1]); t2 = fm(t2, w, XRT_TAB(kLog)[2]);
    t1 = t1 * w;                      t2 = fm(t2, w, XRT_TAB(kLog)[0]);
    double R = fm(t2, z, t1);
    const double hfsq = 0.5 * f * f;
    return fm(dk, XRT_TAB(kMisc)[1], -((hfsq - fm(s, hfsq + R, dk * XRT_TAB(kMisc)[2])) - f));
}

// sum_{i < N} c[i] t^i as two interleaved Horner chains in t^2 (even and odd powers)
template <int N>
XRT_HD double poly_even_odd(const double *c, double t) {
    const double t2 = t * t;
    double pe = c[(N - 1) & ~1], po = c[((N - 2) & ~1) + 1];
#pragma unroll
    for (int i = ((N - 1) & ~1) - 2; i >= 0; i -= 2) pe = fm(pe, t2, c[i]);
#pragma unroll
    for (int i = ((N - 2) & ~1) - 1; i >= 1; i -= 2) po = fm(po, t2, c[i]);
    return fm(po, t, pe);
}

// z with Phi(z) = u for u in [2^-54, 1 - 2^-53]: z = sqrt(2) x erfinv(x) / x with x = 2u - 1 (exact for u >= 1/4) and
// w = -log(1 - x^2) = -log(4 u (1 - u)) formed from u itself (1 - u is exact above 1/2, u carries the tail below it).
// 4e-16 relative against 60-digit arithmetic; one branch for 99.9 % of the uniforms (CUDA's normcdfinv: 160 instructions
// per call, 14 - 17 % of the exact kernels).
XRT_HD double inv_normal_cdf(double u) {
    const double x = fm(2.0, u, -1.0);
    double w = -log_pos(4.0 * (u * (1.0 - u)));
    w = w > 0.0 ? w : 0.0;
    double g;
    if (w < 6.25) {
        g = poly_even_odd<25>(XRT_TAB(kNinvC), w - 3.125);
    } else {
        const double sw = sqrt(w);
        g = (w < 16.0) ? poly_even_odd<21>(XRT_TAB(kNinvT1), sw - 3.25) : poly_even_odd<20>(XRT_TAB(kNinvT2), sw - 5.01);
    }
    return 1.4142135623730951 * x * g;
}

// exp(-x) for 0 <= x <= 700
XRT_HD double exp_neg(double x) {
    const double y = -x;
    const double kf = round_even(y * XRT_TAB(kMisc)[3]);
    double r = fm(kf, -XRT_TAB(kMisc)[1], y);
    r = fm(kf, -XRT_TAB(kMisc)[2], r);
    // sum_{k<=13} r^k / k! as two interleaved Horner chains in r^2 (even and odd powers)
    const double r2 = r * r;
    double pe = XRT_TAB(kExp)[1];      // 1/12!
    double po = XRT_TAB(kExp)[0];      // 1/13!
    pe = fm(pe, r2, XRT_TAB(kExp)[3]);  po = fm(po, r2, XRT_TAB(kExp)[2]);    // 1/10!, 1/11!
    pe = fm(pe, r2, XRT_TAB(kExp)[5]);  po = fm(po, r2, XRT_TAB(kExp)[4]);    // 1/8!,  1/9!
    pe = fm(pe, r2, XRT_TAB(kExp)[7]);  po = fm(po, r2, XRT_TAB(kExp)[6]);    // 1/6!,  1/7!
    pe = fm(pe, r2, XRT_TAB(kExp)[9]);  po = fm(po, r2, XRT_TAB(kExp)[8]);    // 1/4!,  1/5!
    pe = fm(pe, r2, XRT_TAB(kExp)[11]); po = fm(po, r2, XRT_TAB(kExp)[10]);   // 1/2!,  1/3!
    pe = fm(pe, r2, 1.0);               po = fm(po, r2, 1.0);                 // 1,     r^1 coefficient
    double p = fm(po, r, pe);
    const int32_t k = (int32_t)kf;
    return from_words(hi_word(p) + k * 1048576, lo_word(p));   // p in [0.7, 1.42], result normal for x <= 700
}

// asin(w) for the difference of two angles: series for the small arguments that matter
// (|w| < 0.05 covers mosaic spreads of a few degrees), the library call otherwise.
// Next term (231/13312) w^13 is below 5e-18 relative at the limit.
XRT_HD double asin_small(double w) {
    const double z = w * w;
    double p = 63.0 / 2816.0;
    p = fm(p, z, 35.0 / 1152.0);
    p = fm(p, z, 15.0 / 336.0);
    p = fm(p, z, 3.0 / 40.0);
    p = fm(p, z, 1.0 / 6.0);
    return fm(w * z, p, w);
}

}  // namespace xrt
