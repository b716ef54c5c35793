// xrt_math.cuh -- small FP64 vector helpers, Philox4x32-10 and table lookup.
#pragma once
#include <stdint.h>
#include <math_constants.h>
#include "xrt_fastmath.cuh"

namespace xrt {

struct V3 { double x, y, z; };

__device__ __forceinline__ V3 v3(double x, double y, double z) { V3 r; r.x = x; r.y = y; r.z = z; return r; }
__device__ __forceinline__ V3 v3(const double *p) { return v3(p[0], p[1], p[2]); }
__device__ __forceinline__ V3 operator+(V3 a, V3 b) { return v3(a.x + b.x, a.y + b.y, a.z + b.z); }
__device__ __forceinline__ V3 operator-(V3 a, V3 b) { return v3(a.x - b.x, a.y - b.y, a.z - b.z); }
__device__ __forceinline__ V3 operator*(V3 a, double s) { return v3(a.x * s, a.y * s, a.z * s); }
__device__ __forceinline__ V3 operator*(double s, V3 a) { return v3(a.x * s, a.y * s, a.z * s); }
__device__ __forceinline__ double dot(V3 a, V3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
__device__ __forceinline__ V3 cross(V3 a, V3 b) {
    return v3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x);
}
// Square root, reciprocal square root and reciprocal without the library's special-case path.
// The library versions add a range check, a branch and an out-of-line slow path for subnormal /
// huge arguments to every call; the ray code only meets normal-range values (lengths of order
// one, 1 - x^2 with |x| <= 1), for which the hardware seed (MUFU, ~2^-22 relative) plus one
// third-order step and a residual correction is correctly rounded to within an ulp.
//   x < 0 or NaN -> NaN (as sqrt), x == 0 -> 0 for the square root.
__device__ __forceinline__ double rsqrt_seed(double x) {
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    return y;
}
__device__ __forceinline__ double fast_rsqrt(double x) {
    const double y = rsqrt_seed(x);
    const double e = fma(-x * y, y, 1.0);                       // 1 - x y^2
    return fma(y * e, fma(e, 0.375, 0.5), y);                   // y (1 + e/2 + 3 e^2 / 8)
}
__device__ __forceinline__ double fast_sqrt(double x) {
    const double y = fast_rsqrt(x);
    double g = x * y;
    g = fma(fma(-g, g, x), 0.5 * y, g);                         // residual correction
    return (x == 0.0) ? 0.0 : g;
}
__device__ __forceinline__ V3 unit(V3 a) { return a * fast_rsqrt(dot(a, a)); }
__device__ __forceinline__ V3 nan3() { return v3(CUDART_NAN, CUDART_NAN, CUDART_NAN); }

// rows of a 3x3 orientation (x, y, z axes of an element)
// local = R . v   (reference _GeometryObject.py:156-168, einsum 'ji,ki->kj')
__device__ __forceinline__ V3 to_local(const double *R, V3 v) {
    return v3(R[0] * v.x + R[1] * v.y + R[2] * v.z,
              R[3] * v.x + R[4] * v.y + R[5] * v.z,
              R[6] * v.x + R[7] * v.y + R[8] * v.z);
}
// external = R^T . v  (reference _GeometryObject.py:143-154, einsum 'ij,ki->kj')
__device__ __forceinline__ V3 to_external(const double *R, V3 v) {
    return v3(R[0] * v.x + R[3] * v.y + R[6] * v.z,
              R[1] * v.x + R[4] * v.y + R[7] * v.z,
              R[2] * v.x + R[5] * v.y + R[8] * v.z);
}

// ---------------------------------------------------------------------------
// Philox4x32-10 (Salmon et al. 2011), counter-based: no state, any ray can be
// regenerated from (key, ray id, draw site).

__device__ __forceinline__ uint4 philox4x32_10(uint4 c, uint2 k) {
    const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u;
    const uint32_t W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        uint32_t hi0 = __umulhi(M0, c.x), lo0 = M0 * c.x;
        uint32_t hi1 = __umulhi(M1, c.z), lo1 = M1 * c.z;
        c = make_uint4(hi1 ^ c.y ^ k.x, lo1, hi0 ^ c.w ^ k.y, lo0);
        k.x += W0;
        k.y += W1;
    }
    return c;
}

// The same generator with the ten round keys (k + r W) precomputed on the host: the key is the
// same for every ray of a launch, so the kernels take the schedule as a __grid_constant__
// parameter and the XORs read it straight from the constant bank (18 integer adds per block less).
struct PhiloxKeys { uint32_t rk[20]; };

__host__ __device__ inline void philox_round_keys(uint64_t seed, uint64_t stream_id, PhiloxKeys &K) {
    uint32_t kx = (uint32_t)seed, ky = (uint32_t)(seed >> 32) ^ (uint32_t)(stream_id >> 32);
    for (int r = 0; r < 10; ++r) {
        K.rk[2 * r] = kx;
        K.rk[2 * r + 1] = ky;
        kx += 0x9E3779B9u;
        ky += 0xBB67AE85u;
    }
}

#ifndef XRT_PHILOX_ROUNDS
#define XRT_PHILOX_ROUNDS 10
#endif
__device__ __forceinline__ uint4 philox4x32_10(uint4 c, const PhiloxKeys &K) {
    const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u;
#pragma unroll
    for (int r = 0; r < XRT_PHILOX_ROUNDS; ++r) {
        uint32_t hi0 = __umulhi(M0, c.x), lo0 = M0 * c.x;
        uint32_t hi1 = __umulhi(M1, c.z), lo1 = M1 * c.z;
        c = make_uint4(hi1 ^ c.y ^ K.rk[2 * r], lo1, hi0 ^ c.w ^ K.rk[2 * r + 1], lo0);
    }
    return c;
}

// Uniform in [0, 1) from two 32-bit words: the 52 mantissa bits of a double in [1, 2) are
// filled with random bits and 1 is subtracted (two integer ops and one DADD; the
// int -> double conversions of the textbook construction run on the quarter-rate XU pipe).
__device__ __forceinline__ double u01(uint32_t a, uint32_t b) {
    return __hiloint2double((int)(0x3ff00000u | (a >> 12)), (int)((a << 20) | (b >> 12))) - 1.0;
}
// three uniforms of 42 bits each from one 128-bit block (a box of 1 m is resolved to 2e-13 m)
__device__ __forceinline__ void u01_42x3(uint4 r, double &u0, double &u1, double &u2) {
    u0 = __hiloint2double((int)(0x3ff00000u | (r.x >> 12)), (int)(((r.x & 0xfffu) << 20) | ((r.y >> 22) << 10))) - 1.0;
    u1 = __hiloint2double((int)(0x3ff00000u | ((r.y >> 2) & 0xfffffu)), (int)(((r.y & 3u) << 30) | ((r.z >> 12) << 10))) - 1.0;
    u2 = __hiloint2double((int)(0x3ff00000u | ((r.z & 0xfffu) << 8) | (r.w >> 24)), (int)((r.w & 0x00fffffcu) << 8)) - 1.0;
}
// 44 random bits: the low 12 bits of a and all of b
__device__ __forceinline__ double u01_44(uint32_t a, uint32_t b) {
    return __hiloint2double((int)(0x3ff00000u | ((a & 0xfffu) << 8) | (b >> 24)), (int)(b << 8)) - 1.0;
}
// the same with 40 random bits (a: 32, top 8 of b) and with the low 24 bits of b
__device__ __forceinline__ double u01_40(uint32_t a, uint32_t b) {
    return __hiloint2double((int)(0x3ff00000u | (a >> 12)), (int)((a << 20) | ((b >> 24) << 12))) - 1.0;
}
__device__ __forceinline__ double u01_24(uint32_t b) {
    return __hiloint2double((int)(0x3ff00000u | ((b & 0xffffffu) >> 4)), (int)(b << 28)) - 1.0;
}

// two independent standard normals from two uniforms (Box-Muller); 1 - u1 is in (0, 1]
__device__ __forceinline__ void box_muller(double u1, double u2, double &z1, double &z2) {
    double r = fast_sqrt(-2.0 * log_pos(1.0 - u1));
    double s, c;
    sincos_2pi(u2, s, c);
    z1 = r * c;
    z2 = r * s;
}

// np.interp(x, xp, fp) for x inside [xp[0], xp[n-1]] (caller handles the outside)
__device__ __forceinline__ double interp_inside(double x, const double *__restrict__ xp,
                                                const double *__restrict__ fp, int n) {
    // largest j with xp[j] <= x
    int lo = 0, hi = n - 1;
    while (hi - lo > 1) {
        int mid = (lo + hi) >> 1;
        if (__ldg(xp + mid) <= x) lo = mid; else hi = mid;
    }
    double x0 = __ldg(xp + lo), x1 = __ldg(xp + lo + 1);
    double f0 = __ldg(fp + lo), f1 = __ldg(fp + lo + 1);
    if (x == x1) return f1;
    double slope = (f1 - f0) / (x1 - x0);
    return slope * (x - x0) + f0;
}

}  // namespace xrt
