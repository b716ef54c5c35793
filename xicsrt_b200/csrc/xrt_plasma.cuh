// xrt_plasma.cuh -- the per-iteration bundle table of a plasma source, one thread per bundle.
//
// Restates, per bundle, what the reference does in whole-array numpy plus a Python loop over
// bundles (0.5 ms each): setup_bundles (_XicsrtPlasmaGeneric.py:176-231), bundle_filter
// (:246-250, _XicsrtBundleFilterSightline.py:31-56), bundle_generate of the Cubic / Toroidal /
// ToroidalDatafile classes, the intensity of create_sources (:301-319) and the ray count each
// per-bundle XicsrtSourceFocused draws at initialize (_XicsrtSourceGeneric.py:188-196).
#pragma once
#include "../../include/xrt.h"
#include "xrt_math.cuh"

namespace xrt {

constexpr uint32_t kSiteBundle = 0xB0000000u;     // Philox counter word 2: disjoint from every ray draw site

// np.interp(x, xp, fp, left=0, right=0)
__device__ __forceinline__ double interp_zero_outside(double x, const double *xp, const double *fp, int n) {
    if (!(x >= __ldg(xp)) || !(x <= __ldg(xp + n - 1))) return (x == x) ? 0.0 : CUDART_NAN;
    return interp_inside(x, xp, fp, n);
}

// Poisson variate from counter-based uniforms: inversion by sequential search for small means,
// Hoermann's transformed rejection (PTRS, the algorithm numpy uses) otherwise.
struct BundleUniforms {
    uint2 key;
    uint32_t lo, hi, stream, next;
    double spare;
    bool has_spare;
    __device__ __forceinline__ double draw() {
        if (has_spare) { has_spare = false; return spare; }
        uint4 r = philox4x32_10(make_uint4(lo, hi, kSiteBundle | next, stream), key);
        ++next;
        spare = u01(r.z, r.w);
        has_spare = true;
        return u01(r.x, r.y);
    }
};

__device__ __forceinline__ long long poisson_draw(double lam, BundleUniforms &g) {
    if (!(lam > 0.0)) return 0;
    if (lam < 10.0) {
        const double u = g.draw();
        double p = exp(-lam), cdf = p;
        long long k = 0;
        while (u > cdf && k < 200) {
            ++k;
            p *= lam / (double)k;
            cdf += p;
        }
        return k;
    }
    const double slam = sqrt(lam), loglam = log(lam);
    const double b = 0.931 + 2.53 * slam;
    const double a = -0.059 + 0.02483 * b;
    const double invalpha = 1.1239 + 1.1328 / (b - 3.4);
    const double vr = 0.9277 - 3.6224 / (b - 2.0);
    for (int attempt = 0; attempt < 1000; ++attempt) {
        const double U = g.draw() - 0.5;
        const double V = g.draw();
        const double us = 0.5 - fabs(U);
        const double kf = floor((2.0 * a / us + b) * U + lam + 0.43);
        if (us >= 0.07 && V <= vr) return (long long)kf;
        if (kf < 0.0 || (us < 0.013 && V > us)) continue;
        if ((log(V) + log(invalpha) - log(a / (us * us) + b)) <= (-lam + kf * loglam - lgamma(kf + 1.0)))
            return (long long)kf;
    }
    return (long long)floor(lam);
}

__global__ void __launch_bounds__(256)
k_bundles(const XrtPlasmaDesc p, const uint64_t seed, const uint64_t stream_id, const uint64_t n,
          XrtBundle *__restrict__ table, double *__restrict__ intensity, long long *__restrict__ counts) {
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        BundleUniforms g;
        g.key = make_uint2((uint32_t)seed, (uint32_t)(seed >> 32) ^ (uint32_t)(stream_id >> 32));
        g.stream = (uint32_t)stream_id;
        g.lo = (uint32_t)i;
        g.hi = (uint32_t)(i >> 32);
        g.next = 0;
        g.has_spare = false;

        // ---- centre: uniform in the box, then to external coordinates (:193-199)
        double u0, u1, u2;
        if (p.inject_u) { u0 = p.inject_u[i]; u1 = p.inject_u[n + i]; u2 = p.inject_u[2 * n + i]; }
        else { u0 = g.draw(); u1 = g.draw(); u2 = g.draw(); g.has_spare = false; }
        const double hx = -1.0 * p.size[0] / 2, hy = -1.0 * p.size[1] / 2, hz = -1.0 * p.size[2] / 2;
        const V3 off = v3(hx + (p.size[0] / 2 - hx) * u0, hy + (p.size[1] / 2 - hy) * u1, hz + (p.size[2] / 2 - hz) * u2);
        const V3 org = to_external(p.orient, off) + v3(p.origin);

        // ---- emission cone (:206-231)
        double spread = p.spread;
        if (p.use_spread_radius) {
            const V3 t = org - v3(p.target);
            spread = atan(p.spread_radius / sqrt(dot(t, t)));
        }
        const double sh = sin(spread / 2);
        const double solid_angle = 4.0 * CUDART_PI * sh * sh;

        // ---- bundle filters
        bool keep = true;
        for (int f = 0; f < p.n_sightlines; ++f) {
            const XrtSightline &sl = p.sightlines[f];
            const V3 l0 = v3(sl.origin) - org;
            const V3 ax = v3(sl.axis);
            const V3 perp = l0 - ax * dot(ax, l0);
            keep = keep && (sl.radius >= sqrt(dot(perp, perp)));
        }

        // ---- plasma parameters at the bundle
        double temperature = 1.0, emissivity = 1.0;
        V3 vel = v3(0.0, 0.0, 0.0);
        if (p.kind == XRT_PLASMA_CUBIC) {
            temperature = p.temperature;
            emissivity = p.emissivity;
        } else if (p.kind == XRT_PLASMA_TOROIDAL || p.kind == XRT_PLASMA_DATAFILE) {
            if (keep) {
                // normalised minor radius; the division by minor_radius (not its square) is the reference's
                const V3 q = org - v3(p.torus_origin);
                const double dd = sqrt(q.x * q.x + q.y * q.y) - p.major_radius;
                const double rr = sqrt(q.z * q.z + dd * dd);
                const double rho = sqrt(rr * rr / p.minor_radius);
                double t = p.temperature, e = p.emissivity;
                if (p.kind == XRT_PLASMA_DATAFILE) {
                    t = interp_zero_outside(rho, p.profile_t_rho, p.profile_t_val, p.n_profile_t);
                    e = interp_zero_outside(rho, p.profile_e_rho, p.profile_e_val, p.n_profile_e);
                }
                temperature = t * p.temperature_scale;
                emissivity = e * p.emissivity_scale;
                vel = v3(p.velocity) * p.velocity_scale;
                keep = isfinite(temperature);
            }
        }

        // ---- expected photons and the ray count
        const double inten = emissivity * p.intensity_factor * solid_angle;
        long long count = 0;
        if (keep) count = p.use_poisson ? poisson_draw(inten, g) : (long long)inten;
        if (count < 0) count = 0;

        XrtBundle b;
        b.origin[0] = org.x; b.origin[1] = org.y; b.origin[2] = org.z;
        // cone parameter of the per-bundle source (scene.py:_fill_cone for a scalar spread)
        b.cos_spread = p.cone == XRT_CONE_ISOTROPIC ? cos(spread) : p.cone == XRT_CONE_ISOTROPIC_XY ? sin(spread) : tan(spread);
        // natural linewidth > 0: a bundle at exactly T = 0 is given 1 eV (_XicsrtSourceGeneric.py:333-339)
        const double t_line = (p.thermal_line == 2 && temperature == 0.0) ? 1.0 : temperature;
        b.wave_sigma = (p.thermal_line && t_line > 0.0) ? sqrt(t_line) * p.sigma_factor : 0.0;
        b.velocity_c[0] = vel.x * p.inv_c; b.velocity_c[1] = vel.y * p.inv_c; b.velocity_c[2] = vel.z * p.inv_c;
        table[i] = b;
        intensity[i] = keep ? inten : -1.0;
        counts[i] = count;
    }
}

// ---------------------------------------------------------------------------
// ray id -> bundle hint: hint[i] = first bundle whose inclusive prefix sum exceeds i << shift
// (i.e. the bundle of that ray id); hint[n_buckets] = n_bundles - 1.
__global__ void __launch_bounds__(256)
k_bundle_hint(const uint64_t *__restrict__ end, const uint64_t n_bundles, const int shift, const uint64_t n_buckets,
              uint32_t *__restrict__ hint) {
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i <= n_buckets; i += stride) {
        uint64_t lo = 0, hi = n_bundles - 1;
        if (i < n_buckets) {
            const uint64_t index = i << shift;
            while (lo < hi) {
                const uint64_t mid = (lo + hi) >> 1;
                if (__ldg(end + mid) > index) hi = mid; else lo = mid + 1;
            }
        } else {
            lo = hi;
        }
        hint[i] = (uint32_t)lo;
    }
}

// ---------------------------------------------------------------------------
// Per-bundle Voigt inverse-CDF tables (xicsrt/tools/xicsrt_voigt.py:30-92).
//
// The reference builds one 1000-bin table per bundle source with scipy's Faddeeva function.
// Here Re w(x + iy), y > 0, comes from the trapezoid rule for (i/pi) Int exp(-t^2) / (z - t) dt
// with step h = 1/2, whose error is exp(-pi^2 / h^2) = 7e-18 once the pole term
// 2 exp(-z^2) / (1 -+ exp(-2 pi i z / h)) is added for y < pi / h (Matta & Reichel 1971).  Of
// the two grids t = n h and t = (n + 1/2) h the one whose nodes are farther from x is used, so
// the pole term never cancels against a nearby node (checked against scipy.special.wofz to
// 1.2e-13 relative over |x| < 30, 1e-12 < y < 100: tests/test_plasma.py).
constexpr int kVoigtNodes = 13;          // nodes -13 .. 13 (+ 1/2): exp(-6.5^2) = 4e-19
constexpr double kVoigtH = 0.5;

__device__ __forceinline__ double faddeeva_re(double x, double y, const double *__restrict__ node_w) {
    const double q = x / kVoigtH;
    const double fr = q - floor(q);
    const bool half = (fr < 0.25) || (fr > 0.75);
    const double *w = node_w + (half ? 2 * kVoigtNodes + 1 : 0);
    const double t0 = half ? 0.5 * kVoigtH : 0.0;
    const double y2 = y * y;
    double sum = 0.0;
#pragma unroll 1
    for (int k = -kVoigtNodes; k <= kVoigtNodes; ++k) {
        const double a = x - (t0 + k * kVoigtH);
        sum += w[k + kVoigtNodes] * (y / fma(a, a, y2));
    }
    double res = sum * (kVoigtH / CUDART_PI);
    if (y < CUDART_PI / kVoigtH) {
        // 2 exp(-z^2) / (1 - sgn E (cos th - i sin th)),  E = exp(2 pi y / h), th = 2 pi x / h
        const double mag = 2.0 * exp(y2 - x * x);
        if (mag > 0.0) {
            double s2, c2, st, ct;
            sincos(2.0 * x * y, &s2, &c2);
            sincospi(2.0 * x / kVoigtH, &st, &ct);
            const double E = (half ? -1.0 : 1.0) * exp(2.0 * CUDART_PI * y / kVoigtH);
            const double dr = 1.0 - E * ct, di = E * st;          // denominator
            // Re[(c2 - i s2) * conj(dr + i di)] / |d|^2
            res += mag * (c2 * dr - s2 * di) / (dr * dr + di * di);
        }
    }
    return res;
}

// one block per bundle (grid-stride); thread t owns bins 4t .. 4t+3 of the n_table bins
constexpr int kVoigtBlock = 256;

__global__ void __launch_bounds__(kVoigtBlock)
k_voigt_tables(const XrtBundle *__restrict__ table, const long long *__restrict__ counts, const uint64_t n_bundles,
               const double gamma, const int n_table, double *__restrict__ x_out, double *__restrict__ cdf_out) {
    __shared__ double node_w[2 * (2 * kVoigtNodes + 1)];
    __shared__ double warp_sum[kVoigtBlock / 32];
    for (int i = threadIdx.x; i < 2 * (2 * kVoigtNodes + 1); i += blockDim.x) {
        const int g = i / (2 * kVoigtNodes + 1), k = i % (2 * kVoigtNodes + 1) - kVoigtNodes;
        const double t = (k + 0.5 * g) * kVoigtH;
        node_w[i] = exp(-t * t);
    }
    __syncthreads();
    const int per = (n_table + kVoigtBlock - 1) / kVoigtBlock;
    for (uint64_t b = blockIdx.x; b < n_bundles; b += gridDim.x) {
        if (counts[b] <= 0) continue;                 // uniform over the block
        const double sigma = table[b].wave_sigma;
        // grid (:52-72): `value` = 50 * hwhm_max / 5, stretched so that the last edge is the cutoff
        const double cutoff = 1e-5;
        const double gauss_hw = sqrt(2.0 * log(2.0)) * sigma;
        const double hw_max = sqrt(gauss_hw * gauss_hw + gamma * gamma);
        const double value = 100 / 2 * (hw_max / 5.0);
        const double lorentz_cut = gamma * sqrt(1.0 / cutoff - 1.0);
        const double gauss_cut = sqrt(-1.0 * sigma * sigma * 2.0 * log(cutoff * sigma * sqrt(2.0 * CUDART_PI)));
        const double ln_base = 0.1 * log(fmax(lorentz_cut, gauss_cut) / value);
        const double step = (value - (-value)) / n_table;
        const double inv_s2 = 1.0 / (sqrt(2.0) * sigma);
        const double norm = 1.0 / (sqrt(2.0 * CUDART_PI) * sigma);
        auto edge = [&](int k) {
            const double lin = (k == n_table) ? value : -value + k * step;        // np.linspace
            return lin * exp(ln_base * fabs(lin / value * 10.0));
        };
        const int k0 = threadIdx.x * per;
        double part[8];
        double run = 0.0;
        double e_lo = (k0 < n_table) ? edge(k0) : 0.0;
        for (int j = 0; j < per; ++j) {
            const int k = k0 + j;
            if (k < n_table) {
                const double e_hi = edge(k + 1);
                const double mid = (e_lo + e_hi) / 2;
                const double pdf = faddeeva_re(mid * inv_s2, gamma * inv_s2, node_w) * norm;
                run += pdf * (e_hi - e_lo);
                x_out[b * n_table + k] = e_hi;
                e_lo = e_hi;
            }
            if (j < 8) part[j] = run;
        }
        // exclusive block scan of the per-thread totals
        double incl = run;
        const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
        for (int o = 1; o < 32; o <<= 1) {
            const double v = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += v;
        }
        if (lane == 31) warp_sum[wid] = incl;
        __syncthreads();
        double before = incl - run;
        for (int w = 0; w < wid; ++w) before += warp_sum[w];
        __syncthreads();
        for (int j = 0; j < per && j < 8; ++j) {
            const int k = k0 + j;
            if (k < n_table) cdf_out[b * n_table + k] = before + part[j];
        }
    }
}

}  // namespace xrt
