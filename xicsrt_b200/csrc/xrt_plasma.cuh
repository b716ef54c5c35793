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
        b.cos_spread = cos(spread);
        b.wave_sigma = (p.thermal_line && temperature > 0.0) ? sqrt(temperature) * p.sigma_factor : 0.0;
        b.velocity_c[0] = vel.x * p.inv_c; b.velocity_c[1] = vel.y * p.inv_c; b.velocity_c[2] = vel.z * p.inv_c;
        table[i] = b;
        intensity[i] = keep ? inten : -1.0;
        counts[i] = count;
    }
}

}  // namespace xrt
