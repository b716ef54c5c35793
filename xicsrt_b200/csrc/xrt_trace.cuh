// xrt_trace.cuh -- per-ray device code: source generation and the optic train.
//
// One thread owns one ray; the ray state (origin, direction, wavelength, alive
// flag) stays in registers from the source through every optic.  The scene is a
// __grid_constant__ kernel parameter, so element parameters are read from the
// constant bank with warp-uniform addresses.
//
// Feature mask FT: compile-time pruning of branches a scene does not use, so the
// common spectrometer (plane / sphere optics, Gaussian or step rocking curve)
// does not pay registers for the torus solver, mosaic loop or mesh lookup.
#pragma once
#include "../../include/xrt.h"
#include "xrt_math.cuh"
#include "xrt_fastmath.cuh"

namespace xrt {

enum : uint32_t {
    FT_LOCAL = 1u << 0,     // some optic traces in local coordinates
    FT_CYL = 1u << 1,
    FT_TORUS = 1u << 2,
    FT_MESH = 1u << 3,
    FT_APERTURE = 1u << 4,
    FT_MOSAIC = 1u << 5,
    FT_ROCKTAB = 1u << 6,
    FT_SRC_EXT = 1u << 7,   // anything beyond box + isotropic cone + const/uniform/normal line
    FT_MID = FT_LOCAL | FT_CYL | FT_TORUS | FT_APERTURE | FT_MOSAIC | FT_ROCKTAB | FT_SRC_EXT,
    FT_FULL = FT_MID | FT_MESH,
    FT_MESHLEAN = FT_MESH | FT_LOCAL,   // mesh optics (traced in local coordinates) + plane / sphere, box source
    FT_MOSAICLEAN = FT_MOSAIC,          // mosaic crystal + plane / sphere, box source (config 3)
    FT_SRCLEAN = FT_SRC_EXT,            // plasma bundles / extended sources + plane / sphere (config 5)
};

// Known-structure mask KN: facts about the source and the split optic that a pre-instantiated
// kernel variant may assume at compile time (the host checks them at scene creation).  The
// generic code reads the same facts from the descriptor; with a KN bit set the accessor
// returns the constant, so the branch on it -- and the code of the other cases -- disappears.
enum : uint32_t {
    KN_POINT_SOURCE = 1u << 0,   // source.kind == FIXED_AXIS and a zero-size origin box
    KN_WAVE_NORMAL = 1u << 1,    // source.wave == XRT_WAVE_NORMAL
    KN_SPHERE = 1u << 2,         // split optic: concave sphere
    KN_BOUNDS_XY = 1u << 3,      // split optic: check_size with xsize and ysize, no zsize
    KN_CRYSTAL_GAUSS = 1u << 4,  // split optic: crystal with Bragg test, Gaussian or step rocking curve (read at run time)
    KN_IMAGE = 1u << 5,          // split optic: has a pixel grid
    KN_SPECTROMETER = KN_POINT_SOURCE | KN_WAVE_NORMAL | KN_SPHERE | KN_BOUNDS_XY | KN_CRYSTAL_GAUSS | KN_IMAGE,
};

template <uint32_t KN> __device__ __forceinline__ int shape_of(const XrtOpticDesc &op) {
    if constexpr ((KN & KN_SPHERE) != 0) return XRT_SHAPE_SPHERE; else return op.shape;
}
template <uint32_t KN> __device__ __forceinline__ int interact_of(const XrtOpticDesc &op) {
    if constexpr ((KN & KN_CRYSTAL_GAUSS) != 0) return XRT_INTERACT_CRYSTAL; else return op.interact;
}
template <uint32_t KN> __device__ __forceinline__ int rocking_of(const XrtOpticDesc &op) {
    // known to be GAUSS or STEP, never a table: the choice costs one uniform branch in stage B2 (9 % of the rays)
    if constexpr ((KN & KN_CRYSTAL_GAUSS) != 0) return op.rocking_type == XRT_ROCK_STEP ? XRT_ROCK_STEP : XRT_ROCK_GAUSS;
    else return op.rocking_type;
}
template <uint32_t KN> __device__ __forceinline__ uint32_t flags_of(const XrtOpticDesc &op) {
    uint32_t f = op.flags;
    if constexpr ((KN & KN_SPHERE) != 0) f &= ~(uint32_t)XRT_F_CONVEX;
    if constexpr ((KN & KN_BOUNDS_XY) != 0)
        f = (f | XRT_F_CHECK_SIZE | XRT_F_HAS_XSIZE | XRT_F_HAS_YSIZE) & ~(uint32_t)XRT_F_HAS_ZSIZE;
    if constexpr ((KN & KN_CRYSTAL_GAUSS) != 0) f |= XRT_F_CHECK_BRAGG;
    if constexpr ((KN & KN_IMAGE) != 0) f |= XRT_F_IMAGE;
    return f;
}

struct Ray {
    V3 o, d;
    double w;
    bool alive;
};

// ---------------------------------------------------------------------------
// draw providers

// draw-site numbering for the Philox counter (word 2)
__device__ __forceinline__ uint32_t site_optic(int k, int layer, int which) {
    return ((uint32_t)(k + 1) << 16) | ((uint32_t)layer << 1) | (uint32_t)which;
}
// retries of the isotropic_xy rejection sampler live in their own tagged range (bit 30), disjoint from the optic
// sites ((k + 1) << 16 | ...) for any number of attempts and from the plasma bundle sites (0xB0000000 | ...)
enum : uint32_t { SITE_ORIGIN_XY = 0, SITE_ORIGIN_Z = 1, SITE_CONE = 2, SITE_WAVE = 3,
                  SITE_LOSTKEY = 4, SITE_CONE_RETRY = 0x40000000u };

struct PhiloxDraws {
    const PhiloxKeys *keys;   // round keys of (seed, stream_id): a __grid_constant__ kernel parameter
    uint32_t ray_lo, ray_hi, stream;
    int k_shared;     // the crystal whose first rocking-curve uniform shares the wavelength's Philox block
    mutable uint32_t spare;   // word w of the cone block (valid after cone(0, ..)): top 32 bits of the wavelength uniform

    __device__ __forceinline__ void init(const PhiloxKeys &K, uint64_t stream_id, uint64_t ray, int shared_optic) {
        keys = &K;
        spare = 0;
        stream = (uint32_t)stream_id;
        ray_lo = (uint32_t)ray;
        ray_hi = (uint32_t)(ray >> 32);
        k_shared = shared_optic;
    }
    __device__ __forceinline__ uint4 raw(uint32_t site) const {
        return philox4x32_10(make_uint4(ray_lo, ray_hi, site, stream), *keys);
    }
    __device__ __forceinline__ void pair(uint32_t site, double &a, double &b) const {
        uint4 r = raw(site);
        a = u01(r.x, r.y);
        b = u01(r.z, r.w);
    }
    // ---- source sites
    __device__ __forceinline__ void origin_uniform(double u[3]) const {
        u01_42x3(raw(SITE_ORIGIN_XY), u[0], u[1], u[2]);     // one block: 3 x 42 bits
    }
    __device__ __forceinline__ void origin_gauss(const double sig[3], double off[3]) const {
        double a, b, c, d, z0, z1, z2, z3;
        pair(SITE_ORIGIN_XY, a, b);
        pair(SITE_ORIGIN_Z, c, d);
        box_muller(a, b, z0, z1);
        box_muller(c, d, z2, z3);
        off[0] = sig[0] * z0; off[1] = sig[1] * z1; off[2] = sig[2] * z2;
    }
    // Cone block: words x, y(top 20 bits) -> first uniform (52 bits); y(low 12 bits), z -> second uniform
    // (44 bits: 3.6e-13 rad of azimuth).  Word w of the attempt-0 block is left for the wavelength.
    __device__ __forceinline__ void cone(int attempt, double &a, double &b) const {
        uint4 r = raw(attempt == 0 ? SITE_CONE : SITE_CONE_RETRY + (uint32_t)attempt);
        a = u01(r.x, r.y);
        b = u01_44(r.y, r.z);
        if (attempt == 0) spare = r.w;
    }
    // Wavelength.  A normal deviate is the inverse normal CDF of ONE 53-bit uniform whose top 32
    // bits are word w of the cone block and whose low 21 bits are word x of the SITE_WAVE block,
    // taken at the centre of its 2^-53 bin (never 0 or 1; |z| < 8.3).  The fused kernel's Bragg
    // pre-test needs the deviate only roughly and gets it from the cone block alone -- no second
    // Philox block for the rays it rejects.  Words z, w of the SITE_WAVE block -> the 52-bit
    // rocking-curve uniform of (optic k_shared, layer 0); the block is a pure function of
    // (key, ray, site), so its users share one evaluation after inlining.
    __device__ __forceinline__ double wave_u() const { uint4 r = raw(SITE_WAVE); return u01(r.x, r.y); }
    __device__ __forceinline__ double wave_z() const {
        const uint32_t hi = raw(SITE_CONE).w;
        const uint32_t lo = raw(SITE_WAVE).x;
        return inv_normal_cdf(u01(hi, lo) + 1.1102230246251565e-16);
    }
    // top 32 bits of that uniform, for approximate deviates (valid after cone(0, ..) on this object)
    __device__ __forceinline__ uint32_t wave_hi() const { return spare; }
    __device__ __forceinline__ uint64_t lost_key() const {
        uint4 r = raw(SITE_LOSTKEY);
        return ((uint64_t)r.x << 32) | r.y;
    }
    // ---- optic sites
    __device__ __forceinline__ double bragg_u(int k, int layer) const {
        if (k == k_shared && layer == 0) { uint4 r = raw(SITE_WAVE); return u01(r.z, r.w); }
        double a, b; pair(site_optic(k, layer, 0), a, b); return a;
    }
    // Reflection probability below 2^-57: the ray passes only when the uniform is exactly 0
    // (p >= u).  With Philox that draw is not made at all (probability 1.1e-16 per ray).
    __device__ __forceinline__ bool bragg_u_is_zero(int, int) const { return false; }
    // One Philox block per mosaic layer: words x, y -> the two crystallite offsets (Box-Muller from a
    // 40-bit radius uniform and a 24-bit angle), words z, w -> the rocking-curve uniform of that layer.
    __device__ __forceinline__ void mosaic_xy(int k, int layer, double s, double &x, double &y) const {
        uint4 r = raw(site_optic(k, layer, 1));
        double z0, z1;
        box_muller(u01_40(r.x, r.y), u01_24(r.y), z0, z1);
        x = s * z0; y = s * z1;
    }
    __device__ __forceinline__ double mosaic_u(int k, int layer) const {
        uint4 r = raw(site_optic(k, layer, 1));
        return u01(r.z, r.w);
    }
};

// draws recorded from the reference-order stream, scattered to full length
struct InjectedDraws {
    const XrtInject *inj;
    uint64_t i, n;
    __device__ __forceinline__ double bragg_u(int k, int layer) const {
        return inj->u[k][(uint64_t)layer * n + i];
    }
    __device__ __forceinline__ bool bragg_u_is_zero(int k, int layer) const { return bragg_u(k, layer) == 0.0; }
    __device__ __forceinline__ double mosaic_u(int k, int layer) const { return bragg_u(k, layer); }
    __device__ __forceinline__ void mosaic_xy(int k, int layer, double, double &x, double &y) const {
        const double *p = inj->xy[k] + (uint64_t)layer * 2 * n;
        x = p[i]; y = p[n + i];
    }
};

struct SourceInjectedDraws {
    const XrtSourceInject *inj;
    uint64_t i, n;
    __device__ __forceinline__ void origin_uniform(double u[3]) const {
        u[0] = inj->origin[i]; u[1] = inj->origin[n + i]; u[2] = inj->origin[2 * n + i];
    }
    __device__ __forceinline__ void origin_gauss(const double *, double off[3]) const {
        off[0] = inj->origin[i]; off[1] = inj->origin[n + i]; off[2] = inj->origin[2 * n + i];
    }
    __device__ __forceinline__ void cone(int, double &a, double &b) const { a = inj->cone[i]; b = inj->cone[n + i]; }
    __device__ __forceinline__ double wave_u() const { return inj->wave[i]; }
    __device__ __forceinline__ double wave_z() const { return inj->wave[i]; }
};

// ---------------------------------------------------------------------------
// source: one ray from the draws  (reference _XicsrtSourceGeneric.py:198-393,
// _XicsrtSourceFocused.py:35-44, _XicsrtPlasmaGeneric.py:286-345)
//
// Split in two so that the fused kernel can draw the wavelength lazily: Philox is counter
// based, so a ray's wavelength can be produced at the first optic that needs it (the Bragg
// test) instead of at the source -- rays that miss the crystal never pay for it.

// per-ray source parameters: the source itself, or the plasma bundle the ray belongs to
struct SrcLocal {
    V3 org, vel;
    double cos_spread, wave_sigma, ext0, ext1, ext2;
    uint64_t bundle;      // plasma: index of the ray's bundle (row of the per-bundle wavelength tables)
};

template <uint32_t FT, uint32_t KN = 0>
__device__ __forceinline__ void source_local(const XrtSourceDesc &s, uint64_t index, SrcLocal &L) {
    L.org = v3(s.origin);
    L.vel = v3(s.velocity_c);
    L.cos_spread = s.cone_par[0];
    L.wave_sigma = s.wave_par[1];
    L.bundle = 0;
    if constexpr ((KN & KN_POINT_SOURCE) != 0) { L.ext0 = L.ext1 = L.ext2 = 0.0; }
    else { L.ext0 = s.extent[0]; L.ext1 = s.extent[1]; L.ext2 = s.extent[2]; }
    if constexpr ((FT & FT_SRC_EXT) != 0) {
        if (s.kind == XRT_SRC_BUNDLES) {
            // bundle of this ray: first b with bundle_end[b] > index
            uint64_t lo = 0, hi = s.n_bundles - 1;
            if (s.bundle_hint) {     // bracket from the per-2^shift-ids hint table: usually 0..2 steps left
                const uint64_t j = index >> s.bundle_hint_shift;
                lo = __ldg(s.bundle_hint + j);
                hi = __ldg(s.bundle_hint + j + 1);
            }
            while (lo < hi) {
                uint64_t mid = (lo + hi) >> 1;
                if (__ldg(s.bundle_end + mid) > index) hi = mid; else lo = mid + 1;
            }
            const XrtBundle *b = s.bundles + lo;
            L.bundle = lo;
            L.org = v3(__ldg(&b->origin[0]), __ldg(&b->origin[1]), __ldg(&b->origin[2]));
            L.cos_spread = __ldg(&b->cos_spread);
            L.wave_sigma = __ldg(&b->wave_sigma);
            L.vel = v3(__ldg(&b->velocity_c[0]), __ldg(&b->velocity_c[1]), __ldg(&b->velocity_c[2]));
            L.ext0 = L.ext1 = L.ext2 = s.voxel_size;
        }
    }
}

// origin, direction and the source-level mask
template <uint32_t FT, class DR, uint32_t KN = 0, bool USE_TABLE = false>
__device__ __forceinline__ void generate_geometry(const XrtSourceDesc &s, const SrcLocal &L, const DR &dr, Ray &r,
                                                  const double *sincos_table = nullptr) {
    // ---- origin (:229-255): three draws of U(-size/2, size/2), or N(0, sigma)
    double off[3] = {0.0, 0.0, 0.0};
    bool gauss = false;
    if constexpr ((FT & FT_SRC_EXT) != 0) gauss = (s.spatial == XRT_SPATIAL_GAUSSIAN);
    if (gauss) {
        double sig[3] = {L.ext0, L.ext1, L.ext2};
        dr.origin_gauss(sig, off);
    } else if (L.ext0 != 0.0 || L.ext1 != 0.0 || L.ext2 != 0.0) {
        double u[3];
        dr.origin_uniform(u);
        off[0] = -0.5 * L.ext0 + L.ext0 * u[0];
        off[1] = -0.5 * L.ext1 + L.ext1 * u[1];
        off[2] = -0.5 * L.ext2 + L.ext2 * u[2];
    }
    const double *R = s.orient;
    if constexpr ((KN & KN_POINT_SOURCE) != 0) {
        r.o = L.org;      // origin + 0 x + 0 y + 0 z: the same value (finite axes)
    } else {
        r.o = v3(((L.org.x + off[0] * R[0]) + off[1] * R[3]) + off[2] * R[6],
                 ((L.org.y + off[0] * R[1]) + off[1] * R[4]) + off[2] * R[7],
                 ((L.org.z + off[0] * R[2]) + off[1] * R[5]) + off[2] * R[8]);
    }

    // ---- local cone vector (xicsrt_spread.py:80-294)
    V3 l;
    bool xy_exhausted = false;      // isotropic_xy: no direction accepted in 100000 attempts (window of measure ~0)
    int cone = XRT_CONE_ISOTROPIC;
    if constexpr ((FT & FT_SRC_EXT) != 0) cone = s.cone;
    if (cone == XRT_CONE_ISOTROPIC) {
        // z ~ U(cos(spread), 1), phi ~ U(0, 2 pi): (rho cos phi, rho sin phi, z)
        double a, b;
        dr.cone(0, a, b);
        double z = L.cos_spread + (1.0 - L.cos_spread) * a;
        double rho = fast_sqrt(fma(-z, z, 1.0));
        double sn, cs;
        if constexpr (USE_TABLE) sincos_2pi_tab(b, sincos_table, sn, cs);   // fused kernel: shared-memory table
        else sincos_2pi(b, sn, cs);
        l = v3(rho * cs, rho * sn, z);
    } else if (cone == XRT_CONE_ISOTROPIC_XY) {
        // rejection from the enclosing circular cone (:130-196); a plasma bundle carries sin(spread)
        // of its scalar spread: limits [-v, v, -v, v], enclosing cone asin(sqrt(2 v^2))
        const bool per_bundle = s.kind == XRT_SRC_BUNDLES;
        const double v = L.cos_spread;
        const double cm = per_bundle ? sqrt(fma(-2.0 * v, v, 1.0)) : s.cone_cos_max;
        const double x_lo = per_bundle ? -v : s.cone_par[0], x_hi = per_bundle ? v : s.cone_par[1];
        const double y_lo = per_bundle ? -v : s.cone_par[2], y_hi = per_bundle ? v : s.cone_par[3];
        l = v3(0.0, 0.0, 1.0);
        xy_exhausted = true;
        for (int attempt = 0; attempt < 100000; ++attempt) {
            double a, b;
            dr.cone(attempt, a, b);
            double z = cm + (1.0 - cm) * a;
            double rho = sqrt(fma(-z, z, 1.0));
            double sn, cs;
            sincos_2pi(b, sn, cs);
            double x = rho * cs, y = rho * sn;
            double sx = x / sqrt(x * x + z * z);
            double sy = y / sqrt(y * y + z * z);
            l = v3(x, y, z);
            if (sx > x_lo && sx <= x_hi && sy > y_lo && sy <= y_hi) { xy_exhausted = false; break; }
        }
    } else {
        double a, b, a0, s1, c1;
        dr.cone(0, a, b);
        if (cone == XRT_CONE_FLAT) {
            // L.cos_spread holds tan(spread): cone_par[0] of the source, or the bundle's own
            double rr = sqrt(0.0 + (L.cos_spread - 0.0) * a);
            sincos_2pi(b, s1, c1);
            a0 = atan(rr);
        } else {
            const bool per_bundle = s.kind == XRT_SRC_BUNDLES;
            const double v = L.cos_spread;
            const double x_lo = per_bundle ? -v : s.cone_par[0], x_hi = per_bundle ? v : s.cone_par[1];
            const double y_lo = per_bundle ? -v : s.cone_par[2], y_hi = per_bundle ? v : s.cone_par[3];
            double x = x_lo + (x_hi - x_lo) * a;
            double y = y_lo + (y_hi - y_lo) * b;
            a0 = atan(sqrt(x * x + y * y));
            sincos(atan2(y, x), &s1, &c1);
        }
        double s0, c0;
        sincos(a0, &s0, &c0);
        l = v3(c1 * s0, s1 * s0, c0);
    }

    // ---- cone axis and basis (:262-293): rows (o_2, o_1, axis)
    V3 ax, o1, o2;
    if (((KN & KN_POINT_SOURCE) != 0) || s.kind == XRT_SRC_FIXED_AXIS) {
        // constant for every ray: precomputed on the host with the reference's formula
        o2 = v3(s.axis_basis[0], s.axis_basis[1], s.axis_basis[2]);
        o1 = v3(s.axis_basis[3], s.axis_basis[4], s.axis_basis[5]);
        ax = v3(s.axis_basis[6], s.axis_basis[7], s.axis_basis[8]);
    } else {
        ax = unit(v3(s.target) - r.o);
        V3 xa = v3(R[0], R[1], R[2]), za = v3(R[6], R[7], R[8]);
        o1 = unit(cross(ax, xa) + cross(ax, za));
        o2 = unit(cross(ax, o1));
    }
    r.d = v3(l.x * o2.x + l.y * o1.x + l.z * ax.x,
             l.x * o2.y + l.y * o1.y + l.z * ax.y,
             l.x * o2.z + l.y * o1.z + l.z * ax.z);

    // ---- source-level sightline filters (_XicsrtBundleFilterSightline.py:31-56)
    r.alive = !xy_exhausted;        // the reference loops until it has N accepted rays; a ray that cannot be drawn is dropped
    if constexpr ((FT & FT_SRC_EXT) != 0) {
        for (int f = 0; f < s.n_sightlines; ++f) {
            const XrtSightline &sl = s.sightlines[f];
            V3 l0 = v3(sl.origin) - r.o;
            V3 axs = v3(sl.axis);
            V3 perp = l0 - axs * dot(axs, l0);
            r.alive = r.alive && (sl.radius >= sqrt(dot(perp, perp)));
        }
    }
}

// wavelength (:295-367); `dir` is the ray direction at the source (Doppler shift)
template <class DR, uint32_t KN = 0, bool DOPPLER = true>
__device__ __forceinline__ double generate_wavelength(const XrtSourceDesc &s, const SrcLocal &L, const DR &dr, V3 dir) {
    double w;
    if (((KN & KN_WAVE_NORMAL) != 0) || s.wave == XRT_WAVE_NORMAL) {
        w = s.wave_par[0] + L.wave_sigma * dr.wave_z();
    } else if (s.wave == XRT_WAVE_CONST) {
        w = s.wave_par[0];
    } else if (s.wave == XRT_WAVE_UNIFORM) {
        w = s.wave_par[0] + (s.wave_par[1] - s.wave_par[0]) * dr.wave_u();
    } else if (s.kind == XRT_SRC_BUNDLES) {
        // the bundle's own table (every per-bundle source of the reference builds one)
        const double *cdf = s.bundle_cdf + L.bundle * (uint64_t)s.n_table;
        const double *x = s.bundle_x + L.bundle * (uint64_t)s.n_table;
        const double lo = __ldg(cdf), hi = __ldg(cdf + s.n_table - 1);
        double y = lo + (hi - lo) * dr.wave_u();
        w = interp_inside(y, cdf, x, s.n_table) + s.wave_par[0];
    } else {
        double y = s.wave_par[1] + (s.wave_par[2] - s.wave_par[1]) * dr.wave_u();
        w = interp_inside(y, s.table_cdf, s.table_x, s.n_table) + s.wave_par[0];
    }
    if constexpr (DOPPLER) {
        if (L.vel.x != 0.0 || L.vel.y != 0.0 || L.vel.z != 0.0) w *= 1.0 - dot(L.vel, dir);
    }
    return w;
}

// true when the wavelength does not depend on the direction at the source, i.e. it can be
// drawn later from (seed, ray id) alone
__device__ __forceinline__ bool wavelength_is_lazy(const XrtSourceDesc &s) {
    return s.kind != XRT_SRC_BUNDLES && s.velocity_c[0] == 0.0 && s.velocity_c[1] == 0.0 && s.velocity_c[2] == 0.0;
}

template <uint32_t FT, class DR, uint32_t KN = 0, bool USE_TABLE = false>
__device__ __forceinline__ void generate_ray(const XrtSourceDesc &s, const DR &dr, uint64_t index, Ray &r,
                                             const double *sincos_table = nullptr) {
    SrcLocal L;
    source_local<FT, KN>(s, index, L);
    generate_geometry<FT, DR, KN, USE_TABLE>(s, L, dr, r, sincos_table);
    r.w = generate_wavelength<DR, KN>(s, L, dr, r.d);
}

// ---------------------------------------------------------------------------
// analytic shapes: distance along the ray (NaN / false = no intersection)

// _ShapePlane.py:32-53
__device__ __forceinline__ bool hit_plane(const XrtOpticDesc &op, bool local, V3 o, V3 d, double &t) {
    if (local) {
        t = (0.0 - o.z) / d.z;
    } else {
        V3 z = v3(op.orient + 6);
        t = dot(v3(op.origin) - o, z) / dot(d, z);
    }
    return t >= 0.0;
}

// _ShapeSphere.py:53-100 -- geometric solution; concave takes the larger root, no sign test.
// The reference forms d = sqrt(L.L - tca^2), tests d <= R and then uses d^2 again; here the
// square is kept (d^2 < 0 is the reference's NaN -> miss), which saves one square root.
__device__ __forceinline__ bool hit_sphere(const XrtOpticDesc &op, bool convex, V3 o, V3 d, double &t) {
    V3 L = v3(op.center) - o;
    double tca = dot(L, d);
    double d2 = fma(-tca, tca, dot(L, L));
    double r2 = op.radius * op.radius;
    if (!(d2 >= 0.0 && d2 <= r2)) return false;
    double thc = fast_sqrt(r2 - d2);
    t = convex ? tca - thc : tca + thc;     // min / max of tca -+ thc (thc >= 0)
    return true;
}

// _ShapeCylinder.py:52-109 -- axis = element x axis through `center`
__device__ __forceinline__ bool hit_cylinder(const XrtOpticDesc &op, V3 o, V3 d, double &t) {
    V3 va = v3(op.orient);
    V3 dp = o - v3(op.center);
    V3 A1 = d - va * dot(d, va);
    V3 B1 = dp - va * dot(dp, va);
    double A = dot(A1, A1);
    double B = 2.0 * dot(A1, B1);
    double C = dot(B1, B1) - op.radius * op.radius;
    double dis = B * B - 4.0 * A * C;
    if (!(dis >= 0.0)) return false;
    double sq = sqrt(dis);
    double t0 = (-B - sq) / (2.0 * A), t1 = (-B + sq) / (2.0 * A);
    if (op.flags & XRT_F_CONVEX) t = (t0 < t1) ? t0 : t1;
    else t = (t0 > t1) ? t0 : t1;
    return true;
}

__device__ __forceinline__ double signed_cbrt(double x) { return cbrt(x); }

// One real root of z^3 + p z^2 + r z + c (xicsrt_quartic.py:54-162, all_roots=False)
__device__ __forceinline__ double resolvent_root(double p, double r, double c) {
    const double third = 1.0 / 3.0;
    double a13 = p * third;
    double a2 = a13 * a13;
    double f = third * r - a2;
    double g = a13 * (2.0 * a2 - r) + c;
    double h = 0.25 * g * g + f * f * f;
    if (f == 0.0 && g == 0.0 && h == 0.0) return -signed_cbrt(c);
    if (h <= 0.0) {
        double j = sqrt(-f);
        double k = acos(-0.5 * g / (j * j * j));
        return 2.0 * j * cos(third * k) - a13;
    }
    double sh = sqrt(h);
    return (signed_cbrt(-0.5 * g + sh) + signed_cbrt(-0.5 * g - sh)) - a13;
}

// _ShapeTorus.py:116-183 with xicsrt_quartic.py:165-207 in real arithmetic.
// The reference keeps complex values and declares a slot "no hit" when its
// imaginary part is non-zero; in real terms: s^2 = 2p + 2 z0 must be >= 0 (else
// every slot is complex) and the slot's quadratic discriminant must be >= 0.
// Slots 0,1 come from x^2 + s x + (z0 + t), slots 2,3 from x^2 - s x + (z0 - t),
// each ordered (-sqrt, +sqrt); the optic takes slot `root_idx`.
__device__ __forceinline__ bool hit_torus(const XrtOpticDesc &op, V3 o, V3 d, double &t) {
    const double rmaj = op.torus_major, rmin = op.torus_minor;
    V3 O = to_local(op.orient, o - v3(op.center));
    V3 D = to_local(op.orient, d);
    double OO = dot(O, O), OD = dot(O, D);
    double rsq = rmaj * rmaj + rmin * rmin;
    double rm2 = rmaj * rmaj;
    // monic quartic t^4 + a t^3 + b t^2 + c t + e  (axis of the torus = local y)
    double a = 4.0 * OD;
    double b = 4.0 * OD * OD + 2.0 * OO - 2.0 * rsq + 4.0 * rm2 * D.y * D.y;
    double c = 4.0 * OD * (OO - rsq) + 8.0 * rm2 * D.y * O.y;
    double dm = rm2 - rmin * rmin;
    double e = OO * OO - 2.0 * rsq * OO + 4.0 * rm2 * O.y * O.y + dm * dm;

    double q4 = 0.25 * a;
    double q42 = q4 * q4;
    double p = 3.0 * q42 - 0.5 * b;
    double q = a * q42 - b * q4 + 0.5 * c;
    double r = 3.0 * q42 * q42 - b * q42 + c * q4 - e;
    double z0 = resolvent_root(p, r, p * r - 0.5 * q * q);

    double s2 = 2.0 * p + 2.0 * z0;
    if (!(s2 >= 0.0)) return false;      // s imaginary (or NaN): all four slots complex
    double s = sqrt(s2);
    double tt = (s == 0.0) ? (z0 * z0 + r) : (-q / s);

    const int idx = op.root_idx;
    double half, cst;
    if (idx < 2) { half = -0.5 * s; cst = z0 + tt; }
    else { half = 0.5 * s; cst = z0 - tt; }
    double disc = half * half - cst;
    if (!(disc >= 0.0)) return false;
    double sq = sqrt(disc);
    double root = ((idx & 1) ? (half + sq) : (half - sq)) - q4;
    t = root;
    return isfinite(root) && root > 0.0;
}

// ---------------------------------------------------------------------------
// apertures (xicsrt_aperture.py:13-204): sequential fold over the list

__device__ __forceinline__ bool aperture_inside(const XrtAperture &ap, double x, double y) {
    double dx = x - ap.origin[0], dy = y - ap.origin[1];
    switch (ap.shape) {
    case XRT_AP_CIRCLE: return (dx * dx + dy * dy) < ap.size[0] * ap.size[0];
    case XRT_AP_SQUARE: return (fabs(dx) < ap.size[0] / 2) && (fabs(dy) < ap.size[0] / 2);
    case XRT_AP_RECTANGLE: return (fabs(dx) < ap.size[0] / 2) && (fabs(dy) < ap.size[1] / 2);
    case XRT_AP_ELLIPSE: {
        double ex = dx / ap.size[0], ey = dy / ap.size[1];   // full sizes as semi-axes (sic, :182)
        return (ex * ex + ey * ey) < 1.0;
    }
    case XRT_AP_TRIANGLE: {
        const double *v = ap.vert;   // p0 = (v0,v1) p1 = (v2,v3) p2 = (v4,v5)
        double area = 0.5 * (-v[3] * v[4] + v[1] * (-v[2] + v[4]) + v[0] * (v[3] - v[5]) + v[2] * v[5]);
        double k = 1.0 / (2.0 * area);
        double a = k * (v[1] * v[4] - v[0] * v[5] + (v[5] - v[1]) * x + (v[0] - v[4]) * y);
        double b = k * (v[0] * v[3] - v[1] * v[2] + (v[1] - v[3]) * x + (v[2] - v[0]) * y);
        double c = 1.0 - a - b;
        return (a >= 0.0) && (b >= 0.0) && (c >= 0.0);
    }
    default: return true;
    }
}

__device__ __forceinline__ bool aperture_fold(const XrtOpticDesc &op, double x, double y) {
    bool acc = true;
    for (int i = 0; i < op.n_aperture; ++i) {
        const XrtAperture ap = op.apertures[i];
        bool test = aperture_inside(ap, x, y);
        switch (ap.logic) {
        case XRT_LOGIC_AND: acc = acc && test; break;
        case XRT_LOGIC_NOT: acc = acc && !test; break;
        case XRT_LOGIC_OR: acc = acc || test; break;
        case XRT_LOGIC_NAND: acc = !(acc && test); break;
        case XRT_LOGIC_NOR: acc = !(acc || test); break;
        case XRT_LOGIC_XOR: acc = acc != test; break;
        default: acc = !(acc != test); break;   // XNOR
        }
    }
    return acc;
}

// ---------------------------------------------------------------------------
// Bragg reflection test (_InteractCrystal.py:96-196)
//
// The reference forms theta_B = asin(lambda / 2d) and theta_i = pi/2 - acos(|D.n| / |D|) and
// feeds theta_i - theta_B to the rocking curve.  With s = lambda / 2d and c = |D.n| / |D|
// (both in [0, 1]) that difference is  asin(c) - asin(s) = asin(c sqrt(1 - s^2) - s sqrt(1 - c^2)),
// an argument of order 1e-5..1e-3 wherever the rocking curve is not zero, for which asin is a
// short series.  Two square roots replace an asin and an acos; 1 - x^2 is formed with one
// fma, i.e. correctly rounded, so the conditioning is that of the reference's own acos.
__device__ __forceinline__ double bragg_dtheta(const XrtOpticDesc &op, V3 d, double w, V3 n) {
    double s = w * op.inv_two_d;
    double c = fabs(dot(d, n)) * fast_rsqrt(dot(d, d));
    double x = c * fast_sqrt(fma(-s, s, 1.0)) - s * fast_sqrt(fma(-c, c, 1.0));
    return (fabs(x) < 0.05) ? asin_small(x) : asin(x);      // NaN (lambda > 2d) goes to asin -> NaN
}

// Conservative pre-test of the Bragg condition for the fused kernel (history-off path).
//
// In a spectrometer ~98 % of the rays that reach the crystal fail the rocking-curve test, most
// of them by many widths.  theta_B - theta_i is bounded from below without any inverse
// trigonometry: with sB = lambda / 2d = sin(theta_B) and sI = |D.n| = sin(theta_i), both
// angles in [0, pi/2],  |sB - sI| <= |theta_B - theta_i| cos(min(theta_B, theta_i)).  So
//     (sB - sI)^2 > T^2 (1 - min(sB, sI)^2)   ==>   |theta_B - theta_i| > T,
// and with T a little beyond the angle where the rocking curve is zero (step) or below 2^-57
// (Gaussian, x >= 40: bragg_pass returns false there) the ray is lost whatever its uniform is.
// The wavelength of the test is approximate -- the normal deviate of the exact path evaluated
// in FP32 from the top bits of the same uniform (|z32 - z| < 1e-5, normal_approx below) --
// and op.cull_err = 2e-3 sigma_lambda / 2d + 1e-9 covers that and the rounding of the exact
// path many times over; T carries a 5 % + 2e-6 rad margin.  Rays that pass the pre-test take
// the exact FP64 path; rays that fail it would have failed there too (checked ray for ray against
// the replay kernel in tests/test_gpu_statistics.py).  Sphere crystals, |D| = 1.  The pre-test runs
// before the bounds test: a ray it rejects is lost at this optic either way.
__device__ __forceinline__ float lg2_approx(float x) {
    float y;
    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float sqrt_approx(float x) {
    float y;
    asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float rsqrt_approx(float x) {
    float y;
    asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}


// Standard normal deviate of wave_z() from the top bits of its uniform, in FP32: z = sqrt(2) erfinv(2u - 1) with a
// central-branch polynomial of the form of Giles' single-precision erfinv (w = -ln(1 - x^2) < 5, i.e. |z| < 2.93),
// refitted at degree 4: the pre-test's margin (cull_err) allows |z32 - z| < 2e-3, the fit is within 3.4e-5 and the
// 23-bit argument adds 2^-23 / pdf(2.93) = 2.2e-5 (checked against scipy in tests/test_fastmath.py).  x = 2u - 1 is
// built from the top 23 bits as a float in [2, 4) minus 3: no integer-to-float conversion (XU pipe).
__device__ __forceinline__ float normal_approx(uint32_t hi, bool &usable) {
    const float x = __uint_as_float(0x40000000u | (hi >> 9)) - 3.0f;       // -1 + (hi >> 9) 2^-22, exact
    float w = -0.6931471805599453f * lg2_approx(fmaf(-x, x, 1.0f));
    usable = w < 5.0f;                                                      // false for x = -1 (w = inf) as well
    w -= 2.5f;
    float p = 0.000186555306f;
    p = fmaf(p, w, -0.00126819056f);
    p = fmaf(p, w, -0.00410411088f);
    p = fmaf(p, w, 0.246652201f);
    p = fmaf(p, w, 1.50138509f);
    return 1.4142135623730951f * p * x;
}

// cos^2(min(theta_B, theta_i)) <= (1 - sI^2) + 2 |sB - sI|  (equality to first order when sB < sI), which
// avoids the FP64 min / max selects.
__device__ __forceinline__ bool bragg_cull_test(const XrtOpticDesc &op, double sB, double sI, bool usable, double err) {
    const double gap = fabs(sB - sI);
    const double diff = gap - err;
    const double c2 = fma(2.0, gap, fma(-sI, sI, 1.0));
    return usable & (diff > 0.0) & (diff * diff > op.cull_t2 * c2);
}

// sphere hit from a point source: sI = |D.n| = thc / R, the half chord of hit_sphere over the radius (|D| = 1).
// Also hands back the two numbers of the bound, gap = |sB - sI| (-1 when the approximate deviate is not usable) and
// c2 >= cos^2(min angle), for the second, uniform-dependent test of stage B1 (bragg_cull_uniform).
__device__ __forceinline__ bool bragg_cull_sphere(const XrtSourceDesc &s, const XrtOpticDesc &op, uint32_t wave_hi, double thc,
                                                  double &gap_out, double &c2_out) {
    bool usable;
    const float z = normal_approx(wave_hi, usable);
    const double sB = fma((double)z, s.wave_par[1], s.wave_par[0]) * op.inv_two_d;
    const double sI = thc * op.cull_inv_r;
    const double gap = fabs(sB - sI);
    const double diff = gap - op.cull_err;
    const double c2 = fma(2.0, gap, fma(-sI, sI, 1.0));
    gap_out = usable ? gap : -1.0;
    c2_out = c2;
    return usable & (diff > 0.0) & (diff * diff > op.cull_t2 * c2);
}

// Second level of the pre-test, for a Gaussian rocking curve, once the ray's rocking-curve uniform u is known:
// the ray is reflected iff exp(-x) reflectivity >= u, x = dtheta^2 / 2 sigma^2, i.e. iff x <= ln(reflectivity / u).
// With the lower bound on |dtheta| of the first level, x >= (gap - err)^2 / (c2 two_sigma2); if that exceeds
// ln(reflectivity / u) (taken 1e-3 + 0.1 % too large: FP32 logarithm of a rounded u) the ray is lost.
// u = 0 gives an infinite bound: never rejected here.
__device__ __forceinline__ bool bragg_cull_uniform(const XrtOpticDesc &op, double gap, double c2, double err, double u) {
    const float lim = 0.6931471805599453f * (lg2_approx((float)op.reflectivity) - lg2_approx((float)u));   // ln(refl / u)
    const double bound = (double)fmaf(fabsf(lim), 1e-3f, lim + 1e-3f) * op.rock_two_sigma2;   // on dtheta^2
    const double diff = gap - err;
    return (diff > 0.0) & (diff * diff > bound * c2) & (lim == lim);
}

// general form for a sphere traced in global coordinates: X = intersection point, d = unit direction.
//   WAVE_EXACT    `lambda` is the ray's wavelength, already drawn
//   WAVE_APPROX   constant or normal line of the source, deviate from normal_approx (lazy wavelength)
//   WAVE_DEFERRED plasma bundle / Doppler-shifted normal line whose exact deviate is left to stage B:
//                 `lambda` holds the Doppler factor 1 - v.D / c and `sigma` the line's sigma for this ray
enum { WAVE_EXACT = 0, WAVE_APPROX = 1, WAVE_DEFERRED = 2 };
__device__ __forceinline__ bool bragg_cull_general(const XrtSourceDesc &s, const XrtOpticDesc &op, int mode,
                                                   double lambda, double sigma, uint32_t wave_hi, V3 X, V3 d,
                                                   double &gap_out, double &c2_out, double &err_out) {
    bool usable = true;
    double err = op.cull_err;
    if (mode == WAVE_APPROX) {
        lambda = s.wave_par[0];
        if (s.wave == XRT_WAVE_NORMAL) lambda = fma((double)normal_approx(wave_hi, usable), s.wave_par[1], lambda);
    } else if (mode == WAVE_DEFERRED) {
        lambda *= fma((double)normal_approx(wave_hi, usable), sigma, s.wave_par[0]);
        err = fma(2e-3 * fabs(sigma), fabs(op.inv_two_d), err);       // |z32 - z| < 2e-3, Doppler factor <= 2
        err += err;
    }
    const double sI = fabs(dot(d, v3(op.center) - X)) * op.cull_inv_r;
    const double sB = lambda * op.inv_two_d;
    const double gap = fabs(sB - sI);
    const double diff = gap - err;
    const double c2 = fma(2.0, gap, fma(-sI, sI, 1.0));
    gap_out = usable ? gap : -1.0;          // -1: no second level either (the approximate deviate is out of range)
    c2_out = c2;
    err_out = err;
    return usable & (diff > 0.0) & (diff * diff > op.cull_t2 * c2);
}

// true = reflected.  p = rocking(dtheta) * reflectivity, keep when p >= u (:186-196).
template <uint32_t FT, class DR, uint32_t KN = 0, bool MOSAIC = false>
__device__ __forceinline__ bool bragg_pass(const XrtOpticDesc &op, int k, int layer, const DR &dr, double dth) {
    double p;
    const int rocking = rocking_of<KN>(op);
    if (rocking == XRT_ROCK_GAUSS) {
        // sigma = fwhm / (2 sqrt(2 ln 2)); p = exp(-dth^2 / (2 sigma^2))
        double x = (dth * dth) * op.rock_inv_two_sigma2;
        if (x >= 40.0) return dr.bragg_u_is_zero(k, layer);   // p < 2^-57: only u == 0 passes
        if (!(x >= 0.0)) return false;                        // NaN
        p = exp_neg(x);
    } else if (rocking == XRT_ROCK_STEP) {
        p = (fabs(dth) <= op.rocking_fwhm / 2.0) ? 1.0 : 0.0;
    } else {
        p = 0.0;
        if constexpr ((FT & FT_ROCKTAB) != 0) {
            int n = op.n_rock;
            if (dth >= __ldg(op.rock_dtheta) && dth <= __ldg(op.rock_dtheta + n - 1)) {
                double sg = interp_inside(dth, op.rock_dtheta, op.rock_s, n);
                double pi = interp_inside(dth, op.rock_dtheta, op.rock_p, n);
                p = op.rocking_mix * sg + (1.0 - op.rocking_mix) * pi;
            }
        }
        if (!(dth == dth)) return false;
    }
    p *= op.reflectivity;
    // p == 0 (outside a step curve or a table): only a uniform that is exactly 0 passes `p >= u`; as for the
    // Gaussian tail above, Philox draws never do (the Bragg pre-test relies on it), injected draws are compared
    if (!MOSAIC && p == 0.0) return dr.bragg_u_is_zero(k, layer);
    if constexpr (MOSAIC) return p >= dr.mosaic_u(k, layer);
    else return p >= dr.bragg_u(k, layer);
}

// _InteractMirror.py:29-42
__device__ __forceinline__ void reflect(Ray &r, V3 n) {
    double k = 2.0 * dot(r.d, n);
    r.d = v3(r.d.x - k * n.x, r.d.y - k * n.y, r.d.z - k * n.z);
}

// ---------------------------------------------------------------------------
// surface normals at the intersection point X

// _ShapeSphere.py:102-106
__device__ __forceinline__ V3 normal_sphere(const XrtOpticDesc &op, V3 X) {
    return unit(v3(op.center) - X);
}

// _ShapeCylinder.py:111-133 -- toward the point of the axis at the same x
__device__ __forceinline__ V3 normal_cylinder(const XrtOpticDesc &op, V3 X) {
    V3 pa = v3(op.center), va = v3(op.orient);
    double s = dot(pa - X, va);
    return unit((pa - va * s) - X);
}

// _ShapeTorus.py:186-216 -- away from the nearest point of the axis circle
__device__ __forceinline__ V3 normal_torus(const XrtOpticDesc &op, V3 X) {
    V3 C = v3(op.center), ya = v3(op.orient + 3);
    V3 p = X - C;
    p = p - ya * dot(p, ya);
    V3 Q = C + p * (op.torus_major * fast_rsqrt(dot(p, p)));
    return unit(X - Q);
}

// ---------------------------------------------------------------------------
// mosaic crystallite normal about the nominal normal n
// (_InteractMosaicCrystal.py:109-139, xicsrt_spread.py:297-339)

__device__ __forceinline__ V3 mosaic_normal(V3 n, double x, double y) {
    double inv = fast_rsqrt(x * x + y * y + 1.0);
    double lx = x * inv, ly = y * inv, lz = inv;
    // R0 = n x [1,0,0] + n x [0,0,1];  R1 = n x R0
    V3 r0 = v3(0.0 + n.y, n.z - n.x, -n.y + 0.0);
    r0 = unit(r0);
    V3 r1 = unit(cross(n, r0));
    return v3(lx * r0.x + ly * r1.x + lz * n.x,
              lx * r0.y + ly * r1.y + lz * n.y,
              lx * r0.z + ly * r1.z + lz * n.z);
}

// ---------------------------------------------------------------------------
// pixel binning (_TraceObject.py:234-293): channel = rint(local / pixel + (npix-1)/2)

__device__ __forceinline__ bool pixel_index(const XrtOpticDesc &op, V3 o, uint32_t &idx) {
    V3 pl = to_local(op.orient, o - v3(op.origin));
    double cx = rint(pl.x / op.pixel_size + 0.5 * (double)(op.npix[0] - 1));
    double cy = rint(pl.y / op.pixel_size + 0.5 * (double)(op.npix[1] - 1));
    if (!(cx >= 0.0 && cx < (double)op.npix[0] && cy >= 0.0 && cy < (double)op.npix[1])) return false;
    idx = (uint32_t)cx * (uint32_t)op.npix[1] + (uint32_t)cy;
    return true;
}

}  // namespace xrt

#include "xrt_mesh.cuh"

namespace xrt {

// A ray that is already lost still passes through trace_global of the following optics:
// its origin is overwritten by the (NaN) intersection point and, for optics traced in
// local coordinates, its direction is transformed there and back.
template <uint32_t FT>
__device__ __forceinline__ void pass_lost_ray(const XrtOpticDesc &op, Ray &r) {
    r.o = nan3();
    if constexpr ((FT & FT_LOCAL) != 0) {
        if (op.flags & XRT_F_TRACE_LOCAL) r.d = to_external(op.orient, to_local(op.orient, r.d));
    }
}

// ---------------------------------------------------------------------------
// one optic (_TraceObject.py:135-178), in two halves so that the fused kernel can
// re-pack the surviving rays of a warp between them.
//
// geometry half: to-local, intersect, bounds / aperture.  On return r.o = intersection
// point and r.d = direction, both in the frame the optic traces in (local when
// trace_local), n = surface normal there.  Return value:
//   HIT_MISS    surface missed: r.o = NaN, r.d back in external coordinates, r.alive = false
//   HIT_OUTSIDE hit outside bounds / aperture: r.o, r.d in external coordinates, r.alive = false
//   HIT_INSIDE  candidate for the interaction half (frame not yet converted back)
enum { HIT_MISS = 0, HIT_OUTSIDE = 1, HIT_INSIDE = 2 };

template <uint32_t FT>
__device__ __forceinline__ bool optic_is_local(const XrtOpticDesc &op) {
    if constexpr ((FT & FT_LOCAL) != 0) return (op.flags & XRT_F_TRACE_LOCAL) != 0;
    return false;
}

// surface normal of an analytic shape at X (tracing frame)
template <uint32_t FT, uint32_t KN = 0>
__device__ __forceinline__ V3 analytic_normal(const XrtOpticDesc &op, V3 X) {
    const int shape = shape_of<KN>(op);
    // the plane normal is the element's (global) zaxis even when tracing in local
    // coordinates -- _ShapePlane.py:55-62 does not transform it; replicated.
    if (shape == XRT_SHAPE_PLANE) return v3(op.orient + 6);
    if (shape == XRT_SHAPE_SPHERE) return normal_sphere(op, X);
    V3 n = v3(0.0, 0.0, 1.0);
    if constexpr ((FT & FT_CYL) != 0) { if (shape == XRT_SHAPE_CYLINDER) n = normal_cylinder(op, X); }
    if constexpr ((FT & FT_TORUS) != 0) { if (shape == XRT_SHAPE_TORUS) n = normal_torus(op, X); }
    return n;
}

template <uint32_t FT>
__device__ __forceinline__ void frame_to_external(const XrtOpticDesc &op, bool local, Ray &r) {
    if constexpr ((FT & FT_LOCAL) != 0) {
        if (local) {   // _GeometryObject.py:113-125
            r.o = to_external(op.orient, r.o) + v3(op.origin);
            r.d = to_external(op.orient, r.d);
        }
    }
}

template <uint32_t FT, bool WANT_NORMAL, uint32_t KN = 0, bool INLINE_MESH = false>
__device__ __forceinline__ int optic_geometry(const XrtOpticDesc &op, Ray &r, V3 &n, const double *staged_mesh = nullptr,
                                              const V3 *mesh_resume = nullptr) {
    V3 o = r.o, d = r.d;
    const uint32_t flags = flags_of<KN>(op);
    const bool local = optic_is_local<FT>(op);
    if constexpr ((FT & FT_LOCAL) != 0) {
        if (local) {   // _GeometryObject.py:127-141
            o = to_local(op.orient, o - v3(op.origin));
            d = to_local(op.orient, d);
        }
    }

    // ---- intersect: distance, location (and normal for mesh shapes)
    V3 X;
    bool ok;
    bool analytic = true;
    if constexpr ((FT & FT_MESH) != 0) {
        if (shape_of<KN>(op) == XRT_SHAPE_MESH) {
            analytic = false;
            if constexpr (INLINE_MESH) ok = mesh_intersect_inline(op, o, d, X, n, staged_mesh, mesh_resume);
            else ok = mesh_intersect(op, o, d, X, n, staged_mesh, mesh_resume);
        }
    }
    if (analytic) {
        double t = 0.0;
        const int shape = shape_of<KN>(op);
        if (shape == XRT_SHAPE_PLANE) ok = hit_plane(op, local, o, d, t);
        else if (shape == XRT_SHAPE_SPHERE) ok = hit_sphere(op, (flags & XRT_F_CONVEX) != 0, o, d, t);
        else {
            ok = false;
            if constexpr ((FT & FT_CYL) != 0) { if (shape == XRT_SHAPE_CYLINDER) ok = hit_cylinder(op, o, d, t); }
            if constexpr ((FT & FT_TORUS) != 0) { if (shape == XRT_SHAPE_TORUS) ok = hit_torus(op, o, d, t); }
        }
        X = v3(fma(d.x, t, o.x), fma(d.y, t, o.y), fma(d.z, t, o.z));   // _ShapeObject.py:69-81
    }
    r.alive = false;
    if (!ok) {
        r.o = nan3();
        // ray_to_external acts on every ray, hit or not (_TraceObject.py:135-155): the
        // direction makes the local round trip (visible when the axes are not exactly unit)
        if constexpr ((FT & FT_LOCAL) != 0) { if (local) r.d = to_external(op.orient, d); }
        return HIT_MISS;
    }

    // ---- bounds (_TraceObject.py:180-232): strict |x| < size/2, then apertures
    V3 Xl = local ? X : to_local(op.orient, X - v3(op.origin));
    bool in = true;
    if (flags & XRT_F_CHECK_SIZE) {
        if (flags & XRT_F_HAS_XSIZE) in = in && (fabs(Xl.x) < op.half_size[0]);
        if (flags & XRT_F_HAS_YSIZE) in = in && (fabs(Xl.y) < op.half_size[1]);
        if (flags & XRT_F_HAS_ZSIZE) in = in && (fabs(Xl.z) < op.half_size[2]);
    }
    if constexpr ((FT & FT_APERTURE) != 0) {
        if (in && (flags & XRT_F_CHECK_APERTURE) && op.n_aperture > 0) in = aperture_fold(op, Xl.x, Xl.y);
    }
    r.o = X;
    r.d = d;
    if (!in) {
        frame_to_external<FT>(op, local, r);
        return HIT_OUTSIDE;
    }
    if constexpr (WANT_NORMAL) { if (analytic) n = analytic_normal<FT, KN>(op, X); }
    r.alive = true;
    return HIT_INSIDE;
}

// interaction half (None / Mirror / Crystal / Mosaic) and the way back to external
// coordinates.  r.o, r.d in the tracing frame, r.w valid, n = normal at r.o.
template <uint32_t FT, class DR, uint32_t KN = 0>
__device__ __forceinline__ void optic_interact(const XrtOpticDesc &op, int k, const DR &dr, Ray &r, V3 n) {
    bool alive = true;
    const int ia = interact_of<KN>(op);
    const uint32_t flags = flags_of<KN>(op);
    if (ia == XRT_INTERACT_MIRROR) {
        reflect(r, n);
    } else if (ia == XRT_INTERACT_CRYSTAL) {
        if (flags & XRT_F_CHECK_BRAGG) alive = bragg_pass<FT, DR, KN>(op, k, 0, dr, bragg_dtheta(op, r.d, r.w, n));
        if (alive) reflect(r, n);
    } else if (ia == XRT_INTERACT_MOSAIC) {
        if constexpr ((FT & FT_MOSAIC) != 0) {
            // optional prefilter on the nominal normal (_InteractMosaicCrystal.py:67-75)
            if (flags & XRT_F_MOSAIC_CUTOFF) alive = fabs(bragg_dtheta(op, r.d, r.w, n)) < op.mosaic_angle_cut;
            if (alive) {
                // layers of crystallites: the first one that satisfies Bragg reflects (:83-104)
                bool done = false;
                for (int layer = 0; layer < op.mosaic_depth && !done; ++layer) {
                    double x, y;
                    dr.mosaic_xy(k, layer, op.mosaic_sin_sigma, x, y);
                    V3 nm = mosaic_normal(n, x, y);
                    bool pass = true;
                    if (flags & XRT_F_CHECK_BRAGG) pass = bragg_pass<FT, DR, KN, true>(op, k, layer, dr, bragg_dtheta(op, r.d, r.w, nm));
                    if (pass) { reflect(r, nm); done = true; }
                }
                alive = done;
            }
        } else {
            alive = false;
        }
    }
    frame_to_external<FT>(op, optic_is_local<FT>(op), r);
    r.alive = alive;
}

// both halves.  `r` must be alive on entry.  On exit: r.alive = survived; r.o = intersection
// point (NaN when the surface was missed), r.d reflected only for surviving rays -- exactly
// the state the reference's history holds for this element.
template <uint32_t FT, class DR, uint32_t KN = 0>
__device__ __forceinline__ void trace_optic(const XrtOpticDesc &op, int k, const DR &dr, Ray &r) {
    V3 n;
    if (optic_geometry<FT, true, KN>(op, r, n) == HIT_INSIDE) optic_interact<FT, DR, KN>(op, k, dr, r, n);
}

// One out-of-line copy for the variants whose code is large (mesh optics): the fused kernel reaches trace_optic from
// the optics before the split optic and from stage C, and two inlined copies of every shape and interaction cost more
// in instruction fetch (the mesh variants are 180 kB of code) than the call does.
template <uint32_t FT, class DR>
static __device__ __noinline__ void trace_optic_shared(const XrtOpticDesc &op, int k, const DR &dr, Ray &r) {
    trace_optic<FT, DR, 0>(op, k, dr, r);
}

template <uint32_t FT, class DR>
__device__ __forceinline__ void trace_optic_any(const XrtOpticDesc &op, int k, const DR &dr, Ray &r) {
    if constexpr ((FT & FT_MESH) != 0) trace_optic_shared<FT, DR>(op, k, dr, r);
    else trace_optic<FT, DR, 0>(op, k, dr, r);
}

}  // namespace xrt
