// xrt.cu -- kernels and the C ABI of libxrt.so (see include/xrt.h).
//
// Kernels
//   k_trace<FT>     fused generate -> optic train -> bin, one ray per thread, ray
//                   state in registers from the source to the detector; the only
//                   global traffic is the per-element survivor counters (one
//                   atomic per block), the pixel counters of surviving rays
//                   (warp-aggregated atomics) and the optional found / lost id
//                   lists (ballot + prefix-sum compaction).
//   k_record<FT,..> same ray code, but every element's ray state is stored as
//                   struct-of-arrays history (coalesced 8-byte planes); rays come
//                   either from Philox by id (history of selected rays) or from
//                   caller memory with injected draws (parity entry).
//   k_source<..>    source only (history element 0).
//   k_burn          dependent DFMA chains: the FP64 roofline denominator.
#include <cuda_runtime.h>

#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <vector>

#include "xrt_trace.cuh"
#include "xrt_plasma.cuh"

namespace xrt {

#ifndef XRT_BLOCK
#define XRT_BLOCK 256
#endif
#ifndef XRT_MIN_BLOCKS
#define XRT_MIN_BLOCKS 3
#endif
#ifndef XRT_RECORD_BLOCKS
#define XRT_RECORD_BLOCKS 3
#endif
constexpr int kBlock = XRT_BLOCK;
constexpr unsigned kFull = 0xffffffffu;

// ---------------------------------------------------------------------------
// fused kernel
//
// Each warp works through its share of the ray ids in three stages and re-packs the
// survivors between them in per-warp shared-memory queues (ballot + popc prefix sums), so
// that every stage runs with (nearly) all 32 lanes busy although ~48 % of the rays of a
// typical spectrometer miss the crystal and ~98 % of the rest fail the Bragg test:
//
//   stage A  all rays     source origin + direction, optics before the split optic, geometry
//                         (intersect + bounds) of the split optic           -> queue 1
//   stage B  queue 1      wavelength (drawn here when it does not depend on the source
//                         direction: Philox is counter based), interaction of the split
//                         optic (Bragg / mosaic / mirror), its image          -> queue 2
//   stage C  queue 2      the remaining optics, images, found list
//
// Variants of this scheme:
//   spectrometer (KN)  A32 FP32 broad phase of the Bragg pre-test, all rays        -> queue 0 (ids)
//                      A64+B1 FP64 direction, sphere chord, first level of the pre-test, intersection point,
//                          bounds, second level (rocking uniform)                     -> queue b
//                      (broad phase off: FP64 stage A for every ray -> queue 1 -> B1 -> queue b)
//                      B2  exact wavelength, Bragg angle, rocking curve, reflection   -> queue 2, then C
//   mesh split optic   A1 coarse mesh for every ray -> queue a; A2 refinement + interpolation -> queue 1
//   other scenes       stage A ends with the (FP64) Bragg pre-test where it applies (bragg_cull_general)
//
// The split optic is the first crystal of the train (0 if there is none).  SPLIT >= 0 makes
// its index a compile-time constant, so its parameters are fetched from the constant bank at
// fixed offsets (uniform loads) instead of register-indexed ones.  Warps never wait for one
// another: only __syncwarp and warp-uniform queue counters.

constexpr int kQ1Cap = 64;     // stage A pushes <= 32 per pass, stage B pops 32 when >= 32 are queued
constexpr int kQaCap = 64;     // mesh variants: rays that hit the coarse mesh (id, coarse hit point), between the halves of stage A
constexpr int kQaPlanes = 4;
#ifndef XRT_UNROLL
#define XRT_UNROLL 2
#endif
constexpr int kUnroll = XRT_UNROLL;   // spectrometer variant: groups of 32 rays per stage-A pass (independent chains)
constexpr int kQ1CapSpectro = 32 * (kUnroll + 1);
constexpr int kQ1PlanesSpectro = 7;   // id, direction, distance, and the two numbers of the pre-test bound (gap, c2)
constexpr int kQbPlanes = 5;          // id, direction, distance
#ifndef XRT_UNROLL32
#define XRT_UNROLL32 3
#endif
constexpr int kUnroll32 = XRT_UNROLL32;      // groups of 32 rays per pass of the FP32 broad phase
constexpr int kQ0Cap = 32 * (kUnroll32 + 1);   // spectrometer variant: ids that passed the FP32 broad phase
constexpr int kQbCap = 64;            // spectrometer variant: rays inside the bounds, between the two halves of stage B
constexpr int kQ2Cap = 64;     // stage B pushes <= 32 per pass, stage C pops 32 when >= 32 are queued
constexpr int kQ2Planes = 8;   // id, origin, direction, wavelength

template <uint32_t FT> __host__ __device__ constexpr int q1_planes() {
    // id, intersection point, direction [, wavelength when it can be eager] [, normal for mesh shapes]
    return 7 + (FT != 0 ? 1 : 0) + ((FT & FT_MESH) != 0 ? 3 : 0);
}
template <uint32_t FT, uint32_t KN = 0> __host__ __device__ constexpr int q1_doubles() {
    return ((KN & KN_SPECTROMETER) == KN_SPECTROMETER) ? kQ1PlanesSpectro * kQ1CapSpectro + kQbPlanes * kQbCap + kQ0Cap
                                                        : q1_planes<FT>() * kQ1Cap + ((FT & FT_MESH) != 0 ? kQaPlanes * kQaCap : 0);
}
template <uint32_t FT, uint32_t KN = 0> __host__ __device__ constexpr int warp_queue_doubles() {
    return q1_doubles<FT, KN>() + kQ2Planes * kQ2Cap;
}

// shared-memory copy of the step-1 face operands of a mesh split optic (<= kStageFaces faces)
constexpr int kStageFaces = 128;
template <uint32_t FT, uint32_t KN = 0> __host__ __device__ constexpr size_t block_smem_bytes() {
    return ((size_t)(XRT_BLOCK / 32) * warp_queue_doubles<FT, KN>() + ((FT & FT_MESH) != 0 ? 9 * kStageFaces : 0)) * sizeof(double);
}

struct WarpCtx {
    unsigned lane, lt_mask;
    unsigned long long *s_cnt;
};

__device__ __forceinline__ void count_alive(const WarpCtx &c, int elem, bool alive) {
    unsigned m = __ballot_sync(kFull, alive);
    if (c.lane == 0 && m) atomicAdd(&c.s_cnt[elem], (unsigned long long)__popc(m));
}

// pixel hit: one atomic per distinct pixel among the calling lanes
__device__ __forceinline__ void add_pixel(const XrtOutputs &out, const XrtOpticDesc &op, const Ray &r, unsigned lt_mask) {
    uint32_t pix;
    if (pixel_index(op, r.o, pix)) {
        unsigned act = __activemask();
        unsigned same = __match_any_sync(act, pix);
        if ((same & lt_mask) == 0)
            atomicAdd((unsigned long long *)(out.images + op.image_offset + pix), (unsigned long long)__popc(same));
    }
}

// found list: ballot + prefix-sum compaction, one atomic per warp (all 32 lanes call this)
__device__ __forceinline__ void emit_found(const XrtOutputs &out, const WarpCtx &c, bool found, uint64_t id) {
    if (!out.found_count) return;
    unsigned m = __ballot_sync(kFull, found);
    if (!m) return;
    unsigned long long off = 0;
    if (c.lane == 0) off = atomicAdd((unsigned long long *)out.found_count, (unsigned long long)__popc(m));
    off = __shfl_sync(kFull, off, 0);
    if (found && out.found_ids) {
        unsigned long long slot = off + __popc(m & c.lt_mask);
        if (slot < out.found_capacity) out.found_ids[slot] = id;
    }
}

// lost sample: a lost ray is kept when its 64-bit Philox key is below the threshold
__device__ __forceinline__ void emit_lost(const XrtOutputs &out, const WarpCtx &c, const PhiloxDraws &dr, bool lost,
                                          uint64_t id) {
    if (!out.lost_count) return;
    bool keep = false;
    uint64_t key = 0;
    if (lost) {
        key = dr.lost_key();
        keep = key < out.lost_threshold;
    }
    unsigned m = __ballot_sync(kFull, keep);
    if (!m) return;
    unsigned long long off = 0;
    if (c.lane == 0) off = atomicAdd((unsigned long long *)out.lost_count, (unsigned long long)__popc(m));
    off = __shfl_sync(kFull, off, 0);
    if (keep && out.lost_ids) {
        unsigned long long slot = off + __popc(m & c.lt_mask);
        if (slot < out.lost_capacity) {
            out.lost_ids[slot] = id;
            if (out.lost_keys) out.lost_keys[slot] = key;
        }
    }
}

// ---- stage C: the optics after the split optic, for the `cnt` rays in queue 2
template <uint32_t FT>
__device__ __forceinline__ void stage_c(const XrtSceneDesc &sc, const XrtOutputs &out, const WarpCtx &c, int split,
                                        const PhiloxKeys &pk, uint64_t stream_id, const double *q2, int first, int cnt) {
    const bool active = (int)c.lane < cnt;
    Ray r;
    r.alive = false;
    uint64_t id = 0;
    PhiloxDraws dr;
    if (active) {
        const double *p = q2 + first + c.lane;
        id = (uint64_t)__double_as_longlong(p[0]);
        r.o = v3(p[1 * kQ2Cap], p[2 * kQ2Cap], p[3 * kQ2Cap]);
        r.d = v3(p[4 * kQ2Cap], p[5 * kQ2Cap], p[6 * kQ2Cap]);
        r.w = p[7 * kQ2Cap];
        r.alive = true;
    }
    __syncwarp();
    dr.init(pk, stream_id, id, split);
    const int nopt = sc.n_optics;
    for (int k = split + 1; k < nopt; ++k) {
        const XrtOpticDesc &op = sc.optics[k];
        if (r.alive) {
            trace_optic<FT>(op, k, dr, r);
            if (r.alive && (op.flags & XRT_F_IMAGE) && out.images) add_pixel(out, op, r, c.lt_mask);
        }
        count_alive(c, k + 1, r.alive);
    }
    emit_found(out, c, r.alive, id);
    emit_lost(out, c, dr, active && !r.alive, id);
}

// ---- stage B: interaction of the split optic for `cnt` rays popped from queue 1
template <uint32_t FT, uint32_t KN>
__device__ __forceinline__ void stage_b(const XrtSceneDesc &sc, const XrtOpticDesc &ops, const XrtOutputs &out,
                                        const WarpCtx &c, int split, bool lazy, bool need_wave, bool defer, const PhiloxKeys &pk, uint64_t stream_id,
                                        const double *q1, int first, int cnt, double *q2, int &n2, unsigned &n_split) {
    constexpr int P = ((KN & KN_SPECTROMETER) == KN_SPECTROMETER) ? kQbCap : kQ1Cap;   // spectrometer: q1 = queue b here
    const bool active = (int)c.lane < cnt;
    Ray r;
    r.alive = false;
    r.w = 0.0;
    V3 n = v3(0.0, 0.0, 1.0);
    uint64_t id = 0;
    constexpr bool SPECTRO = (KN & KN_SPECTROMETER) == KN_SPECTROMETER;
    bool inside = active;
    if (active) {
        const double *p = q1 + first + c.lane;
        id = (uint64_t)__double_as_longlong(p[0]);
        if constexpr (SPECTRO) {
            // queue b of the spectrometer variant: (id, direction, distance) of rays inside the bounds
            const V3 d = v3(p[1 * P], p[2 * P], p[3 * P]);
            const double t = p[4 * P];
            const V3 o = v3(sc.source.origin);
            const V3 X = v3(fma(d.x, t, o.x), fma(d.y, t, o.y), fma(d.z, t, o.z));
            r.o = X;
            r.d = d;
        } else {
            r.o = v3(p[1 * P], p[2 * P], p[3 * P]);
            r.d = v3(p[4 * P], p[5 * P], p[6 * P]);
            if constexpr (FT != 0) r.w = p[7 * P];
            if constexpr ((FT & FT_MESH) != 0) n = v3(p[8 * P], p[9 * P], p[10 * P]);
        }
    }
    __syncwarp();
    PhiloxDraws dr;
    dr.init(pk, stream_id, id, split);
    if (inside) {
        if (lazy && need_wave) {      // history is off here: a wavelength nobody tests is not drawn
            SrcLocal L;
            source_local<0, KN>(sc.source, id, L);
            r.w = generate_wavelength<PhiloxDraws, KN, false>(sc.source, L, dr, r.d);   // lazy = no Doppler shift
        }
        if constexpr (FT != 0) {
            if (defer) {              // r.w holds the Doppler factor (stage A); the same expressions as generate_wavelength
                SrcLocal L;
                source_local<FT, KN>(sc.source, id, L);
                const double w0 = sc.source.wave_par[0] + L.wave_sigma * dr.wave_z();
                r.w = (r.w != 1.0) ? w0 * r.w : w0;
            }
        }
        bool analytic = true;
        if constexpr ((FT & FT_MESH) != 0) analytic = ops.shape != XRT_SHAPE_MESH;
        if (analytic) n = analytic_normal<FT, KN>(ops, r.o);
        optic_interact<FT, PhiloxDraws, KN>(ops, split, dr, r, n);
        if (r.alive && (flags_of<KN>(ops) & XRT_F_IMAGE) && out.images) add_pixel(out, ops, r, c.lt_mask);
    }
    const unsigned m = __ballot_sync(kFull, r.alive);
    n_split += __popc(m);                  // survivors of the split optic: per-warp register counter
    emit_lost(out, c, dr, active && !r.alive, id);

    if (split + 1 >= sc.n_optics) {        // the split optic is the last one: survivors are found
        emit_found(out, c, r.alive, id);
        return;
    }
    // the caller keeps n2 <= kQ2Cap - 32
    if (r.alive) {
        double *p = q2 + n2 + __popc(m & c.lt_mask);
        p[0] = __longlong_as_double((long long)id);
        p[1 * kQ2Cap] = r.o.x; p[2 * kQ2Cap] = r.o.y; p[3 * kQ2Cap] = r.o.z;
        p[4 * kQ2Cap] = r.d.x; p[5 * kQ2Cap] = r.d.y; p[6 * kQ2Cap] = r.d.z;
        p[7 * kQ2Cap] = r.w;
    }
    n2 += __popc(m);
    __syncwarp();
}

// ---- spectrometer variant, first half of stage B: intersection point and bounds test
// (_TraceObject.py:180-232, arithmetic of optic_geometry) for `cnt` rays popped from queue 1 -- the ~18 % that
// passed the Bragg pre-test; the ~52 % of those inside the bounds are re-packed into queue b
__device__ __forceinline__ void stage_b1(const XrtSceneDesc &sc, const XrtOpticDesc &ops, const XrtOutputs &out,
                                         const WarpCtx &c, int split, const PhiloxKeys &pk, uint64_t stream_id,
                                         const double *q1, int first, int cnt, double *qb, int &nb) {
    constexpr int P = kQ1CapSpectro;
    const bool active = (int)c.lane < cnt;
    bool inside = false;
    uint64_t id = 0;
    V3 d = v3(0.0, 0.0, 1.0);
    double t = 0.0, gap = -1.0, c2 = 1.0;
    if (active) {
        const double *p = q1 + first + c.lane;
        id = (uint64_t)__double_as_longlong(p[0]);
        d = v3(p[1 * P], p[2 * P], p[3 * P]);
        t = p[4 * P];
        const V3 o = v3(sc.source.origin);
        const V3 X = v3(fma(d.x, t, o.x), fma(d.y, t, o.y), fma(d.z, t, o.z));
        const V3 Xl = to_local(ops.orient, X - v3(ops.origin));
        inside = (fabs(Xl.x) < ops.half_size[0]) && (fabs(Xl.y) < ops.half_size[1]);
        gap = p[5 * P];
        c2 = p[6 * P];
    }
    __syncwarp();
    PhiloxDraws dr;
    dr.init(pk, stream_id, id, split);
    // second level of the Bragg pre-test: with the ray's rocking-curve uniform (the same Philox block stage B2 reads)
    // most of the remaining rays are provably lost: 9.4 % -> about 2 % of the launched rays reach the exact path
    if (ops.cull_t2 > 0.0 && ops.rocking_type != XRT_ROCK_STEP) {
        const double u = dr.bragg_u(split, 0);
        if (bragg_cull_uniform(ops, gap, c2, u)) inside = false;
    }
    if (out.lost_count) emit_lost(out, c, dr, active && !inside, id);
    const unsigned m = __ballot_sync(kFull, inside);
    if (inside) {
        double *p = qb + nb + __popc(m & c.lt_mask);
        p[0] = __longlong_as_double((long long)id);
        p[1 * kQbCap] = d.x; p[2 * kQbCap] = d.y; p[3 * kQbCap] = d.z;
        p[4 * kQbCap] = t;
    }
    nb += __popc(m);
    __syncwarp();
}

// ---- spectrometer variant, FP64 stage A for one ray: direction from the cone block, the two lengths of the sphere
// intersection and the first level of the Bragg pre-test.  Returns true for a ray that goes on to stage B1.
__device__ __forceinline__ bool spectro_stage_a(const XrtSceneDesc &sc, const XrtOpticDesc &ops, const PhiloxKeys &pk,
                                                uint64_t stream_id, int split, const double *s_sincos, uint64_t id, bool valid,
                                                V3 &d_out, double &t_out, double &gap_out, double &c2_out) {
    const XrtSourceDesc &src = sc.source;
    PhiloxDraws dr;
    dr.init(pk, stream_id, id, split);
    double a, b;
    dr.cone(0, a, b);
    const double cs0 = src.cone_par[0];
    const double z = cs0 + (1.0 - cs0) * a;
    const double rho = fast_sqrt(fma(-z, z, 1.0));
    double sn, cs;
    sincos_2pi_tab(b, s_sincos, sn, cs);
    const double lx = rho * cs, ly = rho * sn;
    const double *B = src.axis_basis;
    const V3 d = v3(lx * B[0] + ly * B[3] + z * B[6], lx * B[1] + ly * B[4] + z * B[7], lx * B[2] + ly * B[5] + z * B[8]);
    // hit_sphere, concave
    const V3 Lc = v3(ops.center) - v3(src.origin);
    const double tca = dot(Lc, d);
    const double d2 = fma(-tca, tca, dot(Lc, Lc));
    const double r2 = ops.radius * ops.radius;
    const double thc = fast_sqrt(r2 - d2);          // NaN when d2 > r2
    d_out = d;
    t_out = tca + thc;
    bool cand = valid & (d2 >= 0.0) & (d2 <= r2);
    gap_out = -1.0;
    c2_out = 1.0;
    if (ops.cull_t2 > 0.0) cand &= !bragg_cull_sphere(src, ops, dr.wave_hi(), thc, gap_out, c2_out);
    return cand;
}

// ---- spectrometer variant, FP32 broad phase for one ray (stage A32).  The same Philox block, the same geometry and
// the same first-level test as spectro_stage_a, in single precision with the MUFU units; K = XrtSceneDesc.kn32.
// 1 - z^2 is formed as w (2 - w) with w = 1 - z = (1 - cos spread)(1 - a), 1 - a to 2^-24 relative, so rho keeps its
// relative accuracy near the axis; with that every quantity of the test is within 1e-6 of its FP64 value and K[21] adds
// 2e-5 to the margin of the bound.  true = the ray provably fails the Bragg test (whatever its uniform): lost at the
// crystal.  Everything else -- including rays that miss the sphere or give a NaN here -- is decided in FP64.
__device__ __forceinline__ bool spectro_cull32(const float *K, uint4 r) {
    // 1 - a from all 52 bits of the polar uniform (complement of the mantissa, two exact-or-rounded pieces): its
    // RELATIVE precision is what rho = sqrt(w (2 - w)) near the cone axis needs -- truncating a to 24 bits would put
    // rays within 4e-5 rad of the axis off by that much
    const float a1 = fmaf((float)((~r.y) >> 12), 2.220446049250313e-16f, fmaf((float)(~r.x), 2.3283064365386963e-10f,
                                                                                2.220446049250313e-16f));
    const float w = K[1] * a1;                                                          // 1 - z
    const float z = 1.0f - w;
    float rho;
    asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(rho) : "f"(w * (2.0f - w)));
    const uint32_t b24 = ((r.y & 0xfffu) << 12) | (r.z >> 20);                          // top 24 bits of the azimuth uniform
    const float ang = 6.283185307179586f * ((float)b24 * 5.9604644775390625e-8f - 0.5f);
    const float lx = -rho * __cosf(ang), ly = -rho * __sinf(ang);                       // cos(2 pi b) = -cos(2 pi (b - 1/2))
    const float dx = lx * K[2] + ly * K[5] + z * K[8];
    const float dy = lx * K[3] + ly * K[6] + z * K[9];
    const float dz = lx * K[4] + ly * K[7] + z * K[10];
    const float tca = K[11] * dx + K[12] * dy + K[13] * dz;
    const float d2 = fmaf(-tca, tca, K[14]);
    float thc;
    asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(thc) : "f"(K[15] - d2));
    const float sI = thc * K[16];
    bool usable;
    const float zn = normal_approx(r.w, usable);
    const float sB = fmaf(zn, K[18], K[17]) * K[19];
    const float gap = fabsf(sB - sI);
    const float diff = gap - K[21];
    const float c2 = fmaf(2.0f, gap, fmaf(-sI, sI, 1.0f));
    return usable & (diff > 0.0f) & (diff * diff > K[20] * c2);
}

// Resident blocks per SM: 2 for the mesh variants (face loops and Clough-Tocher cubics keep many values live), for
// the spectrometer variant (two ray groups per pass = two independent chains: 118 registers, no spills) and for the
// lean extended-source variant (bundle lookup + focused cone basis: 116 registers, no spills); 3 otherwise.
template <uint32_t FT, int SPLIT, uint32_t KN>
__global__ void __launch_bounds__(kBlock, (((FT & FT_MESH) != 0 || FT == FT_SRCLEAN ||
                                            ((KN & KN_SPECTROMETER) == KN_SPECTROMETER && XRT_UNROLL > 1)) &&
                                           XRT_MIN_BLOCKS > 2) ? 2 : XRT_MIN_BLOCKS)
k_trace(const __grid_constant__ XrtSceneDesc sc, const __grid_constant__ PhiloxKeys pk, const uint64_t stream_id,
        const uint64_t ray_begin, const uint64_t ray_count, const XrtOutputs out, const int split_rt,
        const int lazy_rt) {
    extern __shared__ double s_queue[];
    __shared__ unsigned long long s_cnt[XRT_MAX_OPTICS + 1];
    if (threadIdx.x <= XRT_MAX_OPTICS) s_cnt[threadIdx.x] = 0ull;
    // (cos, sin)(2 pi k / 256) for sincos_2pi_tab
    __shared__ double s_sincos[2 * kSincosTable];
    for (int i = threadIdx.x; i < kSincosTable; i += kBlock) {
        double sn, cs;
        sincos_2pi((double)i / (double)kSincosTable, sn, cs);
        s_sincos[2 * i] = cs;
        s_sincos[2 * i + 1] = sn;
    }
    __syncthreads();

    constexpr int P = ((KN & KN_SPECTROMETER) == KN_SPECTROMETER) ? kQ1CapSpectro : kQ1Cap;
    WarpCtx c;
    c.lane = threadIdx.x & 31u;
    c.lt_mask = (1u << c.lane) - 1u;
    c.s_cnt = s_cnt;
    const int warp = threadIdx.x >> 5;
    double *q1 = s_queue + (size_t)warp * warp_queue_doubles<FT, KN>();
    double *q2 = q1 + q1_doubles<FT, KN>();

    const int split = (SPLIT >= 0) ? SPLIT : split_rt;
    const XrtOpticDesc &ops = sc.optics[split];
    const bool lazy = (FT == 0) ? true : ((lazy_rt & 1) != 0);
    const bool need_wave = (lazy_rt & 2) != 0;   // some optic from the split optic on reads the wavelength (Bragg test)
    const bool defer = (FT == 0) ? false : ((lazy_rt & 4) != 0);   // eager normal line: exact deviate left to stage B

    // mesh split optic: stage the face operands every ray is tested against in shared memory
    const double *staged = nullptr;
    if constexpr ((FT & FT_MESH) != 0) {
        if (ops.shape == XRT_SHAPE_MESH) {
            const double *geom;
            const int nf = mesh_stage1_faces(ops, geom);
            if (nf <= kStageFaces) {
                double *dst = s_queue + (size_t)(kBlock / 32) * warp_queue_doubles<FT, KN>();
                for (int i = threadIdx.x; i < 9 * nf; i += kBlock) dst[i] = __ldg(geom + i);
                staged = dst;
            }
        }
        __syncthreads();
    }
    constexpr bool SPECTRO_K = (KN & KN_SPECTROMETER) == KN_SPECTROMETER;
    int n1 = 0, n2 = 0;     // queue fill levels, warp-uniform
    int nb = 0;             // spectrometer variant: queue b (inside the bounds), after the planes of queue 1
    double *qb = q1 + kQ1PlanesSpectro * kQ1CapSpectro;
    int n0 = 0;             // spectrometer variant: queue 0 (ids that passed the FP32 broad phase), after queue b
    double *q0 = qb + kQbPlanes * kQbCap;
    const bool broad32 = SPECTRO_K && sc.kn32[0] > 0.0f;
    // mesh variants: queue a (coarse-mesh hits) after the planes of queue 1.  Stage A is split in two when the split
    // optic is the first optic, a refining mesh, and the wavelength is lazy (a ray is rebuilt from its id in stage A2)
    int na = 0;
    double *qa = q1 + q1_planes<FT>() * kQ1Cap;
    bool mesh_staged = false;
    if constexpr ((FT & FT_MESH) != 0 && !SPECTRO_K) {
        mesh_staged = split == 0 && ops.shape == XRT_SHAPE_MESH && (ops.flags & XRT_F_MESH_REFINE) && lazy &&
                      sc.source.kind != XRT_SRC_BUNDLES && sc.source.cone != XRT_CONE_ISOTROPIC_XY;
    }
    unsigned n_src = 0, n_split = 0;   // rays out of the source / the split optic (warp-uniform registers)

    // Ray ids in groups of 32: group g = warp_global + it * n_warps belongs to this warp at iteration it.
    // The loop carries 32-bit counters only; the 64-bit id is one multiply-add per iteration.
    const uint32_t n_warps = gridDim.x * (kBlock / 32);
    const uint32_t warp_global = blockIdx.x * (kBlock / 32) + warp;
    const uint64_t n_groups = (ray_count + 31) >> 5;
    const uint32_t n_it = warp_global < n_groups ? (uint32_t)((n_groups - warp_global + n_warps - 1) / n_warps) : 0u;
    const uint32_t tail = (uint32_t)ray_count & 31u;         // valid lanes of the last group (0 = all)
    const uint32_t tail_it = (tail != 0u && (n_groups - 1) % n_warps == warp_global) ? n_it - 1u : 0xffffffffu;
    const uint64_t id0 = ray_begin + (uint64_t)warp_global * 32u + c.lane;
    const uint32_t stride = n_warps * 32u;
    constexpr bool SPECTRO = (KN & KN_SPECTROMETER) == KN_SPECTROMETER;

    // One loop, one copy of each stage: the deepest stage that has a full warp of work runs
    // first; when the ids are exhausted the queues are drained with partial warps.
    uint32_t it = 0;
    for (;;) {
        const bool more = it < n_it;
        if (n2 >= 32 || (!more && n0 == 0 && n1 == 0 && nb == 0 && na == 0 && n2 > 0)) {
            const int cnt = n2 < 32 ? n2 : 32;
            n2 -= cnt;
            stage_c<FT>(sc, out, c, split, pk, stream_id, q2, n2, cnt);
            continue;
        }
        if constexpr (SPECTRO) {
            if (nb >= 32 || (!more && n0 == 0 && n1 == 0 && nb > 0)) {
                const int cnt = nb < 32 ? nb : 32;
                nb -= cnt;
                stage_b<FT, KN>(sc, ops, out, c, split, lazy, need_wave, defer, pk, stream_id, qb, nb, cnt, q2, n2, n_split);
                continue;
            }
            if (n1 >= 32 || (!more && n0 == 0 && n1 > 0)) {
                const int cnt = n1 < 32 ? n1 : 32;
                n1 -= cnt;
                stage_b1(sc, ops, out, c, split, pk, stream_id, q1, n1, cnt, qb, nb);
                continue;
            }
            if (n0 >= 32 || (!more && n0 > 0)) {
                // ---- stage A64 + B1: the rays the FP32 broad phase could not reject, in FP64 from their ids: first
                // level of the pre-test, intersection point and bounds (arithmetic of optic_geometry), second level
                // with the rocking-curve uniform -- 86 % of them are still candidates after the first level, so the
                // two steps share one pass without a queue in between                               -> queue b
                const int cnt = n0 < 32 ? n0 : 32;
                n0 -= cnt;
                const bool active = (int)c.lane < cnt;
                uint64_t id = 0;
                if (active) id = (uint64_t)__double_as_longlong(q0[n0 + c.lane]);
                __syncwarp();
                V3 d;
                double t, gap, c2;
                bool cand = spectro_stage_a(sc, ops, pk, stream_id, split, s_sincos, id, active, d, t, gap, c2);
                {
                    const V3 o = v3(sc.source.origin);
                    const V3 X = v3(fma(d.x, t, o.x), fma(d.y, t, o.y), fma(d.z, t, o.z));
                    const V3 Xl = to_local(ops.orient, X - v3(ops.origin));
                    cand &= (fabs(Xl.x) < ops.half_size[0]) & (fabs(Xl.y) < ops.half_size[1]);
                }
                PhiloxDraws dr;
                dr.init(pk, stream_id, id, split);
                if (ops.cull_t2 > 0.0 && ops.rocking_type != XRT_ROCK_STEP) {
                    const double u = dr.bragg_u(split, 0);
                    if (bragg_cull_uniform(ops, gap, c2, u)) cand = false;
                }
                if (out.lost_count) emit_lost(out, c, dr, active && !cand, id);
                const unsigned m = __ballot_sync(kFull, cand);
                if (cand) {
                    double *p = qb + nb + __popc(m & c.lt_mask);
                    p[0] = __longlong_as_double((long long)id);
                    p[1 * kQbCap] = d.x; p[2 * kQbCap] = d.y; p[3 * kQbCap] = d.z;
                    p[4 * kQbCap] = t;
                }
                nb += __popc(m);
                __syncwarp();
                continue;
            }
        } else {
            if (n1 >= 32 || (!more && na == 0 && n1 > 0)) {
                const int cnt = n1 < 32 ? n1 : 32;
                n1 -= cnt;
                stage_b<FT, KN>(sc, ops, out, c, split, lazy, need_wave, defer, pk, stream_id, q1, n1, cnt, q2, n2, n_split);
                continue;
            }
            if constexpr ((FT & FT_MESH) != 0) {
                if (na >= 32 || (!more && na > 0)) {
                    // ---- stage A2: the coarse-mesh hits, re-packed: rebuild the ray from its id, finish the mesh
                    // intersection from the coarse hit point (nearest vertex, candidate faces, interpolation), bounds
                    const int cnt = na < 32 ? na : 32;
                    na -= cnt;
                    const bool active = (int)c.lane < cnt;
                    uint64_t id = 0;
                    V3 Xc = nan3();
                    if (active) {
                        const double *p = qa + na + c.lane;
                        id = (uint64_t)__double_as_longlong(p[0]);
                        Xc = v3(p[1 * kQaCap], p[2 * kQaCap], p[3 * kQaCap]);
                    }
                    __syncwarp();
                    PhiloxDraws dr;
                    dr.init(pk, stream_id, id, split);
                    Ray r;
                    r.alive = false;
                    r.w = 0.0;
                    V3 n = v3(0.0, 0.0, 1.0);
                    bool cand = false;
                    if (active) {
                        SrcLocal L;
                        source_local<FT, KN>(sc.source, id, L);
                        generate_geometry<FT, PhiloxDraws, KN, true>(sc.source, L, dr, r, s_sincos);
                        cand = optic_geometry<FT, true, KN>(ops, r, n, staged, &Xc) == HIT_INSIDE;
                    }
                    emit_lost(out, c, dr, active && !cand, id);
                    const unsigned m = __ballot_sync(kFull, cand);
                    if (cand) {
                        double *p = q1 + n1 + __popc(m & c.lt_mask);
                        p[0] = __longlong_as_double((long long)id);
                        p[1 * P] = r.o.x; p[2 * P] = r.o.y; p[3 * P] = r.o.z;
                        p[4 * P] = r.d.x; p[5 * P] = r.d.y; p[6 * P] = r.d.z;
                        p[7 * P] = r.w;
                        p[8 * P] = n.x; p[9 * P] = n.y; p[10 * P] = n.z;
                    }
                    n1 += __popc(m);
                    __syncwarp();
                    continue;
                }
            }
        }
        if (!more) break;

        // ---- stage A
        if constexpr (SPECTRO) {
            // Straight-line code for the spectrometer (point source, concave sphere): direction, the two
            // lengths of the sphere intersection and the Bragg pre-test; no branch, no intersection point.
            // sin(theta_i) = |D.n| is thc / R for a ray of unit direction (D.(C - X) = tca - t = -thc), so the
            // pre-test needs nothing else.  A ray it rejects is lost at the crystal whether or not it is inside
            // the bounds, so the bounds test moves to stage B, behind the queue (18 % of the rays).
            // kUnroll groups of 32 rays per pass: independent dependency chains for the scheduler.
            if (broad32) {
                // ---- stage A32: FP32 broad phase, kUnroll32 groups per pass; the ~20 % it cannot reject go to
                // queue 0 as bare ids
                uint64_t id32[kUnroll32];
                bool valid32[kUnroll32], pass32[kUnroll32];
#pragma unroll
                for (int j = 0; j < kUnroll32; ++j) {
                    const uint32_t itj = it + (uint32_t)j;
                    id32[j] = id0 + (uint64_t)itj * stride;
                    valid32[j] = (itj < n_it) && ((itj != tail_it) || (c.lane < tail));
                    PhiloxDraws dr;
                    dr.init(pk, stream_id, id32[j], split);
                    pass32[j] = valid32[j] & !spectro_cull32(sc.kn32, dr.raw(SITE_CONE));
                }
                it += kUnroll32;
#pragma unroll
                for (int j = 0; j < kUnroll32; ++j) {
                    n_src += __popc(__ballot_sync(kFull, valid32[j]));
                    if (out.lost_count) {
                        PhiloxDraws dr;
                        dr.init(pk, stream_id, id32[j], split);
                        emit_lost(out, c, dr, valid32[j] && !pass32[j], id32[j]);
                    }
                    const unsigned m = __ballot_sync(kFull, pass32[j]);
                    if (pass32[j]) q0[n0 + __popc(m & c.lt_mask)] = __longlong_as_double((long long)id32[j]);
                    n0 += __popc(m);
                }
                __syncwarp();
                continue;
            }
            uint64_t idv[kUnroll];
            bool validv[kUnroll], candv[kUnroll];
            V3 dv[kUnroll];
            double tv[kUnroll];
            double gapv[kUnroll], c2v[kUnroll];
#pragma unroll
            for (int j = 0; j < kUnroll; ++j) {
                const uint32_t itj = it + (uint32_t)j;
                idv[j] = id0 + (uint64_t)itj * stride;
                validv[j] = (itj < n_it) && ((itj != tail_it) || (c.lane < tail));
                candv[j] = spectro_stage_a(sc, ops, pk, stream_id, split, s_sincos, idv[j], validv[j], dv[j], tv[j], gapv[j], c2v[j]);
            }
            it += kUnroll;
#pragma unroll
            for (int j = 0; j < kUnroll; ++j) {
                n_src += __popc(__ballot_sync(kFull, validv[j]));
                if (out.lost_count) {
                    PhiloxDraws dr;
                    dr.init(pk, stream_id, idv[j], split);
                    emit_lost(out, c, dr, validv[j] && !candv[j], idv[j]);
                }
                const unsigned m = __ballot_sync(kFull, candv[j]);
                if (candv[j]) {
                    double *p = q1 + n1 + __popc(m & c.lt_mask);
                    p[0] = __longlong_as_double((long long)idv[j]);
                    p[1 * P] = dv[j].x; p[2 * P] = dv[j].y; p[3 * P] = dv[j].z;
                    p[4 * P] = tv[j];
                    p[5 * P] = gapv[j];
                    p[6 * P] = c2v[j];
                }
                n1 += __popc(m);
            }
            __syncwarp();
        } else {
            const uint64_t id = id0 + (uint64_t)it * stride;
            const bool valid = (it != tail_it) || (c.lane < tail);
            ++it;
            PhiloxDraws dr;
            dr.init(pk, stream_id, id, split);
            Ray r;
            r.alive = false;
            r.w = 0.0;
            double sigma_a = 0.0;
            if (valid) {
                SrcLocal L;
                source_local<FT, KN>(sc.source, id, L);
                generate_geometry<FT, PhiloxDraws, KN, true>(sc.source, L, dr, r, s_sincos);
                if (defer) {
                    // normal line with a Doppler shift and / or a per-bundle sigma: the exact deviate (inverse normal
                    // CDF, a second Philox block) is left to stage B; here the Doppler factor, and sigma for the pre-test
                    const bool moving = L.vel.x != 0.0 || L.vel.y != 0.0 || L.vel.z != 0.0;
                    r.w = moving ? 1.0 - dot(L.vel, r.d) : 1.0;
                    sigma_a = L.wave_sigma;
                } else if (!lazy) {
                    r.w = generate_wavelength<PhiloxDraws, KN>(sc.source, L, dr, r.d);
                }
            }
            n_src += __popc(__ballot_sync(kFull, r.alive));
            for (int k = 0; k < split; ++k) {
                const XrtOpticDesc &op = sc.optics[k];
                if (r.alive) {
                    trace_optic<FT>(op, k, dr, r);
                    if (r.alive && (op.flags & XRT_F_IMAGE) && out.images) add_pixel(out, op, r, c.lt_mask);
                }
                count_alive(c, k + 1, r.alive);
            }
            V3 n = v3(0.0, 0.0, 1.0);
            bool cand = false;
            if constexpr ((FT & FT_MESH) != 0) {
                if (mesh_staged) {
                    // ---- stage A1: coarse mesh only; the hits go to queue a as (id, coarse hit point)
                    V3 Xc = nan3();
                    bool hit = false;
                    if (r.alive) {
                        V3 o = r.o, d = r.d;
                        if (optic_is_local<FT>(ops)) {
                            o = to_local(ops.orient, o - v3(ops.origin));
                            d = to_local(ops.orient, d);
                        }
                        hit = mesh_coarse_hit(ops, o, d, Xc, staged);
                    }
                    emit_lost(out, c, dr, valid && !hit, id);
                    const unsigned mh = __ballot_sync(kFull, hit);
                    if (hit) {
                        double *p = qa + na + __popc(mh & c.lt_mask);
                        p[0] = __longlong_as_double((long long)id);
                        p[1 * kQaCap] = Xc.x; p[2 * kQaCap] = Xc.y; p[3 * kQaCap] = Xc.z;
                    }
                    na += __popc(mh);
                    __syncwarp();
                    continue;
                }
            }
            if (r.alive) cand = optic_geometry<FT, (FT & FT_MESH) != 0, KN>(ops, r, n, staged) == HIT_INSIDE;
            // Bragg pre-test (bragg_cull_general): enabled by xrt_scene_create for a spherical Bragg crystal
            // traced in global coordinates; the survivors take the exact path in stage B
            if (ops.cull_t2 > 0.0) {
                const int mode = defer ? WAVE_DEFERRED : (lazy ? WAVE_APPROX : WAVE_EXACT);
                if (cand && bragg_cull_general(sc.source, ops, mode, r.w, sigma_a, dr.wave_hi(), r.o, r.d)) {
                    cand = false;
                    r.alive = false;
                }
            }
            emit_lost(out, c, dr, valid && !cand, id);

            const unsigned m = __ballot_sync(kFull, cand);
            if (cand) {
                double *p = q1 + n1 + __popc(m & c.lt_mask);
                p[0] = __longlong_as_double((long long)id);
                p[1 * P] = r.o.x; p[2 * P] = r.o.y; p[3 * P] = r.o.z;
                p[4 * P] = r.d.x; p[5 * P] = r.d.y; p[6 * P] = r.d.z;
                if constexpr (FT != 0) p[7 * P] = r.w;
                if constexpr ((FT & FT_MESH) != 0) { p[8 * P] = n.x; p[9 * P] = n.y; p[10 * P] = n.z; }
            }
            n1 += __popc(m);
            __syncwarp();
        }
    }

    if (c.lane == 0) {
        if (n_src) atomicAdd(&s_cnt[0], (unsigned long long)n_src);
        if (n_split) atomicAdd(&s_cnt[split + 1], (unsigned long long)n_split);
    }
    __syncthreads();
    if ((int)threadIdx.x <= sc.n_optics && out.counts) {
        unsigned long long cc = s_cnt[threadIdx.x];
        if (cc) atomicAdd((unsigned long long *)(out.counts + threadIdx.x), cc);
    }
}

// ---------------------------------------------------------------------------
// recording kernel: history of every element, optional counters / images

// streaming stores (evict-first): history planes are written once and read back by the host
__device__ __forceinline__ void store_history(const XrtHistory &h, int elem, uint64_t slot, const Ray &r) {
    if (h.rays) {
        double *p = h.rays + ((uint64_t)elem * 7) * h.capacity + slot;
        const uint64_t c = h.capacity;
        __stcs(p, r.o.x); __stcs(p + c, r.o.y); __stcs(p + 2 * c, r.o.z);
        __stcs(p + 3 * c, r.d.x); __stcs(p + 4 * c, r.d.y); __stcs(p + 5 * c, r.d.z);
        __stcs(p + 6 * c, r.w);
    }
    if (h.mask) h.mask[(uint64_t)elem * h.capacity + slot] = r.alive ? 1 : 0;
}

enum { REC_PHILOX = 0, REC_INJECT = 1 };

// KN != 0: the scene has the known structure (split optic = optic 0), see k_trace
template <uint32_t FT, int MODE, uint32_t KN = 0>
__global__ void __launch_bounds__(kBlock, (FT & FT_MESH) != 0 ? 2 : XRT_RECORD_BLOCKS)   // 3 x 256 at 78 registers: the replay is latency bound
k_record(const __grid_constant__ XrtSceneDesc sc, const __grid_constant__ PhiloxKeys pk, const uint64_t stream_id,
         const uint64_t *__restrict__ ids, const uint64_t ray_begin, const uint64_t n,
         const XrtRaysIn in, const XrtInject inj, const XrtOutputs out, const XrtHistory hist, const int split) {
    const int nopt = sc.n_optics;
    const uint64_t stride = (uint64_t)gridDim.x * kBlock;
    for (uint64_t i = (uint64_t)blockIdx.x * kBlock + threadIdx.x; i < n; i += stride) {
        Ray r;
        PhiloxDraws pdr;
        InjectedDraws idr;
        if constexpr (MODE == REC_PHILOX) {
            const uint64_t id = ids ? ids[i] : ray_begin + i;
            pdr.init(pk, stream_id, id, split);
            generate_ray<FT, PhiloxDraws, KN>(sc.source, pdr, id, r);
        } else {
            r.o = v3(in.origin + 3 * i);
            r.d = v3(in.direction + 3 * i);
            r.w = in.wavelength[i];
            r.alive = in.mask[i] != 0;
            idr.inj = &inj;
            idr.i = i;
            idr.n = n;
        }
        store_history(hist, 0, i, r);
        if (out.counts && r.alive) atomicAdd((unsigned long long *)out.counts, 1ull);

        for (int k = 0; k < nopt; ++k) {
            const XrtOpticDesc &op = sc.optics[k];
            if (r.alive) {
                if constexpr (MODE == REC_PHILOX) {
                    if (KN != 0 && k == 0) trace_optic<FT, PhiloxDraws, KN>(op, k, pdr, r);
                    else trace_optic<FT>(op, k, pdr, r);
                } else {
                    trace_optic<FT>(op, k, idr, r);
                }
                if (r.alive) {
                    if (out.counts) atomicAdd((unsigned long long *)(out.counts + k + 1), 1ull);
                    uint32_t pix;
                    if ((op.flags & XRT_F_IMAGE) && out.images && pixel_index(op, r.o, pix))
                        atomicAdd((unsigned long long *)(out.images + op.image_offset + pix), 1ull);
                }
            } else {
                pass_lost_ray<FT>(op, r);   // lost earlier: the reference carries NaN origins forward
            }
            store_history(hist, k + 1, i, r);
        }
    }
}

// ---------------------------------------------------------------------------
// source only

template <int MODE>
__global__ void __launch_bounds__(kBlock)
k_source(const __grid_constant__ XrtSceneDesc sc, const __grid_constant__ PhiloxKeys pk, const uint64_t stream_id,
         const uint64_t ray_begin, const uint64_t n, const XrtSourceInject sinj, const XrtHistory hist) {
    const uint64_t stride = (uint64_t)gridDim.x * kBlock;
    for (uint64_t i = (uint64_t)blockIdx.x * kBlock + threadIdx.x; i < n; i += stride) {
        Ray r;
        if constexpr (MODE == REC_PHILOX) {
            PhiloxDraws dr;
            dr.init(pk, stream_id, ray_begin + i, -1);
            generate_ray<FT_FULL>(sc.source, dr, ray_begin + i, r);
        } else {
            SourceInjectedDraws dr;
            dr.inj = &sinj;
            dr.i = i;
            dr.n = n;
            generate_ray<FT_FULL>(sc.source, dr, ray_begin + i, r);
        }
        store_history(hist, 0, i, r);
    }
}

// ---------------------------------------------------------------------------
// FP64 pipe microbenchmark: 8 independent dependent-FMA chains per thread

constexpr int kBurnChains = 8;

__global__ void __launch_bounds__(kBlock) k_burn(uint64_t iters, double *sink) {
    double a[kBurnChains];
    const double m = 1.0 + 1e-9 * (double)(threadIdx.x & 7), c = 1e-12;
#pragma unroll
    for (int j = 0; j < kBurnChains; ++j) a[j] = 1.0 + (double)j * 1e-3 + (double)threadIdx.x * 1e-6;
    for (uint64_t i = 0; i < iters; ++i) {
#pragma unroll
        for (int j = 0; j < kBurnChains; ++j) a[j] = fma(a[j], m, c);
    }
    double s = 0.0;
#pragma unroll
    for (int j = 0; j < kBurnChains; ++j) s += a[j];
    if (s == 123.456) sink[0] = s;   // never true; keeps the chains alive
}

}  // namespace xrt

// ===========================================================================
// host side

using namespace xrt;

struct XrtScene {
    XrtSceneDesc dev;               // descriptor whose pointers are device pointers
    std::vector<void *> allocs;
    void *bundle_hint = nullptr;    // ray id -> bundle bracket table of the current bundle table
    uint32_t features;
    int split;                      // first crystal of the train (0 if none): the kernel's re-pack point
    int lazy_wavelength;            // wavelength independent of the source direction: drawn at the crystal
    int need_wavelength;            // an optic at or after the split optic reads the wavelength (Bragg test / mosaic cutoff)
    int defer_wavelength;           // eager normal line + Bragg pre-test: exact deviate drawn in stage B
    uint32_t known;                 // KN_* facts that hold for this scene (source + split optic)
    int device;
    int sm_count;
};

static thread_local char g_err[512] = "";

static int fail(int code, const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}

#define CU(call)                                                                         \
    do {                                                                                 \
        cudaError_t e_ = (call);                                                         \
        if (e_ != cudaSuccess)                                                           \
            return fail(XRT_ECUDA, "%s: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
    } while (0)

// Scene tables come from the device's stream-ordered memory pool (cudaMallocAsync): a scene is
// created and destroyed per run, and plain cudaMalloc / cudaFree cost milliseconds each and
// synchronise the device (100 ms per run for a plasma bundle table).  The pool keeps freed
// blocks for the next scene.
static int pool_keep_memory(int device) {
    static thread_local int configured_for = -1;
    if (configured_for == device) return XRT_OK;
    cudaMemPool_t pool;
    CU(cudaDeviceGetDefaultMemPool(&pool, device));
    unsigned long long keep = ~0ull;
    CU(cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep));
    configured_for = device;
    return XRT_OK;
}

template <class T>
static int upload(XrtScene *s, const T *host, size_t count, const T **dev) {
    *dev = nullptr;
    if (host == nullptr || count == 0) return XRT_OK;
    void *p = nullptr;
    CU(cudaMallocAsync(&p, count * sizeof(T), (cudaStream_t)0));
    s->allocs.push_back(p);
    CU(cudaMemcpyAsync(p, host, count * sizeof(T), cudaMemcpyHostToDevice, (cudaStream_t)0));
    *dev = (const T *)p;
    return XRT_OK;
}

#define UP(field, count)                                                   \
    do {                                                                   \
        int rc_ = upload(s, field, (size_t)(count), &field);               \
        if (rc_ != XRT_OK) return rc_;                                     \
    } while (0)

static int upload_mesh(XrtScene *s, const XrtMesh *host, const XrtMesh **dev) {
    XrtMesh m = *host;
    if (m.n_points <= 0 || m.n_faces <= 0 || !m.points || !m.faces || !m.face_normals || !m.face_geom || !m.face_area || !m.face_rec || !m.vertex_faces)
        return fail(XRT_EINVAL, "mesh: points / faces missing");
    size_t cells = (size_t)m.grid_nx * (size_t)m.grid_ny;
    int32_t n_items = 0, n_vitems = 0;
    if (cells && m.grid_start) n_items = m.grid_start[cells];
    if (cells && m.vgrid_start) n_vitems = m.vgrid_start[cells];
    UP(m.points, 3 * (size_t)m.n_points);
    UP(m.faces, 3 * (size_t)m.n_faces);
    UP(m.face_normals, 3 * (size_t)m.n_faces);
    UP(m.face_geom, 9 * (size_t)m.n_faces);
    UP(m.face_area, (size_t)m.n_faces);
    UP(m.face_rec, 16 * (size_t)m.n_faces);
    UP(m.vertex_faces, 8 * (size_t)m.n_points);
    UP(m.coarse_points, 3 * (size_t)m.n_coarse_points);
    UP(m.coarse_faces, 3 * (size_t)m.n_coarse_faces);
    UP(m.coarse_geom, 9 * (size_t)m.n_coarse_faces);
    UP(m.point_faces, 8 * (size_t)m.n_points);
    UP(m.point_faces_mask, 8 * (size_t)m.n_points);
    UP(m.ct_coef, 4 * 19 * (size_t)m.n_tri);
    UP(m.tri_transform, 6 * (size_t)m.n_tri);
    UP(m.grid_start, (cells && m.grid_start) ? cells + 1 : 0);
    UP(m.grid_items, n_items);
    UP(m.vgrid_start, (cells && m.vgrid_start) ? cells + 1 : 0);
    UP(m.vgrid_items, n_vitems);
    UP(m.vgrid_xyz, (cells && m.vgrid_start && m.vgrid_xyz) ? 4 * (size_t)n_vitems : 0);
    const XrtMesh *d = nullptr;
    int rc = upload(s, &m, 1, &d);
    if (rc != XRT_OK) return rc;
    *dev = d;
    return XRT_OK;
}

static uint32_t scene_features(const XrtSceneDesc &d) {
    uint32_t ft = 0;
    const XrtSourceDesc &src = d.source;
    if (src.kind == XRT_SRC_BUNDLES || src.spatial != XRT_SPATIAL_UNIFORM || src.cone != XRT_CONE_ISOTROPIC ||
        src.n_sightlines > 0)
        ft |= FT_SRC_EXT;
    for (int k = 0; k < d.n_optics; ++k) {
        const XrtOpticDesc &op = d.optics[k];
        if (op.flags & XRT_F_TRACE_LOCAL) ft |= FT_LOCAL;
        if (op.shape == XRT_SHAPE_CYLINDER) ft |= FT_CYL;
        if (op.shape == XRT_SHAPE_TORUS) ft |= FT_TORUS;
        if (op.shape == XRT_SHAPE_MESH) ft |= FT_MESH;
        if ((op.flags & XRT_F_CHECK_APERTURE) && op.n_aperture > 0) ft |= FT_APERTURE;
        if (op.interact == XRT_INTERACT_MOSAIC) ft |= FT_MOSAIC;
        if (op.rocking_type == XRT_ROCK_TABLE &&
            (op.interact == XRT_INTERACT_CRYSTAL || op.interact == XRT_INTERACT_MOSAIC))
            ft |= FT_ROCKTAB;
    }
    // compiled variants: lean spectrometer, all analytic features, lean mesh, everything
    if (ft == 0) return 0;
    if (ft == FT_MOSAICLEAN || ft == FT_SRCLEAN) return ft;     // one extra feature: far less code than FT_MID
    if ((ft & ~FT_MID) == 0) return FT_MID;
    if ((ft & ~FT_MESHLEAN) == 0) return FT_MESHLEAN;
    return FT_FULL;
}

extern "C" int xrt_version(void) { return XRT_VERSION; }

extern "C" const char *xrt_last_error(void) { return g_err; }

extern "C" int xrt_scene_destroy(XrtScene *s) {
    if (!s) return XRT_OK;
    // stream-ordered on the legacy stream: waits for kernels of every blocking stream that still read the tables
    for (void *p : s->allocs) cudaFreeAsync(p, (cudaStream_t)0);
    if (s->bundle_hint) cudaFreeAsync(s->bundle_hint, (cudaStream_t)0);
    delete s;
    return XRT_OK;
}

static int build_bundle_hint(XrtScene *s, uint64_t n_rays);

static int scene_build(XrtScene *s, const XrtSceneDesc *desc) {
    s->dev = *desc;
    XrtSceneDesc &d = s->dev;
    XrtSourceDesc &src = d.source;

    if (src.kind < XRT_SRC_FIXED_AXIS || src.kind > XRT_SRC_BUNDLES) return fail(XRT_EINVAL, "source.kind = %d", src.kind);
    if (src.cone < XRT_CONE_ISOTROPIC || src.cone > XRT_CONE_FLAT_XY) return fail(XRT_EINVAL, "source.cone = %d", src.cone);
    if (src.wave < XRT_WAVE_CONST || src.wave > XRT_WAVE_TABLE) return fail(XRT_EINVAL, "source.wave = %d", src.wave);
    if (src.n_sightlines < 0 || src.n_sightlines > XRT_MAX_SIGHTLINES)
        return fail(XRT_EINVAL, "source.n_sightlines = %d", src.n_sightlines);
    if (src.wave == XRT_WAVE_TABLE && src.kind != XRT_SRC_BUNDLES && (src.n_table < 2 || !src.table_cdf || !src.table_x))
        return fail(XRT_EINVAL, "source wavelength table missing");
    if (src.wave == XRT_WAVE_TABLE && src.kind == XRT_SRC_BUNDLES && src.n_table < 2)
        return fail(XRT_EINVAL, "plasma source: n_table = %d", src.n_table);
    const bool host_bundles = src.kind == XRT_SRC_BUNDLES && src.bundles && src.bundle_end && src.n_bundles > 0;
    if (src.kind == XRT_SRC_BUNDLES && !host_bundles) {   // table supplied later by xrt_scene_set_bundles
        src.bundles = nullptr;
        src.bundle_end = nullptr;
        src.n_bundles = 0;
    }
    const bool host_tables = host_bundles && src.wave == XRT_WAVE_TABLE && src.bundle_x && src.bundle_cdf;
    if (src.kind != XRT_SRC_BUNDLES) { src.bundle_x = nullptr; src.bundle_cdf = nullptr; }
    if (src.kind == XRT_SRC_BUNDLES && !host_tables) {    // supplied later by xrt_scene_set_bundle_tables
        src.bundle_x = nullptr;
        src.bundle_cdf = nullptr;
    }
    {
        XrtSourceDesc &m = src;
        const bool one_table = m.wave == XRT_WAVE_TABLE && m.kind != XRT_SRC_BUNDLES;
        if (!one_table) { m.table_cdf = nullptr; m.table_x = nullptr; }
        UP(m.table_cdf, one_table ? m.n_table : 0);
        UP(m.table_x, one_table ? m.n_table : 0);
        UP(m.bundle_x, host_tables ? m.n_bundles * (uint64_t)m.n_table : 0);
        UP(m.bundle_cdf, host_tables ? m.n_bundles * (uint64_t)m.n_table : 0);
        const uint64_t host_rays = host_bundles ? m.bundle_end[m.n_bundles - 1] : 0;
        UP(m.bundles, host_bundles ? m.n_bundles : 0);
        UP(m.bundle_end, host_bundles ? m.n_bundles : 0);
        m.bundle_hint = nullptr;
        m.bundle_hint_shift = 0;
        if (host_bundles) {
            int rc_ = build_bundle_hint(s, host_rays);
            if (rc_ != XRT_OK) return rc_;
        }
    }

    for (int k = 0; k < d.n_optics; ++k) {
        XrtOpticDesc &m = d.optics[k];
        if (m.shape < XRT_SHAPE_PLANE || m.shape > XRT_SHAPE_MESH) return fail(XRT_EINVAL, "optic %d: shape = %d", k, m.shape);
        if (m.interact < XRT_INTERACT_NONE || m.interact > XRT_INTERACT_MOSAIC)
            return fail(XRT_EINVAL, "optic %d: interact = %d", k, m.interact);
        const bool crystal = (m.interact == XRT_INTERACT_CRYSTAL || m.interact == XRT_INTERACT_MOSAIC);
        if (crystal && (m.flags & XRT_F_CHECK_BRAGG)) {
            if (m.rocking_type < XRT_ROCK_STEP || m.rocking_type > XRT_ROCK_TABLE)
                return fail(XRT_EINVAL, "optic %d: rocking_type = %d", k, m.rocking_type);
            if (m.rocking_type == XRT_ROCK_TABLE && (m.n_rock < 2 || !m.rock_dtheta || !m.rock_s || !m.rock_p))
                return fail(XRT_EINVAL, "optic %d: rocking table missing", k);
            if (!(m.two_d > 0.0)) return fail(XRT_EINVAL, "optic %d: crystal_spacing must be > 0", k);
        }
        if (m.shape == XRT_SHAPE_TORUS && (m.root_idx < 0 || m.root_idx > 3))
            return fail(XRT_EINVAL, "optic %d: root_idx = %d", k, m.root_idx);
        if ((m.flags & XRT_F_IMAGE) && (m.npix[0] <= 0 || m.npix[1] <= 0 || !(m.pixel_size > 0.0)))
            return fail(XRT_EINVAL, "optic %d: bad pixel grid", k);
        if (m.n_aperture < 0) return fail(XRT_EINVAL, "optic %d: n_aperture = %d", k, m.n_aperture);
        UP(m.apertures, m.n_aperture);
        const bool tab = crystal && m.rocking_type == XRT_ROCK_TABLE;
        UP(m.rock_dtheta, tab ? m.n_rock : 0);
        UP(m.rock_s, tab ? m.n_rock : 0);
        UP(m.rock_p, tab ? m.n_rock : 0);
        if (m.shape == XRT_SHAPE_MESH) {
            if (!m.mesh) return fail(XRT_EINVAL, "optic %d: mesh tables missing", k);
            if ((m.flags & XRT_F_MESH_REFINE) && (m.mesh->n_coarse_faces <= 0 || !m.mesh->coarse_geom ||
                                                  !m.mesh->vgrid_start || !m.mesh->vgrid_xyz))
                return fail(XRT_EINVAL, "optic %d: mesh refinement needs the coarse mesh and the vertex grid", k);
            if ((m.flags & XRT_F_MESH_INTERP) && (m.mesh->n_tri <= 0 || !m.mesh->ct_coef || !m.mesh->tri_transform ||
                                                  !m.mesh->grid_start))
                return fail(XRT_EINVAL, "optic %d: mesh interpolation needs the Clough-Tocher tables", k);
            int rc = upload_mesh(s, m.mesh, &m.mesh);
            if (rc != XRT_OK) return rc;
        } else {
            m.mesh = nullptr;
        }
    }
    for (int k = 0; k < d.n_optics; ++k) d.optics[k].cull_t2 = d.optics[k].cull_err = d.optics[k].cull_inv_r = 0.0;
    for (int i = 0; i < 32; ++i) d.kn32[i] = 0.0f;
    s->features = scene_features(d);
    s->split = 0;
    for (int k = 0; k < d.n_optics; ++k) {
        if (d.optics[k].interact == XRT_INTERACT_CRYSTAL || d.optics[k].interact == XRT_INTERACT_MOSAIC) {
            s->split = k;
            break;
        }
    }
    s->lazy_wavelength = (src.kind != XRT_SRC_BUNDLES && src.velocity_c[0] == 0.0 && src.velocity_c[1] == 0.0 &&
                          src.velocity_c[2] == 0.0) ? 1 : 0;
    s->need_wavelength = 0;
    s->defer_wavelength = 0;
    for (int k = s->split; k < d.n_optics; ++k) {
        const XrtOpticDesc &o = d.optics[k];
        const bool crystal = o.interact == XRT_INTERACT_CRYSTAL || o.interact == XRT_INTERACT_MOSAIC;
        if (crystal && (o.flags & (XRT_F_CHECK_BRAGG | XRT_F_MOSAIC_CUTOFF))) s->need_wavelength = 1;
    }
    // the lean variant has no wavelength plane in its queue: a source with a Doppler shift takes the lean
    // extended-source variant
    if (s->features == 0 && !s->lazy_wavelength) s->features = FT_SRCLEAN;
    s->known = 0;
    if (d.n_optics > 0) {
        const XrtOpticDesc &o = d.optics[s->split];
        if (src.kind == XRT_SRC_FIXED_AXIS && src.extent[0] == 0.0 && src.extent[1] == 0.0 && src.extent[2] == 0.0)
            s->known |= KN_POINT_SOURCE;
        if (src.wave == XRT_WAVE_NORMAL) s->known |= KN_WAVE_NORMAL;
        if (o.shape == XRT_SHAPE_SPHERE && !(o.flags & XRT_F_CONVEX)) s->known |= KN_SPHERE;
        const uint32_t size_bits = XRT_F_CHECK_SIZE | XRT_F_HAS_XSIZE | XRT_F_HAS_YSIZE | XRT_F_HAS_ZSIZE;
        if ((o.flags & size_bits) == (XRT_F_CHECK_SIZE | XRT_F_HAS_XSIZE | XRT_F_HAS_YSIZE)) s->known |= KN_BOUNDS_XY;
        if (o.interact == XRT_INTERACT_CRYSTAL && (o.flags & XRT_F_CHECK_BRAGG) &&
            (o.rocking_type == XRT_ROCK_GAUSS || o.rocking_type == XRT_ROCK_STEP))
            s->known |= KN_CRYSTAL_GAUSS;
        if (o.flags & XRT_F_IMAGE) s->known |= KN_IMAGE;
        // parameters of the Bragg pre-test (bragg_cull_* in xrt_trace.cuh): a spherical Bragg crystal with a
        // Gaussian or step rocking curve, traced in global coordinates, as split optic; the wavelength is either
        // drawn before the crystal (sources with a Doppler shift, plasma bundles) or it is a constant / normal line
        const bool rock_ok = (o.rocking_type == XRT_ROCK_GAUSS && o.rock_inv_two_sigma2 > 0.0 && std::isfinite(o.rock_inv_two_sigma2)) ||
                             (o.rocking_type == XRT_ROCK_STEP && o.rocking_fwhm >= 0.0 && std::isfinite(o.rocking_fwhm));
        const bool wave_ok = !s->lazy_wavelength || src.wave == XRT_WAVE_NORMAL || src.wave == XRT_WAVE_CONST;
        if (o.shape == XRT_SHAPE_SPHERE && o.interact == XRT_INTERACT_CRYSTAL && (o.flags & XRT_F_CHECK_BRAGG) &&
            !(o.flags & XRT_F_TRACE_LOCAL) && rock_ok && wave_ok && o.radius > 0.0 && std::isfinite(o.inv_two_d) &&
            !(s->features & FT_MESH) && std::getenv("XRT_NO_CULL") == nullptr) {
            XrtOpticDesc &w = d.optics[s->split];
            const double edge = o.rocking_type == XRT_ROCK_GAUSS ? std::sqrt(40.0 / o.rock_inv_two_sigma2) : 0.5 * o.rocking_fwhm;
            const double t = 1.05 * edge + 2e-6;
            w.cull_t2 = t * t;
            const bool approx = s->lazy_wavelength && src.wave == XRT_WAVE_NORMAL;
            w.cull_err = (approx ? 2e-3 * std::fabs(src.wave_par[1]) * std::fabs(o.inv_two_d) : 0.0) + 1e-9;
            w.cull_inv_r = 1.0 / o.radius;
            // FP32 broad phase of the spectrometer variant (spectro_cull32): single-precision copies of its constants.
            // Its arithmetic error on sin(theta_i) is about 1e-6 (|C - O| ~ R); the margin is 2e-5, scaled with
            // |C - O|^2 / R^2, and the phase is left off for a source farther than 2 R from the centre of curvature.
            if ((s->known & KN_SPECTROMETER) == KN_SPECTROMETER && s->features == 0 && approx &&
                std::getenv("XRT_NO_BROAD32") == nullptr) {
                float *K = d.kn32;
                const double lx = o.center[0] - src.origin[0], ly = o.center[1] - src.origin[1], lz = o.center[2] - src.origin[2];
                const double ll = lx * lx + ly * ly + lz * lz, r2 = o.radius * o.radius;
                // ... and for Bragg angles below 6 degrees, where thc -> sI amplifies the error of thc^2 by 1 / (2 sI)
                if (ll <= 4.0 * r2 && std::fabs(src.wave_par[0] * o.inv_two_d) >= 0.1) {
                    K[1] = (float)(1.0 - src.cone_par[0]);
                    for (int i = 0; i < 9; ++i) K[2 + i] = (float)src.axis_basis[i];
                    K[11] = (float)lx; K[12] = (float)ly; K[13] = (float)lz;
                    K[14] = (float)ll;
                    K[15] = (float)r2;
                    K[16] = (float)(1.0 / o.radius);
                    K[17] = (float)src.wave_par[0];
                    K[18] = (float)src.wave_par[1];
                    K[19] = (float)o.inv_two_d;
                    K[20] = (float)w.cull_t2;
                    K[21] = (float)(w.cull_err + 2e-5 * std::fmax(1.0, ll / r2));
                    K[0] = 1.0f;
                }
            }
            // eager normal line (plasma bundles, Doppler shift) with no optic before the crystal: defer the exact deviate
            s->defer_wavelength = (!s->lazy_wavelength && src.wave == XRT_WAVE_NORMAL && s->split == 0) ? 1 : 0;
        }
    }
    return XRT_OK;
}

extern "C" int xrt_scene_create(const XrtSceneDesc *desc, XrtScene **scene) {
    if (!desc || !scene) return fail(XRT_EINVAL, "null argument");
    *scene = nullptr;
    if (desc->version != XRT_VERSION) return fail(XRT_EINVAL, "descriptor version %d, library %d", desc->version, XRT_VERSION);
    if (desc->n_optics < 0 || desc->n_optics > XRT_MAX_OPTICS)
        return fail(XRT_EINVAL, "n_optics = %d (max %d)", desc->n_optics, XRT_MAX_OPTICS);
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
        cudaGetLastError();
        return fail(XRT_ECUDA, "no CUDA device: libxrt has no CPU path");
    }
    XrtScene *s = new (std::nothrow) XrtScene();
    if (!s) return fail(XRT_ENOMEM, "out of host memory");
    cudaError_t e = cudaGetDevice(&s->device);
    if (e == cudaSuccess) e = cudaDeviceGetAttribute(&s->sm_count, cudaDevAttrMultiProcessorCount, s->device);
    if (e != cudaSuccess) {
        delete s;
        return fail(XRT_ECUDA, "cudaGetDevice: %s", cudaGetErrorString(e));
    }
    int rc = pool_keep_memory(s->device);
    if (rc == XRT_OK) rc = scene_build(s, desc);
    // host tables may be released by the caller as soon as this returns
    if (rc == XRT_OK && cudaStreamSynchronize((cudaStream_t)0) != cudaSuccess) rc = fail(XRT_ECUDA, "scene upload failed");
    if (rc != XRT_OK) {
        xrt_scene_destroy(s);
        return rc;
    }
    *scene = s;
    return XRT_OK;
}

// ---- launch helpers -------------------------------------------------------

typedef void (*TraceKernel)(const XrtSceneDesc, const PhiloxKeys, const uint64_t, const uint64_t, const uint64_t,
                            const XrtOutputs, const int, const int);

template <uint32_t FT>
static TraceKernel trace_kernel_ft(int split) {
    switch (split) {
    case 0: return k_trace<FT, 0, 0>;
    case 1: return k_trace<FT, 1, 0>;
    case 2: return k_trace<FT, 2, 0>;
    default: return k_trace<FT, -1, 0>;
    }
}

static TraceKernel trace_kernel(const XrtScene *s, size_t *smem) {
    if (s->features == 0) {
        *smem = block_smem_bytes<0>();
        // pre-instantiated structure: point source with a Gaussian line on a concave spherical
        // Bragg crystal as first optic -- the spherical-crystal spectrometer
        if (s->split == 0 && (s->known & KN_SPECTROMETER) == KN_SPECTROMETER) {
            *smem = block_smem_bytes<0, KN_SPECTROMETER>();
            return k_trace<0, 0, KN_SPECTROMETER>;
        }
        return trace_kernel_ft<0>(s->split);
    }
    if (s->features == FT_MID) {
        *smem = block_smem_bytes<FT_MID>();
        return trace_kernel_ft<FT_MID>(s->split);
    }
    if (s->features == FT_MOSAICLEAN) {
        *smem = block_smem_bytes<FT_MOSAICLEAN>();
        return s->split == 0 ? k_trace<FT_MOSAICLEAN, 0, 0> : k_trace<FT_MOSAICLEAN, -1, 0>;
    }
    if (s->features == FT_SRCLEAN) {
        *smem = block_smem_bytes<FT_SRCLEAN>();
        return s->split == 0 ? k_trace<FT_SRCLEAN, 0, 0> : k_trace<FT_SRCLEAN, -1, 0>;
    }
    if (s->features == FT_MESHLEAN) {
        *smem = block_smem_bytes<FT_MESHLEAN>();
        return trace_kernel_ft<FT_MESHLEAN>(s->split);
    }
    *smem = block_smem_bytes<FT_FULL>();
    return trace_kernel_ft<FT_FULL>(s->split);
}

static int trace_launch_config(const XrtScene *s, TraceKernel *kern, size_t *smem, int *blocks_per_sm, int *regs) {
    *kern = trace_kernel(s, smem);
    CU(cudaFuncSetAttribute(*kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)*smem));
    if (regs) {
        cudaFuncAttributes fa;
        CU(cudaFuncGetAttributes(&fa, *kern));
        *regs = fa.numRegs;
    }
    CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(blocks_per_sm, *kern, kBlock, *smem));
    if (*blocks_per_sm < 1) return fail(XRT_ECUDA, "fused kernel does not fit on an SM (%zu B shared memory)", *smem);
    return XRT_OK;
}

extern "C" int xrt_launch_info(XrtScene *s, int32_t *grid, int32_t *block, int32_t *regs, int32_t *blocks_per_sm) {
    if (!s) return fail(XRT_EINVAL, "null scene");
    TraceKernel kern;
    size_t smem;
    int bps = 0, r = 0;
    int rc = trace_launch_config(s, &kern, &smem, &bps, &r);
    if (rc != XRT_OK) return rc;
    if (grid) *grid = s->sm_count * bps;
    if (block) *block = kBlock;
    if (regs) *regs = r;
    if (blocks_per_sm) *blocks_per_sm = bps;
    return XRT_OK;
}

extern "C" int xrt_trace(XrtScene *s, uint64_t seed, uint64_t stream_id, uint64_t ray_begin, uint64_t ray_count,
                         const XrtOutputs *out, void *stream) {
    if (!s || !out) return fail(XRT_EINVAL, "null argument");
    if (ray_count == 0) return XRT_OK;
    if (s->dev.n_optics < 1) return fail(XRT_EINVAL, "a scene needs at least one optic");
    if (s->dev.source.kind == XRT_SRC_BUNDLES && !s->dev.source.bundles)
        return fail(XRT_EINVAL, "plasma scene without a bundle table: call xrt_scene_set_bundles first");
    if (s->dev.source.kind == XRT_SRC_BUNDLES && s->dev.source.wave == XRT_WAVE_TABLE && !s->dev.source.bundle_cdf)
        return fail(XRT_EINVAL, "plasma scene with a natural linewidth: call xrt_scene_set_bundle_tables first");
    TraceKernel kern;
    size_t smem;
    int bps = 0;
    int rc = trace_launch_config(s, &kern, &smem, &bps, nullptr);
    if (rc != XRT_OK) return rc;
    // one resident wave of blocks, each warp strides over the ray ids
    if (const char *lim = getenv("XRT_BLOCKS_PER_SM")) {      // measurement knob: fewer resident blocks
        int v = atoi(lim);
        if (v >= 1 && v < bps) bps = v;
    }
    uint64_t want = (ray_count + kBlock - 1) / kBlock;
    uint64_t cap = (uint64_t)s->sm_count * (uint64_t)bps;
    int grid = (int)(want < cap ? want : cap);
    PhiloxKeys pk;
    philox_round_keys(seed, stream_id, pk);
    kern<<<grid, kBlock, smem, (cudaStream_t)stream>>>(s->dev, pk, stream_id, ray_begin, ray_count, *out, s->split,
                                                      s->lazy_wavelength | (s->need_wavelength << 1) | (s->defer_wavelength << 2));
    CU(cudaGetLastError());
    return XRT_OK;
}

template <int MODE>
static int launch_record(XrtScene *s, uint64_t seed, uint64_t stream_id, const uint64_t *ids, uint64_t ray_begin,
                         uint64_t n, const XrtRaysIn &in, const XrtInject &inj, const XrtOutputs &out,
                         const XrtHistory &hist, void *stream) {
    if (MODE == REC_PHILOX && s->dev.source.kind == XRT_SRC_BUNDLES && !s->dev.source.bundles)
        return fail(XRT_EINVAL, "plasma scene without a bundle table: call xrt_scene_set_bundles first");
    if (MODE == REC_PHILOX && s->dev.source.kind == XRT_SRC_BUNDLES && s->dev.source.wave == XRT_WAVE_TABLE && !s->dev.source.bundle_cdf)
        return fail(XRT_EINVAL, "plasma scene with a natural linewidth: call xrt_scene_set_bundle_tables first");
    if (hist.rays || hist.mask) {
        if (hist.capacity < n) return fail(XRT_EINVAL, "history capacity %llu < %llu rays",
                                           (unsigned long long)hist.capacity, (unsigned long long)n);
    }
    uint64_t want = (n + kBlock - 1) / kBlock;
    uint64_t cap = (uint64_t)s->sm_count * 8;
    int grid = (int)(want < cap ? want : cap);
    cudaStream_t st = (cudaStream_t)stream;
    PhiloxKeys pk;
    philox_round_keys(seed, stream_id, pk);
    if (s->features == 0 && MODE == REC_PHILOX && s->split == 0 && (s->known & KN_SPECTROMETER) == KN_SPECTROMETER)
        k_record<0, MODE, (MODE == REC_PHILOX ? (uint32_t)KN_SPECTROMETER : 0u)><<<grid, kBlock, 0, st>>>(
            s->dev, pk, stream_id, ids, ray_begin, n, in, inj, out, hist, s->split);
    else if (s->features == 0)
        k_record<0, MODE><<<grid, kBlock, 0, st>>>(s->dev, pk, stream_id, ids, ray_begin, n, in, inj, out, hist, s->split);
    else if (s->features == FT_MID)
        k_record<FT_MID, MODE><<<grid, kBlock, 0, st>>>(s->dev, pk, stream_id, ids, ray_begin, n, in, inj, out, hist, s->split);
    else if (s->features == FT_MOSAICLEAN)
        k_record<FT_MOSAICLEAN, MODE><<<grid, kBlock, 0, st>>>(s->dev, pk, stream_id, ids, ray_begin, n, in, inj, out, hist, s->split);
    else if (s->features == FT_SRCLEAN)
        k_record<FT_SRCLEAN, MODE><<<grid, kBlock, 0, st>>>(s->dev, pk, stream_id, ids, ray_begin, n, in, inj, out, hist, s->split);
    else if (s->features == FT_MESHLEAN)
        k_record<FT_MESHLEAN, MODE><<<grid, kBlock, 0, st>>>(s->dev, pk, stream_id, ids, ray_begin, n, in, inj, out, hist, s->split);
    else
        k_record<FT_FULL, MODE><<<grid, kBlock, 0, st>>>(s->dev, pk, stream_id, ids, ray_begin, n, in, inj, out, hist, s->split);
    CU(cudaGetLastError());
    return XRT_OK;
}

extern "C" int xrt_trace_history(XrtScene *s, uint64_t seed, uint64_t stream_id, const uint64_t *ids,
                                 uint64_t ray_begin, uint64_t n, const XrtHistory *hist, void *stream) {
    if (!s || !hist) return fail(XRT_EINVAL, "null argument");
    if (n == 0) return XRT_OK;
    XrtRaysIn in = {};
    XrtInject inj = {};
    XrtOutputs out = {};
    return launch_record<REC_PHILOX>(s, seed, stream_id, ids, ray_begin, n, in, inj, out, *hist, stream);
}

extern "C" int xrt_trace_injected(XrtScene *s, const XrtRaysIn *rays, const XrtInject *draws, uint64_t n,
                                  const XrtOutputs *out, const XrtHistory *hist, void *stream) {
    if (!s || !rays) return fail(XRT_EINVAL, "null argument");
    if (n == 0) return XRT_OK;
    if (!rays->origin || !rays->direction || !rays->wavelength || !rays->mask)
        return fail(XRT_EINVAL, "incomplete ray input");
    XrtInject inj = {};
    if (draws) inj = *draws;
    for (int k = 0; k < s->dev.n_optics; ++k) {
        const XrtOpticDesc &op = s->dev.optics[k];
        const bool bragg = (op.flags & XRT_F_CHECK_BRAGG) != 0;
        if (op.interact == XRT_INTERACT_CRYSTAL && bragg && !inj.u[k])
            return fail(XRT_EINVAL, "optic %d needs injected uniforms", k);
        if (op.interact == XRT_INTERACT_MOSAIC && (!inj.xy[k] || (bragg && !inj.u[k])))
            return fail(XRT_EINVAL, "optic %d needs injected mosaic draws", k);
    }
    XrtOutputs o = {};
    if (out) o = *out;
    XrtHistory h = {};
    if (hist) h = *hist;
    return launch_record<REC_INJECT>(s, 0, 0, nullptr, 0, n, *rays, inj, o, h, stream);
}

template <int MODE>
static int launch_source(XrtScene *s, uint64_t seed, uint64_t stream_id, uint64_t ray_begin, uint64_t n,
                         const XrtSourceInject &sinj, const XrtHistory *hist, void *stream) {
    if (!s || !hist) return fail(XRT_EINVAL, "null argument");
    if (n == 0) return XRT_OK;
    if (hist->capacity < n) return fail(XRT_EINVAL, "history capacity too small");
    if (MODE == REC_PHILOX && s->dev.source.kind == XRT_SRC_BUNDLES && !s->dev.source.bundles)
        return fail(XRT_EINVAL, "plasma scene without a bundle table: call xrt_scene_set_bundles first");
    if (MODE == REC_PHILOX && s->dev.source.kind == XRT_SRC_BUNDLES && s->dev.source.wave == XRT_WAVE_TABLE && !s->dev.source.bundle_cdf)
        return fail(XRT_EINVAL, "plasma scene with a natural linewidth: call xrt_scene_set_bundle_tables first");
    uint64_t want = (n + kBlock - 1) / kBlock;
    uint64_t cap = (uint64_t)s->sm_count * 8;
    int grid = (int)(want < cap ? want : cap);
    PhiloxKeys pk;
    philox_round_keys(seed, stream_id, pk);
    k_source<MODE><<<grid, kBlock, 0, (cudaStream_t)stream>>>(s->dev, pk, stream_id, ray_begin, n, sinj, *hist);
    CU(cudaGetLastError());
    return XRT_OK;
}

extern "C" int xrt_source_injected(XrtScene *s, const XrtSourceInject *draws, uint64_t n, const XrtHistory *hist,
                                   void *stream) {
    if (!draws) return fail(XRT_EINVAL, "null argument");
    if (s && s->dev.source.cone == XRT_CONE_ISOTROPIC_XY)
        return fail(XRT_EUNSUPPORTED, "injected draws: isotropic_xy has a variable draw count");
    return launch_source<REC_INJECT>(s, 0, 0, 0, n, *draws, hist, stream);
}

extern "C" int xrt_source_generate(XrtScene *s, uint64_t seed, uint64_t stream_id, uint64_t ray_begin, uint64_t n,
                                   const XrtHistory *hist, void *stream) {
    XrtSourceInject none = {};
    return launch_source<REC_PHILOX>(s, seed, stream_id, ray_begin, n, none, hist, stream);
}

// Bracket table for the ray id -> bundle search (XrtSourceDesc.bundle_hint), on the legacy default stream.
static int build_bundle_hint(XrtScene *s, uint64_t n_rays) {
    XrtSourceDesc &src = s->dev.source;
    if (s->bundle_hint) {
        cudaFreeAsync(s->bundle_hint, (cudaStream_t)0);
        s->bundle_hint = nullptr;
    }
    src.bundle_hint = nullptr;
    src.bundle_hint_shift = 0;
    if (n_rays == 0 || src.n_bundles < 64 || src.n_bundles > 0xffffffffull) return XRT_OK;
    int shift = 10;
    while ((n_rays >> shift) > (1ull << 22)) ++shift;        // at most 4 Mi entries (16 MB)
    const uint64_t n_buckets = ((n_rays - 1) >> shift) + 1;
    void *p = nullptr;
    CU(cudaMallocAsync(&p, (n_buckets + 1) * sizeof(uint32_t), (cudaStream_t)0));
    s->bundle_hint = p;
    uint64_t want = (n_buckets + 256) / 256;
    int grid = (int)(want < 4096 ? want : 4096);
    k_bundle_hint<<<grid, 256, 0, (cudaStream_t)0>>>(src.bundle_end, src.n_bundles, shift, n_buckets, (uint32_t *)p);
    CU(cudaGetLastError());
    src.bundle_hint = (const uint32_t *)p;
    src.bundle_hint_shift = shift;
    return XRT_OK;
}

extern "C" int xrt_scene_set_bundles(XrtScene *s, const XrtBundle *table_dev, const uint64_t *end_dev, uint64_t n_bundles,
                                     uint64_t n_rays) {
    if (!s) return fail(XRT_EINVAL, "null scene");
    if (s->dev.source.kind != XRT_SRC_BUNDLES) return fail(XRT_EINVAL, "the scene's source is not a plasma");
    if (!table_dev || !end_dev || n_bundles == 0) return fail(XRT_EINVAL, "empty bundle table");
    s->dev.source.bundles = table_dev;
    s->dev.source.bundle_end = end_dev;
    s->dev.source.n_bundles = n_bundles;
    int prev = 0;
    CU(cudaGetDevice(&prev));
    if (prev != s->device) CU(cudaSetDevice(s->device));
    const int rc = build_bundle_hint(s, n_rays);
    if (prev != s->device) cudaSetDevice(prev);
    return rc;
}

extern "C" int xrt_scene_set_bundle_tables(XrtScene *s, const double *x_dev, const double *cdf_dev, int32_t n_table) {
    if (!s) return fail(XRT_EINVAL, "null scene");
    if (s->dev.source.kind != XRT_SRC_BUNDLES || s->dev.source.wave != XRT_WAVE_TABLE)
        return fail(XRT_EINVAL, "the scene's source is not a plasma with a tabulated line shape");
    if (!x_dev || !cdf_dev || n_table != s->dev.source.n_table) return fail(XRT_EINVAL, "bad bundle wavelength tables");
    s->dev.source.bundle_x = x_dev;
    s->dev.source.bundle_cdf = cdf_dev;
    return XRT_OK;
}

extern "C" int xrt_bundle_voigt_tables(const XrtBundle *table_dev, const int64_t *counts_dev, uint64_t n_bundles,
                                       double gamma, int32_t n_table, double *x_dev, double *cdf_dev, void *stream) {
    if (!table_dev || !counts_dev || !x_dev || !cdf_dev) return fail(XRT_EINVAL, "null argument");
    if (n_table < 2 || n_table > 8 * kVoigtBlock) return fail(XRT_EINVAL, "n_table = %d (2 .. %d)", n_table, 8 * kVoigtBlock);
    if (!(gamma > 0.0)) return fail(XRT_EINVAL, "gamma must be > 0");
    if (n_bundles == 0) return XRT_OK;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
        cudaGetLastError();
        return fail(XRT_ECUDA, "no CUDA device: libxrt has no CPU path");
    }
    int grid = (int)(n_bundles < 148u * 8u ? n_bundles : 148u * 8u);
    k_voigt_tables<<<grid, kVoigtBlock, 0, (cudaStream_t)stream>>>(table_dev, (const long long *)counts_dev, n_bundles, gamma,
                                                                  n_table, x_dev, cdf_dev);
    CU(cudaGetLastError());
    return XRT_OK;
}

extern "C" int xrt_bundles_generate(const XrtPlasmaDesc *desc, uint64_t seed, uint64_t stream_id, uint64_t n_bundles,
                                    XrtBundle *table_dev, double *intensity_dev, int64_t *counts_dev, void *stream) {
    if (!desc || !table_dev || !intensity_dev || !counts_dev) return fail(XRT_EINVAL, "null argument");
    if (n_bundles == 0) return XRT_OK;
    if (desc->kind < XRT_PLASMA_GENERIC || desc->kind > XRT_PLASMA_DATAFILE) return fail(XRT_EINVAL, "plasma kind = %d", desc->kind);
    if (desc->n_sightlines < 0 || desc->n_sightlines > XRT_MAX_SIGHTLINES) return fail(XRT_EINVAL, "n_sightlines = %d", desc->n_sightlines);
    if (desc->kind == XRT_PLASMA_DATAFILE &&
        (desc->n_profile_t < 2 || desc->n_profile_e < 2 || !desc->profile_t_rho || !desc->profile_t_val ||
         !desc->profile_e_rho || !desc->profile_e_val))
        return fail(XRT_EINVAL, "datafile plasma without profile tables");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
        cudaGetLastError();
        return fail(XRT_ECUDA, "no CUDA device: libxrt has no CPU path");
    }
    uint64_t want = (n_bundles + 255) / 256;
    int grid = (int)(want < 65535 ? want : 65535);
    k_bundles<<<grid, 256, 0, (cudaStream_t)stream>>>(*desc, seed, stream_id, n_bundles, table_dev, intensity_dev,
                                                     (long long *)counts_dev);
    CU(cudaGetLastError());
    return XRT_OK;
}

extern "C" int xrt_fp64_burn(uint64_t iters, double *out_dev, double *flops, void *stream) {
    if (!out_dev) return fail(XRT_EINVAL, "null argument");
    int dev = 0, sms = 0, bps = 0;
    CU(cudaGetDevice(&dev));
    CU(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&bps, k_burn, kBlock, 0));
    int grid = sms * bps;
    k_burn<<<grid, kBlock, 0, (cudaStream_t)stream>>>(iters, out_dev);
    CU(cudaGetLastError());
    if (flops) *flops = 2.0 * (double)kBurnChains * (double)iters * (double)grid * (double)kBlock;
    return XRT_OK;
}
