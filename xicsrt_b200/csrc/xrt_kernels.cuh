// xrt_kernels.cuh -- the kernels of libxrt.so (templates; instantiated per scene variant in v_*.cu).
//
//   k_cull32<SRC,HIST>   FP32 broad phase of the Bragg pre-test for scenes whose first optic is a spherical Bragg
//                        crystal: every ray of the launch is generated in single precision from its Philox blocks
//                        and tested with the conservative chord inequality; the rays it cannot reject (~20 % for a
//                        spectrometer, a few % for an extended plasma) leave as 4-byte id offsets in a region-
//                        partitioned list.  Integer / FP32 / MUFU only, ~40 registers: runs at full occupancy.
//   k_trace<FT,..>       fused generate -> optic train -> bin in FP64 for the ids of that list (or for every id of
//                        the launch when the broad phase does not apply); ray state in registers from the source to
//                        the detector; the only global traffic is the id list, the per-element survivor counters
//                        (one atomic per block), the pixel counters of surviving rays (warp-aggregated atomics) and
//                        the optional found / lost id lists (ballot + prefix-sum compaction).
//   k_record<FT,..>      same ray code, but every element's ray state is stored as struct-of-arrays history
//                        (coalesced planes); rays come either from Philox by id (history of selected rays) or from
//                        caller memory with injected draws (parity entry).
//   k_source<..>         source only (history element 0).
//   k_burn               dependent DFMA chains: the FP64 roofline denominator.
#pragma once
#include <cuda_runtime.h>

#include "xrt_trace.cuh"

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
#ifndef XRT_MESH_BLOCKS
#define XRT_MESH_BLOCKS 2      // resident blocks per SM of the mesh variants
#endif
#ifndef XRT_MOSAIC_BLOCKS
#define XRT_MOSAIC_BLOCKS 2    // resident blocks per SM of the lean mosaic variant
#endif
#ifndef XRT_MIN_SCAN
#define XRT_MIN_SCAN 8         // mosaic scan stage: lanes that must still be scanning for another scan iteration
#endif
constexpr int kBlock = XRT_BLOCK;
constexpr unsigned kFull = 0xffffffffu;

// ---------------------------------------------------------------------------
// id source of a k_trace launch: the ray ids [ray_begin, ray_begin + ray_count) are cut into regions of `cap`
// consecutive ids.  Sequence mode (ids == nullptr): every id of a region is traced (cap = 32: one warp pass per
// region).  List mode: region r holds counts[r] surviving id offsets of its range at ids[r * cap ...], written by
// k_cull32.  Warp w takes regions w, w + n_warps, ... in both modes: no atomics, and the order in which a warp
// meets its rays does not depend on scheduling.
// tag_bits > 0: a list entry is (id offset << tag_bits) | tag; the mosaic broad phase passes the first crystallite
// layer it could not reject this way.
struct IdList {
    const uint32_t *ids;
    const uint32_t *counts;
    uint32_t n_regions, cap;
    uint32_t tag_bits;
};

// U groups of <= 32 ids are fetched one pass ahead of their use, so that the list loads (L2 / DRAM latency) overlap the
// arithmetic of the pass before.
template <int U>
struct IdCursor {
    const uint32_t *ids, *counts;
    uint64_t base, n_rays;
    uint32_t reg, n_regions, cap, pos, cnt, step, tag_bits;
    uint64_t pbase[U];    // pending ids of this lane: region base + poff (the sum is formed when the id is handed out, so
    uint32_t poff[U];     // that the list load stays in flight until then)
    bool pvalid[U];
    bool pany;            // warp-uniform: the pending groups hold at least one id

    __device__ __forceinline__ void load() {
        cnt = 0;
        if (reg < n_regions) {
            if (ids) cnt = __ldg(counts + reg);
            else {
                const uint64_t left = n_rays - (uint64_t)reg * cap;
                cnt = left < (uint64_t)cap ? (uint32_t)left : cap;
            }
        }
    }
    // warp-uniform: move to the next non-empty region if the current one is used up
    __device__ __forceinline__ bool advance() {
        while (pos >= cnt) {
            if (reg >= n_regions) return false;
            reg += step;
            pos = 0;
            load();
        }
        return true;
    }
    __device__ __forceinline__ void fetch(unsigned lane) {
        pany = false;
#pragma unroll
        for (int j = 0; j < U; ++j) {
            pbase[j] = base;
            poff[j] = 0u;
            pvalid[j] = false;
            if (!advance()) continue;
            pany = true;
            const uint32_t p = pos + lane;
            pvalid[j] = p < cnt;
            const uint64_t off = (uint64_t)reg * cap + p;
            if (ids) { if (pvalid[j]) poff[j] = __ldg(ids + off); }
            else pbase[j] = base + off;
            pos += 32;
        }
    }
    __device__ __forceinline__ void init(const IdList &L, uint64_t ray_begin, uint64_t ray_count, uint32_t warp_global,
                                         uint32_t n_warps, unsigned lane) {
        ids = L.ids; counts = L.counts; n_regions = L.n_regions; cap = L.cap; tag_bits = L.tag_bits;
        base = ray_begin; n_rays = ray_count;
        reg = warp_global; step = n_warps; pos = 0;
        load();
        fetch(lane);
    }
    __device__ __forceinline__ bool more() const { return pany; }
    // hand out the pending groups and start fetching the ones after them
    __device__ __forceinline__ void take(unsigned lane, uint64_t (&id)[U], bool (&valid)[U]) {
#pragma unroll
        for (int j = 0; j < U; ++j) { id[j] = pbase[j] + poff[j]; valid[j] = pvalid[j]; }
        fetch(lane);
    }
    // tagged entries: id and tag
    __device__ __forceinline__ void take(unsigned lane, uint64_t (&id)[U], bool (&valid)[U], uint32_t (&tag)[U]) {
#pragma unroll
        for (int j = 0; j < U; ++j) {
            id[j] = pbase[j] + (poff[j] >> tag_bits);
            tag[j] = poff[j] & ((1u << tag_bits) - 1u);
            valid[j] = pvalid[j];
        }
        fetch(lane);
    }
};

// ---------------------------------------------------------------------------
// fused kernel
//
// Each warp works through its share of the ray ids in stages and re-packs the survivors between them in per-warp
// shared-memory queues (ballot + popc prefix sums), so that every stage runs with (nearly) all 32 lanes busy
// although ~48 % of the rays of a typical spectrometer miss the crystal and ~98 % of the rest fail the Bragg test:
//
//   stage A  ids          source origin + direction, optics before the split optic, geometry (intersect + bounds)
//                         of the split optic, both levels of the FP64 Bragg pre-test where it applies  -> queue 1
//   stage B  queue 1      wavelength (drawn here when it does not depend on the source direction: Philox is counter
//                         based), interaction of the split optic (Bragg / mosaic / mirror), its image -> queue 2
//   stage C  queue 2      the remaining optics, images, found list
//
// Variants:
//   spectrometer (KN)  AB  FP64 direction from the cone block, sphere chord, first level of the pre-test,
//                          intersection point, bounds, second level (rocking uniform)      -> queue b
//                      B2  exact wavelength, Bragg angle, rocking curve, reflection        -> queue 2, then C
//   mesh split optic   A1 coarse mesh for every ray -> queue a; A2 refinement + interpolation -> queue 1
//   mosaic split optic see stage S / E below
//
// The split optic is the first crystal of the train (0 if there is none).  SPLIT >= 0 makes its index a
// compile-time constant, so its parameters are fetched from the constant bank at fixed offsets (uniform loads)
// instead of register-indexed ones.  Warps never wait for one another: only __syncwarp and warp-uniform counters.
// HIST = false compiles the found / lost list emission out (history-off launches: bench.py's timed launch).

constexpr int kQ1Cap = 64;     // stage A pushes <= 32 per pass, stage B pops 32 when >= 32 are queued
constexpr int kQaCap = 64;     // mesh variants: rays that hit the coarse mesh (id, coarse hit point), between the halves of stage A
constexpr int kQaPlanes = 4;
#ifndef XRT_UNROLL
#define XRT_UNROLL 2
#endif
constexpr int kUnroll = XRT_UNROLL;   // spectrometer variant: groups of 32 rays per stage-AB pass (independent chains)
constexpr int kQbPlanes = 5;          // id, direction, distance
constexpr int kQbCap = 32 * (kUnroll + 1);   // spectrometer variant: rays inside the bounds that passed both pre-test levels
constexpr int kQ2Cap = 64;     // stage B pushes <= 32 per pass, stage C pops 32 when >= 32 are queued
constexpr int kQ2Planes = 8;   // id, origin, direction, wavelength

template <uint32_t FT> __host__ __device__ constexpr int q1_planes() {
    // id, intersection point, direction [, wavelength when it can be eager] [, normal for mesh shapes]
    return 7 + (FT != 0 ? 1 : 0) + ((FT & FT_MESH) != 0 ? 3 : 0);
}
template <uint32_t FT, uint32_t KN = 0> __host__ __device__ constexpr int q1_doubles() {
    return ((KN & KN_SPECTROMETER) == KN_SPECTROMETER) ? kQbPlanes * kQbCap
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
template <bool HIST>
__device__ __forceinline__ void emit_found(const XrtOutputs &out, unsigned lane, unsigned lt_mask, bool found, uint64_t id) {
    if constexpr (!HIST) return;
    if (out.found_bits && found) {
        const uint64_t b = id - out.bits_begin;
        atomicOr(out.found_bits + (b >> 5), 1u << (b & 31u));
    }
    if (!out.found_count) return;
    unsigned m = __ballot_sync(kFull, found);
    if (!m) return;
    unsigned long long off = 0;
    if (lane == 0) off = atomicAdd((unsigned long long *)out.found_count, (unsigned long long)__popc(m));
    off = __shfl_sync(kFull, off, 0);
    if (found && out.found_ids) {
        unsigned long long slot = off + __popc(m & lt_mask);
        if (slot < out.found_capacity) out.found_ids[slot] = id;
    }
}

// lost sample: a lost ray is kept when its 64-bit Philox key is below the threshold
template <bool HIST>
__device__ __forceinline__ void emit_lost(const XrtOutputs &out, unsigned lane, unsigned lt_mask, const PhiloxDraws &dr,
                                          bool lost, uint64_t id) {
    if constexpr (!HIST) return;
    if (!out.lost_count && !out.lost_bits) return;
    bool keep = false;
    uint64_t key = 0;
    if (lost) {
        key = dr.lost_key();
        keep = key < out.lost_threshold;
    }
    if (out.lost_bits && keep) {
        const uint64_t b = id - out.bits_begin;
        atomicOr(out.lost_bits + (b >> 5), 1u << (b & 31u));
    }
    if (!out.lost_count) return;
    unsigned m = __ballot_sync(kFull, keep);
    if (!m) return;
    unsigned long long off = 0;
    if (lane == 0) off = atomicAdd((unsigned long long *)out.lost_count, (unsigned long long)__popc(m));
    off = __shfl_sync(kFull, off, 0);
    if (keep && out.lost_ids) {
        unsigned long long slot = off + __popc(m & lt_mask);
        if (slot < out.lost_capacity) {
            out.lost_ids[slot] = id;
            if (out.lost_keys) out.lost_keys[slot] = key;
        }
    }
}

// ---- stage C: the optics after the split optic, for the `cnt` rays in queue 2
template <uint32_t FT, bool HIST>
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
            trace_optic_any<FT, PhiloxDraws>(op, k, dr, r);
            if (r.alive && (op.flags & XRT_F_IMAGE) && out.images) add_pixel(out, op, r, c.lt_mask);
        }
        count_alive(c, k + 1, r.alive);
    }
    emit_found<HIST>(out, c.lane, c.lt_mask, r.alive, id);
    emit_lost<HIST>(out, c.lane, c.lt_mask, dr, active && !r.alive, id);
}

// ---- stage B: interaction of the split optic for `cnt` rays popped from queue 1
template <uint32_t FT, uint32_t KN, bool HIST>
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
    emit_lost<HIST>(out, c.lane, c.lt_mask, dr, active && !r.alive, id);

    if (split + 1 >= sc.n_optics) {        // the split optic is the last one: survivors are found
        emit_found<HIST>(out, c.lane, c.lt_mask, r.alive, id);
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

// ---- stage S: mosaic crystal as split optic (_InteractMosaicCrystal.py:53-139), for `cnt` rays popped from queue 1.
//
// The reference walks up to mosaic_depth layers of crystallites per ray; in each layer it draws a crystallite normal
// (two normals) and a uniform u and reflects the ray if exp(-dtheta^2 / 2 sigma^2) reflectivity >= u.  Only ~3 % of
// the (ray, layer) pairs of a HOPG-like crystal pass, so the FP64 evaluation of a layer (Box-Muller, two normalisations,
// the Bragg angle, the exponential: ~400 instructions) is preceded by an FP32 pre-test of the same inequality, the one
// of the Bragg pre-test with the layer's own uniform:  reflected  =>  dtheta^2 <= 2 sigma^2 ln(reflectivity / u), and
// |sin(theta_B) - sin(theta_i)| <= |dtheta| cos(min angle).  sin(theta_i) = |D.n_m| / |D| with the crystallite normal
// n_m = (x r_0 + y r_1 + n) / sqrt(x^2 + y^2 + 1) is three precomputed dot products and one rsqrt per layer; (x, y, u)
// come from the layer's Philox block (the block the exact path reads).  Everything is within 5e-7 of the FP64 value
// (margin cull_err = 2e-6, tests/test_host_logic.py), so a layer the pre-test rejects would be rejected by the exact
// test too and results do not change (XRT_NO_CULL runs the plain loop).
//
// Lanes scan their own layers independently; a lane that meets a layer it cannot reject waits as a candidate.  When
// fewer than kMinScan lanes are still scanning, the candidates evaluate their layer exactly (FP64, the code of
// optic_interact): a pass reflects the ray, a fail resumes the scan at the next layer.  When neither is left to do, the
// unfinished rays go back to queue 1 with their layer index (in the top byte of the id word) and their wavelength, and
// are re-packed with new rays.  While the queues drain at the end of the launch the batch runs to completion.
constexpr int kMinScan = XRT_MIN_SCAN;
constexpr uint64_t kIdMask = (1ull << 56) - 1ull;
// queue word of a mosaic ray: id (56 bits) | layer (7 bits) | resumed (1 bit: the ray has been in stage S before, its
// wavelength travels with it)
constexpr uint64_t kResumed = 1ull << 63;

// FP32 pre-test of one crystallite layer (b = the layer's Philox block): true = the layer provably does not reflect
// the ray.  dr0, dr1, dn = D.r_0, D.r_1, D.n for the unit direction D and the basis (r_0, r_1, n) of mosaic_normal;
// sB = sin(theta_B); err = margin for everything that is not exact here.
struct MosaicPre {
    float s32, t2, two_sigma2, lg_refl;
    bool gauss;
};
__device__ __forceinline__ bool mosaic_pretest(const uint4 b, const MosaicPre &M, float dr0, float dr1, float dn, float sB, float err) {
    const uint32_t na = ~b.x;
    const float omu = fmaf((float)na, 2.3283064365386963e-10f, 1.1641532182693481e-10f);   // 1 - u1
    const float rr = sqrt_approx(-1.3862943611198906f * lg2_approx(omu));                 // sqrt(-2 ln(1 - u1))
    const float ang = 6.283185307179586f * (__uint_as_float(0x3f800000u | ((b.y & 0xffffffu) >> 1)) - 1.5f);
    const float x = -M.s32 * rr * __cosf(ang), y = -M.s32 * rr * __sinf(ang);
    const float t = fmaf(x, dr0, fmaf(y, dr1, dn));
    const float sI = fabsf(t) * rsqrt_approx(fmaf(x, x, fmaf(y, y, 1.0f)));
    const float gap = fabsf(sB - sI);
    const float diff = gap - err;
    const float c2 = fmaf(2.0f, gap, fmaf(-sI, sI, 1.0f));
    bool rej = (diff > 0.0f) & (diff * diff > M.t2 * c2);
    if (M.gauss) {
        const float u = __uint_as_float(0x3f800000u | (b.z >> 9)) - 1.0f;          // top 23 bits: u32 <= u
        const float lim = 0.6931471805599453f * (M.lg_refl - lg2_approx(u));       // >= ln(reflectivity / u)
        const float bound = fmaf(fabsf(lim), 1e-3f, lim + 1e-3f) * M.two_sigma2;
        rej |= (diff > 0.0f) & (diff * diff > bound * c2) & (lim == lim);
    }
    // 1 - u1 below 2^-16 (|z| > 4.7): its 32-bit truncation is not precise enough for the radius, the layer is
    // decided exactly (1.5e-5 of the layers)
    return rej & (na >= 65536u);
}

template <uint32_t FT, uint32_t KN, bool HIST>
__device__ __forceinline__ void stage_mosaic(const XrtSceneDesc &sc, const XrtOpticDesc &ops, const XrtOutputs &out,
                                             const WarpCtx &c, int split, bool lazy, bool need_wave, bool defer, const PhiloxKeys &pk,
                                             uint64_t stream_id, double *q1, int &n1, int cnt, bool drain, double *q2, int &n2,
                                             unsigned &n_split) {
    constexpr int P = kQ1Cap;
    const bool active = (int)c.lane < cnt;
    Ray r;
    r.alive = false;
    r.o = r.d = v3(0.0, 0.0, 1.0);
    r.w = 1.0;
    uint64_t id = 0;
    int layer = 0;
    bool fresh = false;
    if (active) {
        const double *p = q1 + n1 + c.lane;
        const uint64_t word = (uint64_t)__double_as_longlong(p[0]);
        id = word & kIdMask;
        layer = (int)((word >> 56) & 0x7fu);
        fresh = (word & kResumed) == 0;
        r.o = v3(p[1 * P], p[2 * P], p[3 * P]);
        r.d = v3(p[4 * P], p[5 * P], p[6 * P]);
        r.w = p[7 * P];
    }
    __syncwarp();
    PhiloxDraws dr;
    dr.init(pk, stream_id, id, split);
    const uint32_t flags = flags_of<KN>(ops);
    const V3 n = analytic_normal<FT, KN>(ops, r.o);
    bool scanning = active;
    if (active && fresh) {
        // first visit (at the layer the broad phase k_mosaic32 found, else layer 0): the wavelength (lazy / deferred, as
        // stage B) and the optional prefilter on the nominal normal
        if (lazy && need_wave) {
            SrcLocal L;
            source_local<0, KN>(sc.source, id, L);
            r.w = generate_wavelength<PhiloxDraws, KN, false>(sc.source, L, dr, r.d);
        }
        if (defer) {
            SrcLocal L;
            source_local<FT, KN>(sc.source, id, L);
            const double w0 = sc.source.wave_par[0] + L.wave_sigma * dr.wave_z();
            r.w = (r.w != 1.0) ? w0 * r.w : w0;
        }
        if (flags & XRT_F_MOSAIC_CUTOFF) scanning = fabs(bragg_dtheta(ops, r.d, r.w, n)) < ops.mosaic_angle_cut;
    }
    // single-precision constants of this ray's scan: D.r_0, D.r_1, D.n over |D| (r_0, r_1: basis of mosaic_normal)
    float dr0, dr1, dn, sB;
    {
        V3 r0 = unit(v3(0.0 + n.y, n.z - n.x, -n.y + 0.0));
        V3 r1 = unit(cross(n, r0));
        const double il = fast_rsqrt(dot(r.d, r.d));
        dr0 = (float)(dot(r.d, r0) * il);
        dr1 = (float)(dot(r.d, r1) * il);
        dn = (float)(dot(r.d, n) * il);
        sB = (float)(r.w * ops.inv_two_d);
    }
    const float err = (float)ops.mosaic_err;
    MosaicPre MP;
    MP.s32 = (float)ops.mosaic_sin_sigma;
    MP.t2 = (float)ops.mosaic_t2;
    MP.gauss = ops.rocking_type != XRT_ROCK_STEP;
    MP.two_sigma2 = (float)ops.rock_two_sigma2;
    MP.lg_refl = lg2_approx((float)ops.reflectivity);
    const int depth = ops.mosaic_depth;
    const int min_scan = drain ? 1 : kMinScan;
    bool cand = false, reflected = false;
    if (layer >= depth) scanning = false;

    // FP32 pre-test of one layer: true = the layer provably does not reflect the ray
    auto pretest = [&](int lay) -> bool {
        return mosaic_pretest(dr.raw(site_optic(split, lay, 1)), MP, dr0, dr1, dn, sB, err);
    };

    for (;;) {
        // ---- scan: two layers per iteration and lane (independent chains; the second is wasted only when the first
        // is a candidate, 3 % of the time)
        while ((int)__popc(__ballot_sync(kFull, scanning)) >= min_scan) {
            if (scanning) {
                const bool rej0 = pretest(layer);
                const bool rej1 = pretest(layer + 1);        // a layer index past the depth is never used
                if (!rej0) {
                    cand = true;
                    scanning = false;
                } else if (layer + 1 >= depth) {
                    layer = depth;
                    scanning = false;
                } else if (!rej1) {
                    layer += 1;
                    cand = true;
                    scanning = false;
                } else {
                    layer += 2;
                    if (layer >= depth) scanning = false;
                }
            }
        }
        if (!__ballot_sync(kFull, cand)) break;
        // ---- exact evaluation of the candidates' layers (optic_interact, one layer)
        if (cand) {
            double x, y;
            dr.mosaic_xy(split, layer, ops.mosaic_sin_sigma, x, y);
            const V3 nm = mosaic_normal(n, x, y);
            const bool pass = bragg_pass<FT, PhiloxDraws, KN, true>(ops, split, layer, dr, bragg_dtheta(ops, r.d, r.w, nm));
            cand = false;
            if (pass) {
                reflect(r, nm);
                reflected = true;
            } else {
                ++layer;
                scanning = layer < depth;
            }
        }
    }

    // ---- reflected rays leave for stage C, lost rays are done, the rest go back to queue 1
    r.alive = reflected;
    if (reflected && (flags & XRT_F_IMAGE) && out.images) add_pixel(out, ops, r, c.lt_mask);
    const unsigned mr = __ballot_sync(kFull, reflected);
    n_split += __popc(mr);
    emit_lost<HIST>(out, c.lane, c.lt_mask, dr, active && !reflected && !scanning, id);
    const unsigned ms = __ballot_sync(kFull, scanning);
    if (scanning) {
        double *p = q1 + n1 + __popc(ms & c.lt_mask);
        p[0] = __longlong_as_double((long long)(id | ((uint64_t)layer << 56) | kResumed));
        p[1 * P] = r.o.x; p[2 * P] = r.o.y; p[3 * P] = r.o.z;
        p[4 * P] = r.d.x; p[5 * P] = r.d.y; p[6 * P] = r.d.z;
        p[7 * P] = r.w;
    }
    n1 += __popc(ms);
    if (split + 1 >= sc.n_optics) {
        emit_found<HIST>(out, c.lane, c.lt_mask, reflected, id);
    } else {
        if (reflected) {
            double *p = q2 + n2 + __popc(mr & c.lt_mask);
            p[0] = __longlong_as_double((long long)id);
            p[1 * kQ2Cap] = r.o.x; p[2 * kQ2Cap] = r.o.y; p[3 * kQ2Cap] = r.o.z;
            p[4 * kQ2Cap] = r.d.x; p[5 * kQ2Cap] = r.d.y; p[6 * kQ2Cap] = r.d.z;
            p[7 * kQ2Cap] = r.w;
        }
        n2 += __popc(mr);
    }
    __syncwarp();
}

// ---- spectrometer variant, stage AB for one ray: direction from the cone block, the two lengths of the sphere
// intersection, first level of the Bragg pre-test (sin(theta_i) = |D.n| is thc / R for a ray of unit direction:
// D.(C - X) = tca - t = -thc, so the pre-test needs nothing else), intersection point and bounds (arithmetic of
// optic_geometry), second level with the ray's rocking-curve uniform (the same Philox block stage B2 reads).
// Returns true for a ray that goes on to stage B2.  A ray the pre-test rejects is lost at the crystal whether or
// not it is inside the bounds.
__device__ __forceinline__ bool spectro_stage_ab(const XrtSceneDesc &sc, const XrtOpticDesc &ops, const PhiloxDraws &dr,
                                                 int split, const double *s_sincos, bool valid, V3 &d_out, double &t_out) {
    const XrtSourceDesc &src = sc.source;
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
    const V3 o = v3(src.origin);
    const V3 Lc = v3(ops.center) - o;
    const double tca = dot(Lc, d);
    const double d2 = fma(-tca, tca, dot(Lc, Lc));
    const double r2 = ops.radius * ops.radius;
    const double thc = fast_sqrt(r2 - d2);          // NaN when d2 > r2
    const double t = tca + thc;
    d_out = d;
    t_out = t;
    bool cand = valid & (d2 >= 0.0) & (d2 <= r2);
    double gap = -1.0, c2 = 1.0;
    if (ops.cull_t2 > 0.0) cand &= !bragg_cull_sphere(src, ops, dr.wave_hi(), thc, gap, c2);
    const V3 X = v3(fma(d.x, t, o.x), fma(d.y, t, o.y), fma(d.z, t, o.z));
    const V3 Xl = to_local(ops.orient, X - v3(ops.origin));
    cand &= (fabs(Xl.x) < ops.half_size[0]) & (fabs(Xl.y) < ops.half_size[1]);
    if (ops.cull_t2 > 0.0 && ops.rocking_type != XRT_ROCK_STEP) {
        const double u = dr.bragg_u(split, 0);
        if (bragg_cull_uniform(ops, gap, c2, ops.cull_err, u)) cand = false;
    }
    return cand;
}

// Resident blocks per SM: 2 for the mesh variants (face loops and Clough-Tocher cubics keep many values live), for
// the mosaic variants (scan state + exact layer evaluation), for the spectrometer variant (two ray groups per pass =
// two independent chains) and for the lean extended-source variant (bundle lookup + focused cone basis); 3 otherwise.
template <uint32_t FT, uint32_t KN> __host__ __device__ constexpr int trace_min_blocks() {
    if ((FT & FT_MESH) != 0) return XRT_MESH_BLOCKS;
    if (FT == FT_MOSAICLEAN) return XRT_MOSAIC_BLOCKS;
    return (((FT & FT_MESH) != 0 || (FT & FT_MOSAIC) != 0 || FT == FT_SRCLEAN ||
             ((KN & KN_SPECTROMETER) == KN_SPECTROMETER && XRT_UNROLL > 1)) &&
            XRT_MIN_BLOCKS > 2) ? 2 : XRT_MIN_BLOCKS;
}

template <uint32_t FT, int SPLIT, uint32_t KN, bool HIST>
__global__ void __launch_bounds__(kBlock, (trace_min_blocks<FT, KN>()))
k_trace(const __grid_constant__ XrtSceneDesc sc, const __grid_constant__ PhiloxKeys pk, const uint64_t stream_id,
        const uint64_t ray_begin, const uint64_t ray_count, const XrtOutputs out, const IdList list, const int split_rt,
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

    constexpr bool SPECTRO = (KN & KN_SPECTROMETER) == KN_SPECTROMETER;
    constexpr int P = SPECTRO ? kQbCap : kQ1Cap;
    WarpCtx c;
    c.lane = threadIdx.x & 31u;
    c.lt_mask = (1u << c.lane) - 1u;
    c.s_cnt = s_cnt;
    const int warp = threadIdx.x >> 5;
    double *q1 = s_queue + (size_t)warp * warp_queue_doubles<FT, KN>();     // spectrometer variant: queue b
    double *q2 = q1 + q1_doubles<FT, KN>();

    const int split = (SPLIT >= 0) ? SPLIT : split_rt;
    const XrtOpticDesc &ops = sc.optics[split];
    const bool lazy = (FT == 0) ? true : ((lazy_rt & 1) != 0);
    const bool need_wave = (lazy_rt & 2) != 0;   // some optic from the split optic on reads the wavelength (Bragg test)
    const bool defer = (FT == 0) ? false : ((lazy_rt & 4) != 0);   // eager normal line: exact deviate left to stage B
    const bool count_src = list.ids == nullptr;  // list mode: k_cull32 has counted the rays out of the source

    // mesh split optic: stage the face operands every ray is tested against in shared memory.  When every ray starts
    // at the same point (point source in front of a refining mesh as first optic) the staged values are the per-face
    // constants of mesh_all_faces_point instead, and stage A1 runs its dot-product pre-selection.
    const double *staged = nullptr;
    bool mesh_staged = false, mesh_point = false;
    if constexpr ((FT & FT_MESH) != 0) {
        if constexpr (!SPECTRO) {
            mesh_staged = split == 0 && ops.shape == XRT_SHAPE_MESH && (ops.flags & XRT_F_MESH_REFINE) &&
                          !(ops.flags & XRT_F_MESH_LOSSLESS) && lazy &&
                          sc.source.kind != XRT_SRC_BUNDLES && sc.source.cone != XRT_CONE_ISOTROPIC_XY;
        }
        if (ops.shape == XRT_SHAPE_MESH) {
            const double *geom;
            const int nf = mesh_stage1_faces(ops, geom);
            if (nf <= kStageFaces) {
                double *dst = s_queue + (size_t)(kBlock / 32) * warp_queue_doubles<FT, KN>();
                const XrtSourceDesc &src = sc.source;
                mesh_point = mesh_staged && src.extent[0] == 0.0 && src.extent[1] == 0.0 && src.extent[2] == 0.0 &&
                             src.spatial == XRT_SPATIAL_UNIFORM && kPointRec * nf <= 9 * kStageFaces;
                if (mesh_point) {
                    V3 o = v3(src.origin);
                    if (optic_is_local<FT>(ops)) o = to_local(ops.orient, o - v3(ops.origin));
                    for (int i = threadIdx.x; i < nf; i += kBlock) mesh_point_constants(geom + 9 * i, o, dst + kPointRec * i);
                } else {
                    for (int i = threadIdx.x; i < 9 * nf; i += kBlock) dst[i] = __ldg(geom + i);
                }
                staged = dst;
            }
        }
        __syncthreads();
    }
    int n1 = 0, n2 = 0;     // queue fill levels, warp-uniform
    // mesh variants: queue a (coarse-mesh hits) after the planes of queue 1.  Stage A is split in two when the split
    // optic is the first optic, a refining mesh, and the wavelength is lazy (a ray is rebuilt from its id in stage A2)
    int na = 0;
    double *qa = q1 + q1_planes<FT>() * kQ1Cap;
    unsigned n_src = 0, n_split = 0;   // rays out of the source / the split optic (warp-uniform registers)

    IdCursor<SPECTRO ? kUnroll : 1> cur;
    cur.init(list, ray_begin, ray_count, blockIdx.x * (kBlock / 32) + warp, gridDim.x * (kBlock / 32), c.lane);

    // One loop, one copy of each stage: the deepest stage that has a full warp of work runs
    // first; when the ids are exhausted the queues are drained with partial warps.
    for (;;) {
        const bool more = cur.more();
        if (n2 >= 32 || (!more && n1 == 0 && na == 0 && n2 > 0)) {
            const int cnt = n2 < 32 ? n2 : 32;
            n2 -= cnt;
            stage_c<FT, HIST>(sc, out, c, split, pk, stream_id, q2, n2, cnt);
            continue;
        }
        if (n1 >= 32 || (!more && na == 0 && n1 > 0)) {
            const int cnt = n1 < 32 ? n1 : 32;
            n1 -= cnt;
            if constexpr ((FT & FT_MOSAIC) != 0 && !SPECTRO) {
                if (ops.mosaic_scan) {
                    stage_mosaic<FT, KN, HIST>(sc, ops, out, c, split, lazy, need_wave, defer, pk, stream_id, q1, n1, cnt,
                                               !more && na == 0, q2, n2, n_split);
                    continue;
                }
            }
            stage_b<FT, KN, HIST>(sc, ops, out, c, split, lazy, need_wave, defer, pk, stream_id, q1, n1, cnt, q2, n2, n_split);
            continue;
        }
        if constexpr ((FT & FT_MESH) != 0 && !SPECTRO) {
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
                    cand = optic_geometry<FT, true, KN>(ops, r, n, mesh_point ? nullptr : staged, &Xc) == HIT_INSIDE;
                }
                emit_lost<HIST>(out, c.lane, c.lt_mask, dr, active && !cand, id);
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
        if (!more) break;

        // ---- stage A
        if constexpr (SPECTRO) {
            // Straight-line code for the spectrometer (point source, concave sphere); kUnroll groups of 32 ids per
            // pass: independent dependency chains for the scheduler.
            uint64_t idv[kUnroll];
            bool validv[kUnroll], candv[kUnroll];
            V3 dv[kUnroll];
            double tv[kUnroll];
            if constexpr (SPECTRO) cur.take(c.lane, idv, validv);
#pragma unroll
            for (int j = 0; j < kUnroll; ++j) {
                PhiloxDraws dr;
                dr.init(pk, stream_id, idv[j], split);
                candv[j] = spectro_stage_ab(sc, ops, dr, split, s_sincos, validv[j], dv[j], tv[j]);
            }
#pragma unroll
            for (int j = 0; j < kUnroll; ++j) {
                if (count_src) n_src += __popc(__ballot_sync(kFull, validv[j]));
                if constexpr (HIST) {
                    if (out.lost_count || out.lost_bits) {
                        PhiloxDraws dr;
                        dr.init(pk, stream_id, idv[j], split);
                        emit_lost<HIST>(out, c.lane, c.lt_mask, dr, validv[j] && !candv[j], idv[j]);
                    }
                }
                const unsigned m = __ballot_sync(kFull, candv[j]);
                if (candv[j]) {
                    double *p = q1 + n1 + __popc(m & c.lt_mask);
                    p[0] = __longlong_as_double((long long)idv[j]);
                    p[1 * P] = dv[j].x; p[2 * P] = dv[j].y; p[3 * P] = dv[j].z;
                    p[4 * P] = tv[j];
                }
                n1 += __popc(m);
            }
            __syncwarp();
        } else {
            uint64_t id1[1];
            bool valid1[1];
            uint32_t tag1[1] = {0u};
            if constexpr (!SPECTRO) {
                if constexpr ((FT & FT_MOSAIC) != 0) cur.take(c.lane, id1, valid1, tag1);
                else cur.take(c.lane, id1, valid1);
            }
            const uint64_t id = id1[0];
            const bool valid = valid1[0];
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
            if (count_src) n_src += __popc(__ballot_sync(kFull, r.alive));
            for (int k = 0; k < split; ++k) {
                const XrtOpticDesc &op = sc.optics[k];
                if (r.alive) {
                    trace_optic_any<FT, PhiloxDraws>(op, k, dr, r);
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
                        hit = mesh_coarse_hit(ops, o, d, Xc, staged, mesh_point);
                    }
                    emit_lost<HIST>(out, c.lane, c.lt_mask, dr, valid && !hit, id);
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
            if (r.alive) cand = optic_geometry<FT, (FT & FT_MESH) != 0, KN>(ops, r, n, mesh_point ? nullptr : staged) == HIT_INSIDE;
            // Bragg pre-test (bragg_cull_general): enabled by xrt_scene_create for a spherical Bragg crystal traced in
            // global coordinates; first level with the (approximate / exact / deferred) wavelength, second level with
            // the ray's rocking-curve uniform; the survivors take the exact path in stage B
            if (ops.cull_t2 > 0.0) {
                if (cand) {
                    const int mode = defer ? WAVE_DEFERRED : (lazy ? WAVE_APPROX : WAVE_EXACT);
                    double gap, c2, err;
                    bool lost = bragg_cull_general(sc.source, ops, mode, r.w, sigma_a, dr.wave_hi(), r.o, r.d, gap, c2, err);
                    if (!lost && ops.rocking_type != XRT_ROCK_STEP && gap >= 0.0)
                        lost = bragg_cull_uniform(ops, gap, c2, err, dr.bragg_u(split, 0));
                    if (lost) {
                        cand = false;
                        r.alive = false;
                    }
                }
            }
            emit_lost<HIST>(out, c.lane, c.lt_mask, dr, valid && !cand, id);

            const unsigned m = __ballot_sync(kFull, cand);
            if (cand) {
                double *p = q1 + n1 + __popc(m & c.lt_mask);
                // mosaic variants: the layer at which stage S starts (k_mosaic32 rejected the layers before it)
                p[0] = __longlong_as_double((long long)(id | ((uint64_t)tag1[0] << 56)));
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
// FP32 broad phase (k_cull32)
//
// For a scene whose first optic is a spherical Bragg crystal (Gaussian or step rocking curve, traced in global
// coordinates) almost every ray fails the rocking-curve test by many widths.  With sB = lambda / 2d = sin(theta_B) and
// sI = |D.n| = sin(theta_i) = thc / R (thc = half chord of the ray through the sphere, |D| = 1),
//     (|sB - sI| - err)^2 > T^2 ((1 - sI^2) + 2 |sB - sI|)   ==>   |theta_B - theta_i| > T,
// T beyond the angle where the rocking curve is zero (step) or below 2^-57 (Gaussian).  Every quantity of that test
// is a function of the ray's first Philox blocks; evaluated in single precision (MUFU sqrt / sin / cos / lg2) from
// the same blocks it is within ~1e-6 of the FP64 value, and `err` carries a 2e-5 margin for it (x |C - O|^2 / R^2;
// tests/test_host_logic.py restates the arithmetic in numpy float32 and checks the error budget).  A ray the test
// rejects is lost at the crystal whatever its uniform; everything else -- including rays that miss the sphere, give
// a NaN here or have a deviate beyond the range of normal_approx -- is written to the id list and decided in FP64
// by k_trace, so results are identical with the phase switched off.
//
// SRC: 0 point source with a fixed axis; 1 box source with a fixed axis; 2 box source focused on a target; 3 plasma
// bundles (per-ray voxel origin, cone, line width and velocity from the bundle table).  Lines: constant or normal,
// with or without a Doppler shift.

enum { CULL_POINT = 0, CULL_BOX = 1, CULL_FOCUSED = 2, CULL_BUNDLES = 3 };

struct Cull32Par {
    float one_m_cos;       // 1 - cos(spread)                                     (not bundles)
    float basis[9];        // fixed axis: rows o_2, o_1, axis of the cone basis
    float Lb[3];           // C - source origin                                   (not bundles)
    float m[3];            // point source: basis . Lb, so that tca = l . m without forming the direction
    float mv[3];           // point source: basis . velocity / c
    float ll;              // point source: |Lb|^2
    float r2, inv_r, inv_r2;   // sphere
    float lam0, sig, inv_two_d;
    float t2, err;         // cull_t2; cull_err (+ the geometric margin for a point source, else added per ray from |C - O|^2)
    float R[9];            // source orientation rows: world offset = off . R
    float ext[3];          // box sizes (bundles: voxel size x 3)
    float Tb[3];           // target - source origin                              (focused)
    float xz[3];           // source xaxis + zaxis: o_1 = unit(axis x xz)
    float vel[3];          // velocity / c                                        (not bundles)
    float err_sig;         // 2e-3 |inv_two_d|: approximate-deviate term of err per unit sigma (bundles)
    int32_t moving;        // Doppler shift present
    int32_t normal_line;   // XRT_WAVE_NORMAL (else constant)
    // second stage (rays the first stage could not reject, re-packed): crystal bounds and the rocking-curve uniform
    int32_t stage2;        // enabled
    int32_t bounds_xy;     // the crystal checks |x| < hx, |y| < hy (and nothing else that the stage does not know)
    int32_t convex;        // sphere root: tca - thc instead of tca + thc
    int32_t gauss;         // Gaussian rocking curve: second pre-test level with the ray's uniform
    float Ob[3];           // source origin - crystal origin                      (not bundles)
    float ox[3], oy[3];    // crystal x and y axes
    float hx, hy;          // half sizes
    float lg_refl;         // log2(reflectivity)
    float two_sigma2;      // 2 sigma^2 of the rocking curve
    double Oc[3];          // crystal origin in FP64 (bundles)
    double C[3], T[3];     // sphere centre and target in FP64 (bundles: per-ray differences formed in FP64 once)
};

struct Cull32Out {
    uint32_t *ids;         // [n_regions][cap]
    uint32_t *counts;      // [n_regions]
    uint32_t n_regions, cap;
    unsigned int *next;        // region counter this launch claims from (zero on entry)
    unsigned int *next_reset;  // the counter of the launch after this one: zeroed here
};

#ifndef XRT_CULL_UNROLL
#define XRT_CULL_UNROLL 2
#endif
#ifndef XRT_CULL_BLOCKS
#define XRT_CULL_BLOCKS 6
#endif

// what the second stage needs of a ray
struct Cull32Full {
    float dx, dy, dz, tca, thc;      // direction, sphere chord
    float px, py, pz;                // ray origin - crystal origin
    float gap, c2, err;              // Bragg pre-test quantities
    float lx, ly, lz;                // sphere centre - ray origin
    float sB;                        // sin(theta_B) of the ray's (approximate) wavelength
    float cl[3];                     // local cone vector (point source: the direction is cl . basis, formed by whoever needs it)
    bool usable;
};
// what cull32_ray hands out besides its verdict: nothing; what the second stage needs (no direction for a point source,
// whose first stage never forms it); everything (k_mosaic32)
enum { CULL_OUT_NONE = 0, CULL_OUT_STAGE2 = 1, CULL_OUT_FULL = 2 };

// true = provably lost at the crystal
// first bundle b with bundle_end[b] > id (the search of source_local), bracketed by the hint table
__device__ __forceinline__ uint64_t cull32_bundle_of(const XrtSourceDesc &src, uint64_t id) {
    uint64_t blo = 0, bhi = src.n_bundles - 1;
    if (src.bundle_hint) {
        const uint64_t j = id >> src.bundle_hint_shift;
        blo = __ldg(src.bundle_hint + j);
        bhi = __ldg(src.bundle_hint + j + 1);
    }
    while (blo < bhi) {
        const uint64_t mid = (blo + bhi) >> 1;
        if (__ldg(src.bundle_end + mid) > id) bhi = mid; else blo = mid + 1;
    }
    return blo;
}

// The per-ray values of a plasma bundle in single precision (differences formed in FP64 once), as the first stage uses
// them: 4 x 16 bytes.  Ray ids are handed out bundle by bundle (inclusive prefix sum of the counts), so the 32 consecutive
// ids of a pass nearly always share the bundle of the pass before: a warp keeps the record of its current bundle in
// shared memory (k_cull32) and looks a bundle up only where a pass crosses a bundle boundary.
struct Cull32Bundle { float4 a, b, c, d; };     // {L, 1 - cos}, {T, sigma}, {v / c, margin}, {origin - crystal origin, -}

__device__ __forceinline__ Cull32Bundle cull32_bundle_record(const Cull32Par &K, const XrtBundle *bd) {
    const double ox = __ldg(&bd->origin[0]), oy = __ldg(&bd->origin[1]), oz = __ldg(&bd->origin[2]);
    Cull32Bundle r;
    const float sig = (float)__ldg(&bd->wave_sigma);
    r.a = make_float4((float)(K.C[0] - ox), (float)(K.C[1] - oy), (float)(K.C[2] - oz), (float)(1.0 - __ldg(&bd->cos_spread)));
    r.b = make_float4((float)(K.T[0] - ox), (float)(K.T[1] - oy), (float)(K.T[2] - oz), sig);
    r.c = make_float4((float)__ldg(&bd->velocity_c[0]), (float)__ldg(&bd->velocity_c[1]), (float)__ldg(&bd->velocity_c[2]),
                      fmaf(K.err_sig, fabsf(sig), K.err));
    r.d = make_float4((float)(ox - K.Oc[0]), (float)(oy - K.Oc[1]), (float)(oz - K.Oc[2]), 0.0f);
    return r;
}

// out of line: runs where a pass of k_cull32 crosses a bundle boundary, and must not cost the passes that do not any
// registers.  Returns the id at which the bundle of `id` ends.
static __device__ __noinline__ uint64_t cull32_bundle_miss(const Cull32Par &K, const XrtSourceDesc &src, uint64_t id, Cull32Bundle *out) {
    const uint64_t b = cull32_bundle_of(src, id);
    *out = cull32_bundle_record(K, src.bundles + b);
    return __ldg(src.bundle_end + b);
}

template <int SRC, int OUT = CULL_OUT_NONE>
__device__ __forceinline__ bool cull32_ray(const Cull32Par &K, const XrtSourceDesc &src, const PhiloxKeys &pk, uint32_t stream,
                                           uint32_t lo, uint32_t hi, Cull32Full *full = nullptr,
                                           const Cull32Bundle *cached = nullptr) {
    const uint4 r = philox4x32_10(make_uint4(lo, hi, SITE_CONE, stream), pk);
    float Px = K.Ob[0], Py = K.Ob[1], Pz = K.Ob[2];

    float one_m_cos = K.one_m_cos, sig = K.sig, err = K.err;
    float Lx = K.Lb[0], Ly = K.Lb[1], Lz = K.Lb[2];
    float Tx = K.Tb[0], Ty = K.Tb[1], Tz = K.Tb[2];
    float vx = K.vel[0], vy = K.vel[1], vz = K.vel[2];
    if constexpr (SRC == CULL_BUNDLES) {
        // bundle of this ray: the warp's cached record (every id of the pass lies in that bundle), else a lookup
        Cull32Bundle rec;
        if (cached) {
            rec = *cached;
        } else {
            const uint64_t id = ((uint64_t)hi << 32) | lo;
            rec = cull32_bundle_record(K, src.bundles + cull32_bundle_of(src, id));
        }
        Lx = rec.a.x; Ly = rec.a.y; Lz = rec.a.z; one_m_cos = rec.a.w;
        Tx = rec.b.x; Ty = rec.b.y; Tz = rec.b.z; sig = rec.b.w;
        vx = rec.c.x; vy = rec.c.y; vz = rec.c.z; err = rec.c.w;
        if constexpr (OUT != CULL_OUT_NONE) { Px = rec.d.x; Py = rec.d.y; Pz = rec.d.z; }
    }

    // ---- origin offset in world coordinates (the exact path: u01_42x3 of one block, off_k = ext_k (u_k - 1/2))
    if constexpr (SRC != CULL_POINT) {
        const uint4 ro = philox4x32_10(make_uint4(lo, hi, SITE_ORIGIN_XY, stream), pk);
        // top 23 bits of each uniform as a float in [1, 2): u - 1/2 = f - 3/2, no integer-to-float conversion
        const float o0 = K.ext[0] * (__uint_as_float(0x3f800000u | (ro.x >> 9)) - 1.5f);
        const float o1 = K.ext[1] * (__uint_as_float(0x3f800000u | ((ro.y << 10) >> 9) | (ro.z >> 31)) - 1.5f);
        const float o2 = K.ext[2] * (__uint_as_float(0x3f800000u | ((ro.z << 20) >> 9) | (ro.w >> 21)) - 1.5f);
        const float wx = o0 * K.R[0] + o1 * K.R[3] + o2 * K.R[6];
        const float wy = o0 * K.R[1] + o1 * K.R[4] + o2 * K.R[7];
        const float wz = o0 * K.R[2] + o1 * K.R[5] + o2 * K.R[8];
        Lx -= wx; Ly -= wy; Lz -= wz;
        Tx -= wx; Ty -= wy; Tz -= wz;
        if constexpr (OUT != CULL_OUT_NONE) { Px += wx; Py += wy; Pz += wz; }
    }

    // ---- local cone vector.  1 - a from the top 32 bits of the polar uniform (their complement): its RELATIVE
    // precision is what rho = sqrt(w (2 - w)) near the cone axis needs; the truncation error 2^-32 moves the direction
    // by 0.09 2^-32 / sqrt(1 - a) < 1e-7 unless 1 - a < 2^-24, and those rays (6e-8 of all) are left to FP64
    const uint32_t na = ~r.x;
    bool usable = na >= 256u;
    const float a1 = fmaf((float)na, 2.3283064365386963e-10f, 1.1641532182693481e-10f);
    const float w = one_m_cos * a1;                                                     // 1 - z
    const float z = 1.0f - w;
    const float rho = sqrt_approx(w * (2.0f - w));
    // azimuth 2 pi (b - 1/2) from the top 23 bits of b as a float in [1, 2): cos(2 pi b) = -cos(2 pi (b - 1/2))
    const float ang = 6.283185307179586f * (__uint_as_float(0x3f800000u | ((r.y & 0xfffu) << 11) | (r.z >> 21)) - 1.5f);
    const float lx = -rho * __cosf(ang), ly = -rho * __sinf(ang);

    // ---- direction and sphere chord
    float tca, ll, vd = 0.0f;
    float dx = 0.0f, dy = 0.0f, dz = 0.0f;
    if constexpr (SRC == CULL_POINT) {
        // fixed basis and fixed origin: L . D = l . (basis L), v . D = l . (basis v); |L|^2 is a constant
        tca = lx * K.m[0] + ly * K.m[1] + z * K.m[2];
        ll = K.ll;
        if constexpr (OUT == CULL_OUT_FULL) {
            dx = lx * K.basis[0] + ly * K.basis[3] + z * K.basis[6];
            dy = lx * K.basis[1] + ly * K.basis[4] + z * K.basis[7];
            dz = lx * K.basis[2] + ly * K.basis[5] + z * K.basis[8];
        }
    } else {
        if constexpr (SRC == CULL_BOX) {
            dx = lx * K.basis[0] + ly * K.basis[3] + z * K.basis[6];
            dy = lx * K.basis[1] + ly * K.basis[4] + z * K.basis[7];
            dz = lx * K.basis[2] + ly * K.basis[5] + z * K.basis[8];
        } else {
            // axis = unit(target - origin); o_1 = unit(axis x (xaxis + zaxis)); o_2 = axis x o_1 (unit up to rounding)
            const float it = rsqrt_approx(Tx * Tx + Ty * Ty + Tz * Tz);
            const float ax = Tx * it, ay = Ty * it, az = Tz * it;
            float px = ay * K.xz[2] - az * K.xz[1], py = az * K.xz[0] - ax * K.xz[2], pz = ax * K.xz[1] - ay * K.xz[0];
            const float ip = rsqrt_approx(px * px + py * py + pz * pz);
            px *= ip; py *= ip; pz *= ip;
            const float qx = ay * pz - az * py, qy = az * px - ax * pz, qz = ax * py - ay * px;
            dx = lx * qx + ly * px + z * ax;
            dy = lx * qy + ly * py + z * ay;
            dz = lx * qz + ly * pz + z * az;
        }
        tca = Lx * dx + Ly * dy + Lz * dz;
        ll = Lx * Lx + Ly * Ly + Lz * Lz;
        if (SRC == CULL_BUNDLES || K.moving) vd = vx * dx + vy * dy + vz * dz;
        // geometric margin 2e-5 max(1, |C - O|^2 / R^2); beyond 2 R from the centre of curvature the ray is left to FP64
        const float q = ll * K.inv_r2;
        err = fmaf(2e-5f, fmaxf(1.0f, q), err);
        usable &= q <= 4.0f;
    }
    const float d2 = fmaf(-tca, tca, ll);
    const float thc = sqrt_approx(K.r2 - d2);
    const float sI = thc * K.inv_r;

    // ---- sin(theta_B)
    float lam = K.lam0;
    if (K.normal_line) {
        bool in_range;
        lam = fmaf(normal_approx(r.w, in_range), sig, lam);
        usable &= in_range;
    }
    if constexpr (SRC == CULL_POINT) {
        // Doppler factor formed here, inside the one uniform branch that uses it (most scenes are at rest)
        if (K.moving) lam = fmaf(-lam, lx * K.mv[0] + ly * K.mv[1] + z * K.mv[2], lam);
    } else {
        if (SRC == CULL_BUNDLES || K.moving) lam = fmaf(-lam, vd, lam);
    }
    const float sB = lam * K.inv_two_d;

    const float gap = fabsf(sB - sI);
    const float diff = gap - err;
    const float c2 = fmaf(2.0f, gap, fmaf(-sI, sI, 1.0f));
    if constexpr (OUT != CULL_OUT_NONE) {
        full->cl[0] = lx; full->cl[1] = ly; full->cl[2] = z;
        full->dx = dx; full->dy = dy; full->dz = dz; full->tca = tca; full->thc = thc;
        full->px = Px; full->py = Py; full->pz = Pz;
        full->gap = gap; full->c2 = c2; full->err = err; full->usable = usable;
        full->lx = Lx; full->ly = Ly; full->lz = Lz; full->sB = sB;
    }
    return usable & (diff > 0.0f) & (diff * diff > K.t2 * c2);
}

// Second stage for one ray that survived the first, from the values the first stage left in the warp's queue (the ray
// is not generated again):
//   bounds   X = O + t D in the crystal's frame; a ray farther outside |x| < hx, |y| < hy than the rounding of the
//            chord arithmetic allows (4e-7 of the lengths that cancel in t = tca +- thc) is lost at the crystal;
//   level 2  with the ray's rocking-curve uniform u (words z, w of the SITE_WAVE block, the block the exact path reads):
//            reflected => dtheta^2 <= 2 sigma^2 ln(reflectivity / u); the first stage's lower bound on |dtheta| beyond
//            that means lost (bragg_cull_uniform in single precision, u truncated to 23 bits -- downwards, which only
//            widens the limit).
// Queue record (planes of 64 16-byte entries per warp): {offset | usable << 31, local cone vector}, {tca, thc, gap, c2}
// for the point source (its direction is cone vector . basis, origin and margin are constants); {offset | usable << 31,
// direction}, {tca, thc, gap, c2}, {origin - crystal origin, margin} for the box.
// Sources whose second stage is fed from the queue; for the others (per-ray focused basis, plasma bundles: more values,
// and few rays survive the first stage at all) the queue holds the offset alone and the second stage generates the ray
// again -- measured: hand-off +6.5 % for the point source, +2.5 % for the box, -2 % / -5 % for focused / bundles.
template <int SRC> __host__ __device__ constexpr bool cull_handoff() { return SRC == CULL_POINT || SRC == CULL_BOX; }
template <int SRC> __host__ __device__ constexpr int cull_planes() { return SRC == CULL_POINT ? 8 : (SRC == CULL_BOX ? 12 : 1); }
constexpr int kCullQ = 64;

template <int SRC>
__device__ __forceinline__ void cull32_push(uint32_t *q, int slot, uint32_t off, const Cull32Full &f) {
    const uint32_t word = off | (f.usable ? 0x80000000u : 0u);
    if constexpr (!cull_handoff<SRC>()) {
        q[slot] = word;
    } else {
        // 16-byte records in planes of kCullQ: two (point) or three (box) STS.128 per ray instead of 8 / 12 STS.32 --
        // consecutive slots are consecutive 16-byte words, so a quarter warp covers the 32 banks once
        float4 *q4 = (float4 *)q;
        if constexpr (SRC == CULL_POINT) {
            q4[slot] = make_float4(__uint_as_float(word), f.cl[0], f.cl[1], f.cl[2]);
        } else {
            q4[slot] = make_float4(__uint_as_float(word), f.dx, f.dy, f.dz);
            q4[2 * kCullQ + slot] = make_float4(f.px, f.py, f.pz, f.err);
        }
        q4[kCullQ + slot] = make_float4(f.tca, f.thc, f.gap, f.c2);
    }
}

// true = provably lost at the crystal
template <int SRC>
__device__ __forceinline__ bool cull32_stage2(const Cull32Par &K, const XrtSourceDesc &src, const PhiloxKeys &pk, uint32_t stream,
                                              uint32_t lo, uint32_t hi, const uint32_t *q, int slot, const float4 &head, bool usable) {
    const float4 *q4 = (const float4 *)q;
    float tca, thc, gap, c2, dx, dy, dz, px, py, pz, err;
    if constexpr (!cull_handoff<SRC>()) {
        Cull32Full f;
        if (cull32_ray<SRC, CULL_OUT_FULL>(K, src, pk, stream, lo, hi, &f)) return true;
        tca = f.tca; thc = f.thc; gap = f.gap; c2 = f.c2;
        dx = f.dx; dy = f.dy; dz = f.dz; px = f.px; py = f.py; pz = f.pz; err = f.err;
        usable = f.usable;
    } else {
        const float4 b = q4[kCullQ + slot];
        tca = b.x; thc = b.y; gap = b.z; c2 = b.w;
    }
    if constexpr (!cull_handoff<SRC>()) {
    } else if constexpr (SRC == CULL_POINT) {
        const float lx = head.y, ly = head.z, z = head.w;
        dx = lx * K.basis[0] + ly * K.basis[3] + z * K.basis[6];
        dy = lx * K.basis[1] + ly * K.basis[4] + z * K.basis[7];
        dz = lx * K.basis[2] + ly * K.basis[5] + z * K.basis[8];
        px = K.Ob[0]; py = K.Ob[1]; pz = K.Ob[2];
        err = K.err;
    } else {
        const float4 c = q4[2 * kCullQ + slot];
        dx = head.y; dy = head.z; dz = head.w;
        px = c.x; py = c.y; pz = c.z; err = c.w;
    }
    bool lost = false;
    if (K.bounds_xy) {
        const float t = K.convex ? tca - thc : tca + thc;
        const float X = fmaf(t, dx, px), Y = fmaf(t, dy, py), Z = fmaf(t, dz, pz);
        const float xl = X * K.ox[0] + Y * K.ox[1] + Z * K.ox[2];
        const float yl = X * K.oy[0] + Y * K.oy[1] + Z * K.oy[2];
        const float slack = fmaf(4e-7f, fabsf(tca) + fabsf(thc) + fabsf(px) + fabsf(py) + fabsf(pz), 1e-7f);
        lost = (fabsf(xl) > K.hx + slack) | (fabsf(yl) > K.hy + slack);       // NaN (sphere missed): not lost here
    }
    if (K.gauss && usable) {
        const uint4 b = philox4x32_10(make_uint4(lo, hi, SITE_WAVE, stream), pk);
        const float u = __uint_as_float(0x3f800000u | (b.z >> 9)) - 1.0f;
        const float lim = 0.6931471805599453f * (K.lg_refl - lg2_approx(u));
        const float bound = fmaf(fabsf(lim), 1e-3f, lim + 1e-3f) * K.two_sigma2;
        const float diff = gap - err;
        lost |= (diff > 0.0f) & (diff * diff > bound * c2) & (lim == lim);
    }
    return lost;
}

// groups of 32 consecutive ids per pass (independent chains): one for the bundle lookup, whose loads already overlap
template <int SRC> __host__ __device__ constexpr int cull_unroll() { return SRC == CULL_BUNDLES ? 1 : XRT_CULL_UNROLL; }

// stage 2 for `cnt` records popped from the warp's queue; survivors are appended to the region's list
template <int SRC, bool HIST>
__device__ __forceinline__ void cull32_drain(const Cull32Par &K, const XrtSourceDesc &src, const PhiloxKeys &pk, uint64_t stream_id,
                                             const XrtOutputs &out, unsigned lane, unsigned lt_mask, uint64_t id_first,
                                             uint32_t off_first, const uint32_t *q, int first, int cnt, uint32_t *dst, uint32_t &kept) {
    const bool active = (int)lane < cnt;
    const int slot = active ? first + (int)lane : first;
    float4 head = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
    uint32_t word;
    if constexpr (cull_handoff<SRC>()) {
        head = ((const float4 *)q)[slot];
        word = __float_as_uint(head.x);
    } else {
        word = q[slot];
    }
    const uint32_t off = word & 0x7fffffffu;
    const uint64_t id = id_first + off;
    const bool pass = active && !cull32_stage2<SRC>(K, src, pk, (uint32_t)stream_id, (uint32_t)id, (uint32_t)(id >> 32), q, slot, head, (word >> 31) != 0u);
    __syncwarp();
    if constexpr (HIST) {
        if (out.lost_count || out.lost_bits) {
            PhiloxDraws dr;
            dr.init(pk, stream_id, id, 0);
            emit_lost<true>(out, lane, lt_mask, dr, active && !pass, id);
        }
    }
    const unsigned m = __ballot_sync(kFull, pass);
    if (pass) dst[kept + __popc(m & lt_mask)] = off_first + off;
    kept += __popc(m);
}

// Lane 0 looks up the bundle of id_first + off and leaves its record in the warp's cache; returns the offset (from
// id_first, saturated) at which that bundle ends.
template <int SRC>
__device__ __forceinline__ uint32_t cull32_bundle_cache(const Cull32Par &K, const XrtSourceDesc &src, uint64_t id_first,
                                                        uint32_t off, unsigned lane, Cull32Bundle *bc) {
    uint32_t end = 0;
    if constexpr (SRC == CULL_BUNDLES) {
        __syncwarp();
        if (lane == 0) {
            const uint64_t e = cull32_bundle_miss(K, src, id_first + off, bc);
            // ids past the last bundle do not occur (ray_count is the total of the counts); e <= id can only mean that
            end = e > id_first + off ? (uint32_t)min(e - id_first, (uint64_t)0xffffffffu) : 0u;
        }
        end = __shfl_sync(kFull, end, 0);
        __syncwarp();                           // lane 0's stores before the other lanes' reads
    }
    return end;
}

// cull_unroll groups of 32 consecutive ids: first stage; what it cannot reject goes to the second stage through the
// warp's queue (K.stage2) or straight to the region's list
template <int SRC, bool HIST, bool CHECK>
__device__ __forceinline__ void cull32_pass(const Cull32Par &K, const XrtSourceDesc &src, const PhiloxKeys &pk, uint64_t stream_id,
                                            const XrtOutputs &out, unsigned lane, unsigned lt_mask, uint64_t id_first,
                                            uint32_t off_first, uint32_t g, uint32_t n_here, uint32_t *dst, uint32_t &kept,
                                            uint32_t *q, int &nq, Cull32Bundle *bc, uint32_t &bc_end) {
    constexpr int U = cull_unroll<SRC>();
#pragma unroll
    for (int j = 0; j < U; ++j) {
        const uint32_t off = g + 32u * j + lane;
        const bool valid = CHECK ? off < n_here : true;
        const uint64_t id = id_first + (valid ? off : 0u);      // lanes past the end re-test the region's first ray
        Cull32Full f;
        f.usable = true;
        bool pass;
        if constexpr (SRC == CULL_BUNDLES) {
            // offsets below bc_end lie in the warp's cached bundle (the region is walked upwards)
            // The warp's cached bundle covers the offsets below bc_end (the region is walked upwards).  A pass that crosses
            // a bundle boundary (3 in 1000 at 1e4 rays per bundle) looks every lane's bundle up, leaves the record in the
            // lane's own slot, and caches the bundle of the next pass's first id (if the region has one).
            const Cull32Bundle *rec = bc;
            const uint32_t next = g + 32u * j + 32u;
            if (next > bc_end) {
                Cull32Bundle *mine = bc + 1 + lane;
                cull32_bundle_miss(K, src, id, mine);
                rec = mine;
                bc_end = next < n_here ? cull32_bundle_cache<SRC>(K, src, id_first, next, lane, bc) : 0u;
            }
            pass = !cull32_ray<SRC, CULL_OUT_NONE>(K, src, pk, (uint32_t)stream_id, (uint32_t)id, (uint32_t)(id >> 32), &f, rec);
        } else {
            pass = !cull32_ray<SRC, cull_handoff<SRC>() ? CULL_OUT_STAGE2 : CULL_OUT_NONE>(K, src, pk, (uint32_t)stream_id, (uint32_t)id, (uint32_t)(id >> 32), &f);
        }
        if constexpr (CHECK) pass = pass && valid;
        if constexpr (HIST) {
            if (out.lost_count || out.lost_bits) {
                PhiloxDraws dr;
                dr.init(pk, stream_id, id_first + off, 0);
                emit_lost<true>(out, lane, lt_mask, dr, valid && !pass, id_first + off);
            }
        }
        const unsigned m = __ballot_sync(kFull, pass);
        if (K.stage2) {
            if (pass) cull32_push<SRC>(q, nq + __popc(m & lt_mask), off, f);
            nq += __popc(m);
            __syncwarp();
            if (nq >= 32) {         // at most 63 queued: one full pop keeps the queue below 32 + the next push
                nq -= 32;
                cull32_drain<SRC, HIST>(K, src, pk, stream_id, out, lane, lt_mask, id_first, off_first, q, nq, 32, dst, kept);
            }
        } else {
            if (pass) dst[kept + __popc(m & lt_mask)] = off_first + off;
            kept += __popc(m);
        }
    }
}

// resident blocks per SM: the lost-sample emission of history-on launches needs more registers than the rest (4 x 64);
// the plasma source with the warp's bundle record cached runs without spills at 80 registers, and 3 resident blocks of
// that beat 4 blocks with 160 B of spills (9.3e10 against 8.8e10 rays/s on config 5; 8.4e10 before the cache)
#ifndef XRT_CULL_BLOCKS_HIST
#define XRT_CULL_BLOCKS_HIST 4
#endif
#ifndef XRT_CULL_BLOCKS_BUNDLES
#define XRT_CULL_BLOCKS_BUNDLES 3
#endif
template <int SRC, bool HIST>
__global__ void __launch_bounds__(kBlock, (SRC == CULL_BUNDLES ? XRT_CULL_BLOCKS_BUNDLES : (HIST ? XRT_CULL_BLOCKS_HIST : XRT_CULL_BLOCKS)))
k_cull32(const __grid_constant__ Cull32Par K, const __grid_constant__ XrtSourceDesc src, const __grid_constant__ PhiloxKeys pk,
         const uint64_t stream_id, const uint64_t ray_begin, const uint64_t ray_count, const Cull32Out lst,
         const XrtOutputs out) {
    const unsigned lane = threadIdx.x & 31u;
    const unsigned lt_mask = (1u << lane) - 1u;
    const uint32_t n_warps = gridDim.x * (kBlock / 32);
    const uint32_t warp_global = blockIdx.x * (kBlock / 32) + (threadIdx.x >> 5);
    constexpr uint32_t kPass = 32u * cull_unroll<SRC>();
    __shared__ __align__(16) uint32_t s_q2[kBlock / 32][cull_planes<SRC>() * kCullQ];     // per warp: records waiting for the second stage
    uint32_t *q = s_q2[threadIdx.x >> 5];
    int nq = 0;
    // per warp: record of its current bundle + one slot per lane for the passes that cross a bundle boundary
    __shared__ Cull32Bundle s_bc[SRC == CULL_BUNDLES ? (kBlock / 32) * 33 : 1];
    Cull32Bundle *bc = s_bc + (SRC == CULL_BUNDLES ? (threadIdx.x >> 5) * 33 : 0);
    uint32_t bc_end = 0;
    // Regions are claimed from a global counter: with a static share per warp the scheduler's oldest-first policy lets
    // the old warps of an SM finish early and the SM runs its last third at half occupancy (ncu: 49 % achieved of 75 %).
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        *lst.next_reset = 0u;
        // rays out of the source: every region of the launch is claimed by exactly one warp
        if (out.counts && ray_count) atomicAdd((unsigned long long *)out.counts, (unsigned long long)ray_count);
    }
    (void)n_warps; (void)warp_global;
    for (;;) {
        uint32_t reg = 0;
        if (lane == 0) reg = atomicAdd(lst.next, 1u);
        reg = __shfl_sync(kFull, reg, 0);
        if (reg >= lst.n_regions) break;
        const uint64_t first = (uint64_t)reg * lst.cap;
        const uint64_t left = ray_count - first;
        const uint32_t n_here = left < (uint64_t)lst.cap ? (uint32_t)left : lst.cap;
        uint32_t *dst = lst.ids + first;
        uint32_t kept = 0;
        uint32_t g = 0;
        bc_end = cull32_bundle_cache<SRC>(K, src, ray_begin + first, 0u, lane, bc);
        for (; g + kPass <= n_here; g += kPass)
            cull32_pass<SRC, HIST, false>(K, src, pk, stream_id, out, lane, lt_mask, ray_begin + first, (uint32_t)first, g, n_here, dst, kept, q, nq, bc, bc_end);
        if (g < n_here)
            cull32_pass<SRC, HIST, true>(K, src, pk, stream_id, out, lane, lt_mask, ray_begin + first, (uint32_t)first, g, n_here, dst, kept, q, nq, bc, bc_end);
        if (nq > 0) {               // the queue holds offsets of this region only: drain it before the next one
            cull32_drain<SRC, HIST>(K, src, pk, stream_id, out, lane, lt_mask, ray_begin + first, (uint32_t)first, q, 0, nq, dst, kept);
            nq = 0;
        }
        if (lane == 0) lst.counts[reg] = kept;
    }
}

}  // namespace xrt
#include "xrt_meshsort.cuh"
namespace xrt {

// ---------------------------------------------------------------------------
// FP32 broad phase of a spherical MOSAIC crystal as first optic (k_mosaic32)
//
// A HOPG-like crystal reflects a ray at the first of its mosaic_depth crystallite layers that satisfies Bragg's law
// with its own random normal; ~97 % of the (ray, layer) pairs fail by many rocking-curve widths.  Stage S of k_trace
// rejects them with an FP32 pre-test per layer (mosaic_pretest), but it does so inside the FP64 kernel: 128
// registers, 16 warps per SM, 140 kB of code.  This kernel runs the same scan for every ray of the launch in a small
// single-precision kernel of its own, with the ray generated in FP32 from its Philox blocks (cull32_ray):
//
//   feed   32 consecutive ids: direction, sphere chord, intersection point, crystal bounds (a ray farther outside
//          than the rounding allows is lost at the crystal), the frame (n, r_0, r_1) of mosaic_normal at the point and
//          the three dot products of the pre-test -> per-warp queue (offset, dr0, dr1, dn, sB, err)
//   scan   every lane owns one queued ray and tests two layers per iteration; a lane whose ray is finished (first
//          layer that cannot be rejected found, or all layers rejected = lost) takes the next ray from the queue, so
//          the scan runs with nearly all lanes busy
//
// A ray with a candidate layer leaves as (id offset << 5 | layer) in the region's list; k_trace starts its exact FP64
// scan at that layer.  Rays the single-precision geometry cannot judge (sphere missed or nearly so, deviate outside
// the range of normal_approx, NaN) leave with layer 0.  Every FP32 quantity is within ~1e-6 of its FP64 value; the
// margin err carries 2e-5 (x |C - O|^2 / R^2) + the approximate-deviate term + stage S's own 2e-6, so a layer rejected
// here is rejected by the exact test as well and results are identical with the phase off (tests/test_gpu_scale.py).

struct Mosaic32Par {
    float sin_sigma, err, t2, two_sigma2, lg_refl;
    int32_t depth, gauss;
};

#ifndef XRT_MOSAIC32_BLOCKS
#define XRT_MOSAIC32_BLOCKS 4
#endif
constexpr int kMosaicTagBits = 5;          // layer index in a list entry: mosaic_depth <= 32

template <int SRC, bool HIST>
__global__ void __launch_bounds__(kBlock, XRT_MOSAIC32_BLOCKS)
k_mosaic32(const __grid_constant__ Cull32Par K, const __grid_constant__ Mosaic32Par M, const __grid_constant__ XrtSourceDesc src,
           const __grid_constant__ PhiloxKeys pk, const uint64_t stream_id, const uint64_t ray_begin, const uint64_t ray_count,
           const Cull32Out lst, const XrtOutputs out) {
    const unsigned lane = threadIdx.x & 31u;
    const unsigned lt_mask = (1u << lane) - 1u;
    constexpr int kQ = 64;
    __shared__ uint32_t s_off[kBlock / 32][kQ];
    __shared__ float s_par[kBlock / 32][5][kQ];
    uint32_t *q_off = s_off[threadIdx.x >> 5];
    float(*q_par)[kQ] = s_par[threadIdx.x >> 5];
    MosaicPre MP;
    MP.s32 = M.sin_sigma;
    MP.t2 = M.t2;
    MP.two_sigma2 = M.two_sigma2;
    MP.lg_refl = M.lg_refl;
    MP.gauss = M.gauss != 0;
    const int depth = M.depth;
    const uint32_t stream = (uint32_t)stream_id;
    unsigned long long n_src = 0;
    if (blockIdx.x == 0 && threadIdx.x == 0) *lst.next_reset = 0u;
    for (;;) {
        uint32_t reg = 0;
        if (lane == 0) reg = atomicAdd(lst.next, 1u);
        reg = __shfl_sync(kFull, reg, 0);
        if (reg >= lst.n_regions) break;
        const uint64_t first = (uint64_t)reg * lst.cap;
        const uint64_t left = ray_count - first;
        const uint32_t n_here = left < (uint64_t)lst.cap ? (uint32_t)left : lst.cap;
        const uint64_t id_first = ray_begin + first;
        const uint32_t off_first = (uint32_t)first;
        uint32_t *dst = lst.ids + first;
        uint32_t kept = 0;
        n_src += n_here;
        int nq = 0;
        uint32_t g = 0;
        // the ray this lane is scanning
        bool busy = false;
        uint32_t off = 0;
        int layer = 0;
        float dr0 = 0.0f, dr1 = 0.0f, dn = 0.0f, sB = 0.0f, err = 0.0f;
        for (;;) {
            if (g < n_here && nq < 32) {
                // ---- feed: 32 consecutive ids
                const uint32_t o1 = g + lane;
                const bool valid = o1 < n_here;
                const uint64_t id = id_first + (valid ? o1 : 0u);
                Cull32Full f;
                cull32_ray<SRC, CULL_OUT_FULL>(K, src, pk, stream, (uint32_t)id, (uint32_t)(id >> 32), &f);
                const float t = K.convex ? f.tca - f.thc : f.tca + f.thc;
                // frame of mosaic_normal at the intersection point: n = (C - X) / R with C - X = L - t D
                const float nx = (f.lx - t * f.dx) * K.inv_r, ny = (f.ly - t * f.dy) * K.inv_r, nz = (f.lz - t * f.dz) * K.inv_r;
                float ax = ny, ay = nz - nx, az = -ny;                                          // r_0 ~ n x x + n x z
                const float a2 = ax * ax + ay * ay + az * az;
                // t is NaN when the chord is (sphere missed in FP32); a nearly degenerate frame (n along x + z) would
                // amplify the rounding of n
                const bool judged = f.usable && (t == t) && a2 > 0.01f;
                bool outside = false;
                if (K.bounds_xy) {
                    const float X = fmaf(t, f.dx, f.px), Y = fmaf(t, f.dy, f.py), Z = fmaf(t, f.dz, f.pz);
                    const float xl = X * K.ox[0] + Y * K.ox[1] + Z * K.ox[2];
                    const float yl = X * K.oy[0] + Y * K.oy[1] + Z * K.oy[2];
                    const float slack = fmaf(4e-7f, fabsf(f.tca) + fabsf(f.thc) + fabsf(f.px) + fabsf(f.py) + fabsf(f.pz), 1e-7f);
                    outside = (fabsf(xl) > K.hx + slack) | (fabsf(yl) > K.hy + slack);
                }
                const bool lost = valid && judged && outside;
                const bool scan = valid && judged && !outside;
                const bool pass0 = valid && !judged;             // left to k_trace, from layer 0
                if constexpr (HIST) {
                    if (out.lost_count || out.lost_bits) {
                        PhiloxDraws dr;
                        dr.init(pk, stream_id, id, 0);
                        emit_lost<true>(out, lane, lt_mask, dr, lost, id);
                    }
                }
                const unsigned mp = __ballot_sync(kFull, pass0);
                if (pass0) dst[kept + __popc(mp & lt_mask)] = (off_first + o1) << kMosaicTagBits;
                kept += __popc(mp);
                const unsigned ms = __ballot_sync(kFull, scan);
                if (scan) {
                    const float ia = rsqrt_approx(a2);
                    ax *= ia; ay *= ia; az *= ia;
                    float bx = ny * az - nz * ay, by = nz * ax - nx * az, bz = nx * ay - ny * ax;   // r_1 ~ n x r_0
                    const float ib = rsqrt_approx(bx * bx + by * by + bz * bz);
                    const int slot = nq + __popc(ms & lt_mask);
                    q_off[slot] = o1;
                    q_par[0][slot] = f.dx * ax + f.dy * ay + f.dz * az;
                    q_par[1][slot] = (f.dx * bx + f.dy * by + f.dz * bz) * ib;
                    q_par[2][slot] = f.dx * nx + f.dy * ny + f.dz * nz;
                    q_par[3][slot] = f.sB;
                    q_par[4][slot] = f.err + M.err;
                }
                nq += __popc(ms);
                g += 32u;
                __syncwarp();
            }
            // ---- refill: idle lanes take the last entries of the queue
            {
                const unsigned mi = __ballot_sync(kFull, !busy);
                const int r = __popc(mi & lt_mask);
                if (!busy && r < nq) {
                    const int slot = nq - 1 - r;
                    off = q_off[slot];
                    dr0 = q_par[0][slot]; dr1 = q_par[1][slot]; dn = q_par[2][slot]; sB = q_par[3][slot]; err = q_par[4][slot];
                    layer = 0;
                    busy = true;
                }
                const int taken = __popc(mi);
                nq -= taken < nq ? taken : nq;
                __syncwarp();
            }
            if (!__ballot_sync(kFull, busy)) {
                if (g >= n_here) break;          // region done: nothing queued, nothing being scanned
                continue;
            }
            // ---- scan: two layers (independent chains; the second is wasted only when the first is a candidate)
            bool found = false, gone = false;
            if (busy) {
                const uint64_t id = id_first + off;
                const uint32_t lo = (uint32_t)id, hi = (uint32_t)(id >> 32);
                const uint4 b0 = philox4x32_10(make_uint4(lo, hi, site_optic(0, layer, 1), stream), pk);
                const uint4 b1 = philox4x32_10(make_uint4(lo, hi, site_optic(0, layer + 1, 1), stream), pk);
                const bool rej0 = mosaic_pretest(b0, MP, dr0, dr1, dn, sB, err);
                const bool rej1 = mosaic_pretest(b1, MP, dr0, dr1, dn, sB, err);
                if (!rej0) found = true;
                else if (layer + 1 >= depth) gone = true;
                else if (!rej1) { layer += 1; found = true; }
                else { layer += 2; gone = layer >= depth; }
            }
            if constexpr (HIST) {
                if (out.lost_count || out.lost_bits) {
                    PhiloxDraws dr;
                    dr.init(pk, stream_id, id_first + off, 0);
                    emit_lost<true>(out, lane, lt_mask, dr, gone, id_first + off);
                }
            }
            const unsigned mf = __ballot_sync(kFull, found);
            if (found) dst[kept + __popc(mf & lt_mask)] = ((off_first + off) << kMosaicTagBits) | (uint32_t)layer;
            kept += __popc(mf);
            if (found || gone) busy = false;
        }
        if (lane == 0) lst.counts[reg] = kept;
    }
    __shared__ unsigned long long s_src;
    if (threadIdx.x == 0) s_src = 0ull;
    __syncthreads();
    if (lane == 0 && n_src) atomicAdd(&s_src, n_src);
    __syncthreads();
    if (threadIdx.x == 0 && s_src && out.counts) atomicAdd((unsigned long long *)out.counts, s_src);
}

// ---------------------------------------------------------------------------
// recording kernel: history of every element, optional counters / images

// streaming stores (evict-first): history planes are written once and read back by the host
__device__ __forceinline__ void store_history(const XrtHistory &h, int elem, uint64_t slot, const Ray &r) {
    if (h.rays) {
        const uint64_t c = h.capacity;
        double *e = h.rays + ((uint64_t)elem * 7) * c;
        if (h.layout == XRT_HIST_ROWS) {
            // the reference's (n, 3) row arrays: a warp writes 768 contiguous bytes per array, three 8-byte stores per
            // thread that the L2 merges into full lines; the host then needs no transposition at all
            double *po = e + 3 * slot, *pd = e + 3 * c + 3 * slot;
            __stcs(po, r.o.x); __stcs(po + 1, r.o.y); __stcs(po + 2, r.o.z);
            __stcs(pd, r.d.x); __stcs(pd + 1, r.d.y); __stcs(pd + 2, r.d.z);
            __stcs(e + 6 * c + slot, r.w);
        } else {
            double *p = e + slot;
            __stcs(p, r.o.x); __stcs(p + c, r.o.y); __stcs(p + 2 * c, r.o.z);
            __stcs(p + 3 * c, r.d.x); __stcs(p + 4 * c, r.d.y); __stcs(p + 5 * c, r.d.z);
            __stcs(p + 6 * c, r.w);
        }
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
    // the same azimuth routine as the fused kernel (sincos_2pi_tab), so that a replayed ray is bit for bit the ray
    // the fused kernel classified
    __shared__ double s_sincos[MODE == REC_PHILOX ? 2 * kSincosTable : 2];
    if constexpr (MODE == REC_PHILOX) {
        for (int i = threadIdx.x; i < kSincosTable; i += kBlock) {
            double sn, cs;
            sincos_2pi((double)i / (double)kSincosTable, sn, cs);
            s_sincos[2 * i] = cs;
            s_sincos[2 * i + 1] = sn;
        }
        __syncthreads();
    }
    const int nopt = sc.n_optics;
    const uint64_t stride = (uint64_t)gridDim.x * kBlock;
    for (uint64_t i = (uint64_t)blockIdx.x * kBlock + threadIdx.x; i < n; i += stride) {
        Ray r;
        PhiloxDraws pdr;
        InjectedDraws idr;
        if constexpr (MODE == REC_PHILOX) {
            const uint64_t id = ids ? ids[i] : ray_begin + i;
            pdr.init(pk, stream_id, id, split);
            generate_ray<FT, PhiloxDraws, KN, true>(sc.source, pdr, id, r, s_sincos);
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
    __shared__ double s_sincos[MODE == REC_PHILOX ? 2 * kSincosTable : 2];
    if constexpr (MODE == REC_PHILOX) {
        for (int i = threadIdx.x; i < kSincosTable; i += kBlock) {
            double sn, cs;
            sincos_2pi((double)i / (double)kSincosTable, sn, cs);
            s_sincos[2 * i] = cs;
            s_sincos[2 * i + 1] = sn;
        }
        __syncthreads();
    }
    const uint64_t stride = (uint64_t)gridDim.x * kBlock;
    for (uint64_t i = (uint64_t)blockIdx.x * kBlock + threadIdx.x; i < n; i += stride) {
        Ray r;
        if constexpr (MODE == REC_PHILOX) {
            PhiloxDraws dr;
            dr.init(pk, stream_id, ray_begin + i, -1);
            generate_ray<FT_FULL, PhiloxDraws, 0, true>(sc.source, dr, ray_begin + i, r, s_sincos);
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

}  // namespace xrt
