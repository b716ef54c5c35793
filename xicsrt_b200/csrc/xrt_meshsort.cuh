// xrt_meshsort.cuh -- sorted path for a refining mesh as first optic (included by xrt_kernels.cuh).
//
// The refinement of a coarse-mesh hit (nearest fine vertex -> the faces around it -> Clough-Tocher interpolation)
// is a chain of dependent table reads, ~1 kB per ray out of several MB of tables.  With the rays of a warp spread
// over the whole crystal every one of those reads goes to L2 and the fused kernel waits on the scoreboard 55 % of
// the time (ncu, profiles/r02_c4_*).  The sorted path orders the rays by WHERE they hit before refining them:
//
//   k_mesh_coarse   every ray of the launch: source geometry + step 1 of the intersection (coarse mesh).  A hit is
//                   written as a packed entry (id offset << 5 | coarse face) with the spatial bin of its hit point
//                   (tiles of the vertex grid); per-bin histogram.  Light kernel, no table reads.
//   k_mesh_scan     exclusive scan of the histogram (one block).
//   k_mesh_scatter  counting-sort scatter of the entries by bin (shared-memory ranks, one global atomic per block,
//                   region and bin).
//   k_mesh_refine   32 consecutive sorted entries per warp pass: rebuilds the ray from its id (Philox is counter
//                   based) and the coarse hit point from the recorded face, finishes the intersection (nearest
//                   vertex, candidate faces, interpolation), interaction, remaining optics, images, found / lost
//                   lists.  No shared-memory queues (nearly every coarse hit is a crystal hit, so there is nothing to
//                   re-pack), which leaves the SM's 256 kB to L1: the warps resident at any time work on one or two
//                   bins, whose tables (~25 kB each) stay in L1.
//
// Only the order in which rays are processed changes: counters, images and found / lost sets are the sums and sets
// of the unsorted path (tests/test_gpu_scale.py compares them with XRT_NO_MESH_SORT=1).
#pragma once

namespace xrt {

constexpr int kMeshMaxBins = 8192;
constexpr int kMeshFaceBits = 5;            // packed entry: coarse face in the low bits (<= 32 coarse faces)
constexpr uint64_t kMeshMaxLaunch = 1ull << (32 - kMeshFaceBits);

// Direction grid of a point source with a fixed cone axis: which coarse faces a ray can hit is a function of its
// direction alone.  In the gnomonic coordinates of the cone frame, (p, q) = (D.ex, D.ey) / D.ez, the three edge
// functions of a face (mesh_all_faces_point) are linear, so a cell of a uniform (p, q) grid is tested against a face
// by interval arithmetic on its four linear forms; mask[cell] holds the faces that cannot be excluded (k_mesh_dirgrid,
// built once per scene).  k_mesh_coarse then runs the pre-selection on the 1-6 faces of the ray's cell instead of all
// of them; the faces that survive are decided by the reference's arithmetic as before.
constexpr int kDirGrid = 64;
struct MeshDirGrid {
    const uint32_t *mask;      // [kDirGrid][kDirGrid], nullptr = no grid for this scene
    double ex[3], ey[3], ez[3];   // cone frame (rows o_2, o_1, axis of generate_geometry) in the optic's tracing frame
    double half, inv_h;        // p, q in [-half, half]; cells per unit
};

__device__ __forceinline__ int mesh_faces_point_masked(const double *__restrict__ naq, const double *__restrict__ geom,
                                                       unsigned todo, V3 o, V3 d, V3 &X) {
    const double eps = 1e-15, tol = 1e-9;
    unsigned cand = 0u;
    while (todo) {
        const int j = __ffs(todo) - 1;
        todo &= todo - 1u;
        const double *c = naq + kPointRec * j;
        const double a = d.x * c[0] + d.y * c[1] + d.z * c[2];
        double ua = d.x * c[3] + d.y * c[4] + d.z * c[5];
        double va = d.x * c[6] + d.y * c[7] + d.z * c[8];
        const double aa = fabs(a);
        if (a < 0.0) { ua = -ua; va = -va; }
        const double slack = fma(tol, aa, c[9]);
        const double lo = -slack, hi = aa + slack;
        const bool out_side = (ua < lo) | (ua > hi) | (va < lo) | (ua + va > hi);
        if (!out_side || aa < 4.0 * eps) cand |= 1u << j;
    }
    int hit = -1;
    while (cand) {
        const int j = __ffs(cand) - 1;
        cand &= cand - 1u;
        V3 P;
        if (mesh_test_face_mt(geom + 9 * j, o, d, P)) { hit = j; X = P; }
    }
    return hit;
}

struct MeshSortOut {
    uint32_t *entries;         // [n_regions][cap] unsorted packed entries
    uint16_t *bins;            // [n_regions][cap] bin of each entry
    uint32_t *counts;          // [n_regions]
    uint32_t n_regions, cap;
    unsigned int *next;        // region counter this launch claims from (zero on entry)
    unsigned int *next_reset;  // the counter of the launch after this one: zeroed here
    unsigned int *hist;        // [n_bins] entries per bin (zero on entry; k_mesh_scan zeroes it again)
    int32_t n_bins, tile, tiles_x, sub;
    MeshDirGrid dg;
};

template <uint32_t FT, bool HIST>
__global__ void __launch_bounds__(kBlock, 3)
k_mesh_coarse(const __grid_constant__ XrtSceneDesc sc, const __grid_constant__ PhiloxKeys pk, const uint64_t stream_id,
              const uint64_t ray_begin, const uint64_t ray_count, const MeshSortOut lst, const XrtOutputs out) {
    extern __shared__ double s_dyn[];           // step-1 face operands (or point constants), then the block's histogram
    __shared__ double s_sincos[2 * kSincosTable];
    for (int i = threadIdx.x; i < kSincosTable; i += kBlock) {
        double sn, cs;
        sincos_2pi((double)i / (double)kSincosTable, sn, cs);
        s_sincos[2 * i] = cs;
        s_sincos[2 * i + 1] = sn;
    }
    const XrtOpticDesc &ops = sc.optics[0];
    const XrtSourceDesc &src = sc.source;
    const double *geom;
    const int nf = mesh_stage1_faces(ops, geom);
    // the host enables this path for nf <= 32 (5-bit face tag), so the operands always fit the staging area
    const bool mesh_point = src.kind != XRT_SRC_BUNDLES && src.extent[0] == 0.0 && src.extent[1] == 0.0 && src.extent[2] == 0.0 &&
                            src.spatial == XRT_SPATIAL_UNIFORM && kPointRec * nf <= 9 * kStageFaces;
    if (mesh_point) {
        V3 o = v3(src.origin);
        if (optic_is_local<FT>(ops)) o = to_local(ops.orient, o - v3(ops.origin));
        for (int i = threadIdx.x; i < nf; i += kBlock) mesh_point_constants(geom + 9 * i, o, s_dyn + kPointRec * i);
    } else {
        for (int i = threadIdx.x; i < 9 * nf; i += kBlock) s_dyn[i] = __ldg(geom + i);
    }
    unsigned int *s_hist = (unsigned int *)(s_dyn + 9 * kStageFaces);
    for (int i = threadIdx.x; i < lst.n_bins; i += kBlock) s_hist[i] = 0u;
    uint32_t *s_grid = (uint32_t *)(s_hist + kMeshMaxBins);
    const bool use_grid = mesh_point && lst.dg.mask != nullptr;
    if (use_grid)
        for (int i = threadIdx.x; i < kDirGrid * kDirGrid; i += kBlock) s_grid[i] = __ldg(lst.dg.mask + i);
    const uint32_t all_faces = nf >= 32 ? 0xffffffffu : ((1u << nf) - 1u);
    __syncthreads();

    const unsigned lane = threadIdx.x & 31u;
    const unsigned lt_mask = (1u << lane) - 1u;
    const XrtMesh &mesh = *ops.mesh;
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
        uint32_t kept = 0;
        for (uint32_t g = 0; g < n_here; g += 32u) {
            const uint32_t off = g + lane;
            const bool valid = off < n_here;
            const uint64_t id = ray_begin + first + (valid ? off : 0u);
            PhiloxDraws dr;
            dr.init(pk, stream_id, id, 0);
            Ray r;
            r.alive = false;
            r.w = 0.0;
            if (valid) {
                SrcLocal L;
                source_local<FT, 0>(src, id, L);
                generate_geometry<FT, PhiloxDraws, 0, true>(src, L, dr, r, s_sincos);
            }
            n_src += __popc(__ballot_sync(kFull, r.alive));
            int face = -1;
            V3 Xc = nan3();
            if (r.alive) {
                V3 o = r.o, d = r.d;
                if (optic_is_local<FT>(ops)) {
                    o = to_local(ops.orient, o - v3(ops.origin));
                    d = to_local(ops.orient, d);
                }
                if (use_grid) {
                    const MeshDirGrid &G = lst.dg;
                    const double dz = d.x * G.ez[0] + d.y * G.ez[1] + d.z * G.ez[2];
                    const double iz = 1.0 / dz;
                    const double p = (d.x * G.ex[0] + d.y * G.ex[1] + d.z * G.ex[2]) * iz;
                    const double q = (d.x * G.ey[0] + d.y * G.ey[1] + d.z * G.ey[2]) * iz;
                    uint32_t mask = all_faces;
                    // a ray outside the grid by more than the cells' own margin (never, for an isotropic cone) or
                    // with a direction behind the cone plane takes every face
                    if (dz > 0.0 && fabs(p) <= G.half * 1.0000001 && fabs(q) <= G.half * 1.0000001) {
                        const int ix = min(max((int)floor((p + G.half) * G.inv_h), 0), kDirGrid - 1);
                        const int iy = min(max((int)floor((q + G.half) * G.inv_h), 0), kDirGrid - 1);
                        mask = s_grid[iy * kDirGrid + ix];
                    }
                    face = mesh_faces_point_masked(s_dyn, geom, mask, o, d, Xc);
                } else {
                    face = mesh_coarse_face(ops, o, d, Xc, s_dyn, mesh_point);
                }
            }
            const bool hit = face >= 0;
            emit_lost<HIST>(out, lane, lt_mask, dr, valid && !hit, id);
            const unsigned m = __ballot_sync(kFull, hit);
            if (hit) {
                const int bin = mesh_bin_of(mesh, Xc, lst.sub, lst.tile, lst.tiles_x);
                const uint64_t slot = first + kept + __popc(m & lt_mask);
                lst.entries[slot] = ((uint32_t)(first + off) << kMeshFaceBits) | (uint32_t)face;
                lst.bins[slot] = (uint16_t)bin;
                atomicAdd(&s_hist[bin], 1u);
            }
            kept += __popc(m);
        }
        if (lane == 0) lst.counts[reg] = kept;
    }
    __shared__ unsigned long long s_src;
    if (threadIdx.x == 0) s_src = 0ull;
    __syncthreads();
    if (lane == 0 && n_src) atomicAdd(&s_src, n_src);
    __syncthreads();
    if (threadIdx.x == 0 && s_src && out.counts) atomicAdd((unsigned long long *)out.counts, s_src);
    for (int i = threadIdx.x; i < lst.n_bins; i += kBlock) {
        const unsigned int v = s_hist[i];
        if (v) atomicAdd(lst.hist + i, v);
    }
}

#ifdef XRT_MESHSORT_HOST_KERNELS      // defined by xrt.cu alone: the sort and grid kernels are not templates
// one thread per cell of the direction grid; o = the point source in the optic's tracing frame
__global__ void __launch_bounds__(kBlock) k_mesh_dirgrid(const double *__restrict__ geom, const int nf, const V3 o,
                                                         const MeshDirGrid G, uint32_t *__restrict__ mask) {
    const int cell = blockIdx.x * kBlock + threadIdx.x;
    if (cell >= kDirGrid * kDirGrid) return;
    const int ix = cell % kDirGrid, iy = cell / kDirGrid;
    const double h = 1.0 / G.inv_h;
    const double pc = -G.half + (ix + 0.5) * h, qc = -G.half + (iy + 0.5) * h;
    const double hh = 0.5 * h * 1.001;                      // half cell, 0.1 % margin on every side
    const V3 ex = v3(G.ex), ey = v3(G.ey), ez = v3(G.ez);
    // |D| of the un-normalised direction (p, q, 1) is at most sqrt(1 + 2 half^2): absolute slacks scale with it
    const double scale = sqrt(1.0 + 2.0 * G.half * G.half);
    uint32_t m = 0u;
    for (int j = 0; j < nf; ++j) {
        double c[kPointRec];
        mesh_point_constants(geom + 9 * j, o, c);
        double lo[4], hi[4];      // a, ua, va, w = a - ua - va over the cell
        double wx = 0.0, wy = 0.0, wz = 0.0;
        // the per-ray pre-selection allows tol |a| (tol = 1e-9) on every edge: |a| <= abound over the grid
        const V3 vn = v3(c[0], c[1], c[2]);
        const double abound = (fabs(dot(vn, ex)) + fabs(dot(vn, ey))) * G.half + fabs(dot(vn, ez));
        for (int f = 0; f < 3; ++f) {
            const V3 v = v3(c[3 * f], c[3 * f + 1], c[3 * f + 2]);
            const double fx = dot(v, ex), fy = dot(v, ey), fz = dot(v, ez);
            const double sg = f == 0 ? 1.0 : -1.0;
            wx += sg * fx; wy += sg * fy; wz += sg * fz;
            const double mid = pc * fx + qc * fy + fz;
            const double rad = (fabs(fx) + fabs(fy)) * hh + 1e-6 * (fabs(fx) * G.half + fabs(fy) * G.half + fabs(fz)) +
                               2.0 * c[9] * scale + 4e-9 * abound;
            lo[f] = mid - rad; hi[f] = mid + rad;
        }
        {
            const double mid = pc * wx + qc * wy + wz;
            const double rad = (fabs(wx) + fabs(wy)) * hh + 1e-6 * (fabs(wx) * G.half + fabs(wy) * G.half + fabs(wz)) +
                               6.0 * c[9] * scale + 1.2e-8 * abound;
            lo[3] = mid - rad; hi[3] = mid + rad;
        }
        const bool pos = hi[0] > 0.0 && hi[1] >= 0.0 && hi[2] >= 0.0 && hi[3] >= 0.0;
        const bool neg = lo[0] < 0.0 && lo[1] <= 0.0 && lo[2] <= 0.0 && lo[3] <= 0.0;
        const bool flat = lo[0] <= 0.0 && hi[0] >= 0.0;     // a ray in the face plane somewhere in the cell
        const bool nan = !(lo[0] == lo[0]) || !(hi[3] == hi[3]);
        if (pos || neg || flat || nan) m |= 1u << j;
    }
    mask[cell] = m;
}


// one block: cursor[b] = number of entries in the bins before b; *total = all entries; the histogram is zeroed for
// the next launch
__global__ void __launch_bounds__(kBlock) k_mesh_scan(unsigned int *hist, unsigned int *cursor, uint32_t *total, int n_bins) {
    __shared__ unsigned int s_part[kBlock];
    const int per = (n_bins + kBlock - 1) / kBlock;
    const int b0 = threadIdx.x * per;
    unsigned int sum = 0;
    for (int k = 0; k < per; ++k) {
        const int b = b0 + k;
        if (b < n_bins) sum += hist[b];
    }
    s_part[threadIdx.x] = sum;
    __syncthreads();
    for (int d = 1; d < kBlock; d <<= 1) {          // inclusive Hillis-Steele scan of the per-thread sums
        const unsigned int v = (int)threadIdx.x >= d ? s_part[threadIdx.x - d] : 0u;
        __syncthreads();
        s_part[threadIdx.x] += v;
        __syncthreads();
    }
    unsigned int run = s_part[threadIdx.x] - sum;
    for (int k = 0; k < per; ++k) {
        const int b = b0 + k;
        if (b < n_bins) {
            const unsigned int h = hist[b];
            cursor[b] = run;
            run += h;
            hist[b] = 0u;
        }
    }
    if (threadIdx.x == kBlock - 1) *total = s_part[kBlock - 1];
}

// kScatterBatch consecutive regions per block and pass: ranks inside the batch by shared-memory atomics, one global
// atomic per bin that occurs in it
constexpr uint32_t kScatterBatch = 16;
__global__ void __launch_bounds__(kBlock) k_mesh_scatter(const uint32_t *__restrict__ entries, const uint16_t *__restrict__ bins,
                                                         const uint32_t *__restrict__ counts, const uint32_t n_regions,
                                                         const uint32_t cap, unsigned int *cursor, uint32_t *__restrict__ sorted,
                                                         const int n_bins) {
    extern __shared__ unsigned int s_sc[];      // count / fill per bin, then the batch's base per bin
    unsigned int *s_cnt = s_sc, *s_base = s_sc + n_bins;
    const uint32_t n_batches = (n_regions + kScatterBatch - 1) / kScatterBatch;
    for (uint32_t batch = blockIdx.x; batch < n_batches; batch += gridDim.x) {
        const uint32_t r0 = batch * kScatterBatch;
        const uint32_t r1 = r0 + kScatterBatch < n_regions ? r0 + kScatterBatch : n_regions;
        for (int i = threadIdx.x; i < n_bins; i += kBlock) s_cnt[i] = 0u;
        __syncthreads();
        for (uint32_t reg = r0; reg < r1; ++reg) {
            const uint32_t n = counts[reg];
            const uint64_t first = (uint64_t)reg * cap;
            for (uint32_t i = threadIdx.x; i < n; i += kBlock) atomicAdd(&s_cnt[bins[first + i]], 1u);
        }
        __syncthreads();
        for (int i = threadIdx.x; i < n_bins; i += kBlock) {
            const unsigned int cn = s_cnt[i];
            s_base[i] = cn ? atomicAdd(cursor + i, cn) : 0u;
            s_cnt[i] = 0u;
        }
        __syncthreads();
        for (uint32_t reg = r0; reg < r1; ++reg) {
            const uint32_t n = counts[reg];
            const uint64_t first = (uint64_t)reg * cap;
            for (uint32_t i = threadIdx.x; i < n; i += kBlock) {
                const unsigned int b = bins[first + i];
                sorted[s_base[b] + atomicAdd(&s_cnt[b], 1u)] = entries[first + i];
            }
        }
        __syncthreads();
    }
}
#endif  // XRT_MESHSORT_HOST_KERNELS

// Blocks of 128 threads, 5 resident per SM: 20 warps at 96 registers.  Measured on config 4 (rays/s): 256 x 2 at 128
// registers 1.00e10, 256 x 3 at 80 registers (224 B of stack) 1.06e10, 128 x 4 at 128 registers 1.04e10, 128 x 5 at 96
// registers 1.10e10 (mesh sphere +2.6 %, plasma source -> mesh -2.5 %).
#ifndef XRT_REFINE_BLOCKS
#define XRT_REFINE_BLOCKS 5
#endif
#ifndef XRT_REFINE_THREADS
#define XRT_REFINE_THREADS 128
#endif
constexpr int kRefineBlock = XRT_REFINE_THREADS;
template <uint32_t FT, bool HIST>
__global__ void __launch_bounds__(kRefineBlock, XRT_REFINE_BLOCKS)
k_mesh_refine(const __grid_constant__ XrtSceneDesc sc, const __grid_constant__ PhiloxKeys pk, const uint64_t stream_id,
              const uint64_t ray_begin, const uint32_t *__restrict__ sorted, const uint32_t *__restrict__ total,
              const XrtOutputs out, const int lazy_rt) {
    __shared__ unsigned long long s_cnt[XRT_MAX_OPTICS + 1];
    __shared__ double s_sincos[2 * kSincosTable];
    if (threadIdx.x <= XRT_MAX_OPTICS) s_cnt[threadIdx.x] = 0ull;
    for (int i = threadIdx.x; i < kSincosTable; i += kRefineBlock) {
        double sn, cs;
        sincos_2pi((double)i / (double)kSincosTable, sn, cs);
        s_sincos[2 * i] = cs;
        s_sincos[2 * i + 1] = sn;
    }
    __syncthreads();
    WarpCtx c;
    c.lane = threadIdx.x & 31u;
    c.lt_mask = (1u << c.lane) - 1u;
    c.s_cnt = s_cnt;
    const XrtOpticDesc &ops = sc.optics[0];
    const bool lazy = (lazy_rt & 1) != 0;
    const bool need_wave = (lazy_rt & 2) != 0;
    const int nopt = sc.n_optics;
    const uint32_t n = __ldg(total);
    const uint32_t n_groups = (n + 31u) / 32u;
    const uint32_t n_warps = gridDim.x * (kRefineBlock / 32);
    const double *coarse_geom = ops.mesh->coarse_geom;
    unsigned n_split = 0;
    for (uint32_t g = blockIdx.x * (kRefineBlock / 32) + (threadIdx.x >> 5); g < n_groups; g += n_warps) {
        const uint32_t idx = g * 32u + c.lane;
        const bool valid = idx < n;
        const uint32_t e = valid ? __ldg(sorted + idx) : 0u;
        const uint64_t id = ray_begin + (e >> kMeshFaceBits);
        const uint32_t face = e & ((1u << kMeshFaceBits) - 1u);
        PhiloxDraws dr;
        dr.init(pk, stream_id, id, 0);
        Ray r;
        r.alive = false;
        r.w = 0.0;
        if (valid) {
            SrcLocal L;
            source_local<FT, 0>(sc.source, id, L);
            generate_geometry<FT, PhiloxDraws, 0, true>(sc.source, L, dr, r, s_sincos);
            // a wavelength that depends on the direction at the source or on the bundle (Doppler shift, plasma) is drawn
            // with the ray, as in stage A of k_trace
            if (!lazy) r.w = generate_wavelength<PhiloxDraws, 0>(sc.source, L, dr, r.d);
            V3 o = r.o, d = r.d;
            if (optic_is_local<FT>(ops)) {
                o = to_local(ops.orient, o - v3(ops.origin));
                d = to_local(ops.orient, d);
            }
            const V3 Xc = mesh_face_point(coarse_geom + 9 * face, o, d);
            V3 nrm = v3(0.0, 0.0, 1.0);
            if (optic_geometry<FT, true, 0, true>(ops, r, nrm, nullptr, &Xc) == HIT_INSIDE) {
                // a lazy wavelength is drawn where it is first read (no Doppler shift)
                if (lazy && need_wave) r.w = generate_wavelength<PhiloxDraws, 0, false>(sc.source, L, dr, r.d);
                optic_interact<FT, PhiloxDraws, 0>(ops, 0, dr, r, nrm);
                if (r.alive && (ops.flags & XRT_F_IMAGE) && out.images) add_pixel(out, ops, r, c.lt_mask);
            }
        }
        n_split += __popc(__ballot_sync(kFull, r.alive));
        for (int k = 1; k < nopt; ++k) {
            const XrtOpticDesc &op = sc.optics[k];
            if (r.alive) {
                trace_optic<FT, PhiloxDraws, 0>(op, k, dr, r);       // inlined: the ray stays in registers
                if (r.alive && (op.flags & XRT_F_IMAGE) && out.images) add_pixel(out, op, r, c.lt_mask);
            }
            count_alive(c, k + 1, r.alive);
        }
        emit_found<HIST>(out, c.lane, c.lt_mask, r.alive, id);
        emit_lost<HIST>(out, c.lane, c.lt_mask, dr, valid && !r.alive, id);
    }
    if (c.lane == 0 && n_split) atomicAdd(&s_cnt[1], (unsigned long long)n_split);
    __syncthreads();
    if ((int)threadIdx.x <= nopt && threadIdx.x >= 1 && out.counts) {
        const unsigned long long cc = s_cnt[threadIdx.x];
        if (cc) atomicAdd((unsigned long long *)(out.counts + threadIdx.x), cc);
    }
}

template <uint32_t FT> __host__ __device__ constexpr size_t mesh_coarse_smem_bytes() {
    return 9 * kStageFaces * sizeof(double) + kMeshMaxBins * sizeof(unsigned int) + kDirGrid * kDirGrid * sizeof(uint32_t);
}

}  // namespace xrt
