// xrt_variants.h -- interface between the host side (xrt.cu) and the per-feature-set translation units (v_*.cu).
// Each v_*.cu instantiates k_trace / k_record for one compiled feature set, so the feature sets compile in parallel
// and none of them carries the code of the others.
#pragma once
#include "xrt_kernels.cuh"

namespace xrt {

typedef void (*TraceKernel)(const XrtSceneDesc, const PhiloxKeys, const uint64_t, const uint64_t, const uint64_t,
                            const XrtOutputs, const IdList, const int, const int);

typedef void (*CullKernel)(const Cull32Par, const XrtSourceDesc, const PhiloxKeys, const uint64_t, const uint64_t,
                           const uint64_t, const Cull32Out, const XrtOutputs);

struct RecordLaunch {
    const XrtSceneDesc *sc;
    PhiloxKeys pk;
    uint64_t stream_id;
    const uint64_t *ids;
    uint64_t ray_begin, n;
    XrtRaysIn in;
    XrtInject inj;
    XrtOutputs out;
    XrtHistory hist;
    int split, grid;
    cudaStream_t st;
};

// hist = false: history-off launch (no found / lost lists): kernels with that code compiled out where they exist
#define XRT_DECLARE_VARIANT(tag)                                                                   \
    TraceKernel trace_kernel_##tag(int split, uint32_t known, bool hist, size_t *smem);            \
    void record_launch_##tag(int mode, uint32_t known, const RecordLaunch &a);
XRT_DECLARE_VARIANT(lean)
XRT_DECLARE_VARIANT(mid)
XRT_DECLARE_VARIANT(mosaic)
XRT_DECLARE_VARIANT(src)
XRT_DECLARE_VARIANT(mesh)
XRT_DECLARE_VARIANT(full)

CullKernel cull_kernel(int src_mode, bool hist);

typedef void (*Mosaic32Kernel)(const Cull32Par, const Mosaic32Par, const XrtSourceDesc, const PhiloxKeys, const uint64_t,
                               const uint64_t, const uint64_t, const Cull32Out, const XrtOutputs);
Mosaic32Kernel mosaic32_kernel(int src_mode, bool hist);

// sorted mesh path: the coarse-mesh kernel of the two feature sets that hold mesh optics
typedef void (*MeshCoarseKernel)(const XrtSceneDesc, const PhiloxKeys, const uint64_t, const uint64_t, const uint64_t,
                                 const MeshSortOut, const XrtOutputs);
typedef void (*MeshRefineKernel)(const XrtSceneDesc, const PhiloxKeys, const uint64_t, const uint64_t, const uint32_t *,
                                 const uint32_t *, const XrtOutputs, const int);
MeshRefineKernel mesh_refine_kernel_mesh(bool hist);
MeshRefineKernel mesh_refine_kernel_full(bool hist);
MeshCoarseKernel mesh_coarse_kernel_mesh(bool hist, size_t *smem);
MeshCoarseKernel mesh_coarse_kernel_full(bool hist, size_t *smem);

// ---- helpers for the v_*.cu files
template <uint32_t FT, uint32_t KN = 0>
static void record_launch_ft(int mode, const RecordLaunch &a) {
    if (mode == REC_PHILOX)
        k_record<FT, REC_PHILOX, KN><<<a.grid, kBlock, 0, a.st>>>(*a.sc, a.pk, a.stream_id, a.ids, a.ray_begin, a.n, a.in, a.inj,
                                                                  a.out, a.hist, a.split);
    else
        k_record<FT, REC_INJECT, 0><<<a.grid, kBlock, 0, a.st>>>(*a.sc, a.pk, a.stream_id, a.ids, a.ray_begin, a.n, a.in, a.inj,
                                                                 a.out, a.hist, a.split);
}

// split index as a compile-time constant for 0 (and 1, 2 where WIDE); the history-off build exists for split 0
template <uint32_t FT, bool WIDE>
static TraceKernel trace_kernel_ft(int split, bool hist) {
    if (split == 0) return hist ? k_trace<FT, 0, 0, true> : k_trace<FT, 0, 0, false>;
    if constexpr (WIDE) {
        if (split == 1) return k_trace<FT, 1, 0, true>;
        if (split == 2) return k_trace<FT, 2, 0, true>;
    }
    return k_trace<FT, -1, 0, true>;
}

}  // namespace xrt
