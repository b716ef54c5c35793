// xrt.cu -- host side of libxrt.so: the C ABI of include/xrt.h (scene upload, launch configuration, entry points).
// The kernels are templates in xrt_kernels.cuh, instantiated per compiled feature set in v_*.cu.
#include <cuda_runtime.h>

#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <vector>

#define XRT_MESHSORT_HOST_KERNELS
#include "xrt_variants.h"
#include "xrt_plasma.cuh"
#include "xrt_select.cuh"

namespace xrt {

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
    // FP32 broad phase (k_cull32): -1 = does not apply to this scene, else CULL_*
    int cull_mode = -1;
    Cull32Par cull;
    // FP32 broad phase of a mosaic crystal's crystallite scan (k_mosaic32): -1 = does not apply, else CULL_*
    int mosaic32_mode = -1;
    Mosaic32Par mosaic32;
    uint32_t *list_ids = nullptr;   // id list between k_cull32 and k_trace (stream-ordered allocation, grown on demand)
    uint32_t *list_counts = nullptr;
    uint64_t list_ids_cap = 0, list_counts_cap = 0;
    unsigned int *list_next = nullptr;   // [2] region counters of k_cull32, used alternately (each launch resets the other)
    int list_phase = 0;
    // sorted mesh path (xrt_meshsort.cuh): applies when the first optic is a refining mesh (see scene_build)
    int mesh_sort = 0;
    int mesh_bins = 0, mesh_tile = 0, mesh_tiles_x = 0, mesh_sub = 1;
    uint32_t *ms_entries = nullptr, *ms_sorted = nullptr, *ms_counts = nullptr, *ms_total = nullptr;
    uint16_t *ms_bins = nullptr;
    unsigned int *ms_hist = nullptr, *ms_cursor = nullptr;
    uint64_t ms_cap = 0, ms_counts_cap = 0;
    const double *coarse_geom_dev = nullptr;   // of the mesh uploaded last (upload_mesh)
    int32_t coarse_faces = 0;
    MeshDirGrid dirgrid = {};                  // direction grid of a point source in front of the mesh (mask = nullptr: none)
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
    int32_t n_items = 0, n_vitems = 0, n_nb = 0;
    if (cells && m.grid_start) n_items = m.grid_start[cells];
    if (cells && m.vgrid_start) n_vitems = m.vgrid_start[cells];
    if (cells && m.nb_start) n_nb = m.nb_start[cells];
    if (!m.nb_start || !m.nb_rec) { m.nb_start = nullptr; m.nb_rec = nullptr; }
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
    UP(m.nb_start, (cells && m.nb_start) ? cells + 1 : 0);
    UP(m.nb_rec, (cells && m.nb_start) ? 4 * (size_t)n_nb : 0);
    UP(m.tri_rec, (cells && m.grid_start && m.tri_rec) ? 8 * (size_t)n_items : 0);
    UP(m.vertex_face_rec, m.vertex_face_rec ? 128 * (size_t)m.n_points : 0);
    {
        const size_t fcells = (size_t)m.fgrid_nx * (size_t)m.fgrid_ny;
        const bool have = fcells > 0 && m.fgrid_start && m.fgrid_items;
        const int32_t n_f = have ? m.fgrid_start[fcells] : 0;
        if (!have) { m.fgrid_start = nullptr; m.fgrid_items = nullptr; m.fgrid_nx = m.fgrid_ny = 0; }
        UP(m.fgrid_start, have ? fcells + 1 : 0);
        UP(m.fgrid_items, n_f);
    }
    s->coarse_geom_dev = m.coarse_geom;
    s->coarse_faces = m.n_coarse_faces;
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
    if (s->list_ids) cudaFreeAsync(s->list_ids, (cudaStream_t)0);
    if (s->list_counts) cudaFreeAsync(s->list_counts, (cudaStream_t)0);
    if (s->list_next) cudaFreeAsync(s->list_next, (cudaStream_t)0);
    // ms_cursor and ms_total live inside the ms_hist allocation
    for (void *p : {(void *)s->ms_entries, (void *)s->ms_sorted, (void *)s->ms_counts, (void *)s->ms_bins, (void *)s->ms_hist})
        if (p) cudaFreeAsync(p, (cudaStream_t)0);
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
            if (k == 0 && (m.flags & XRT_F_MESH_REFINE) && !(m.flags & XRT_F_MESH_LOSSLESS) &&
                m.mesh->n_coarse_faces <= (1 << kMeshFaceBits) && m.mesh->grid_nx > 0 && m.mesh->grid_ny > 0) {
                // bins of the sorted path: tiles of the vertex grid, 3 x 3 cells unless that gives too many bins
                // bins of the sorted path: the cells of the vertex grid, each cut in sub x sub (rays of one bin share the
                // nearest vertex and the triangle: warp-uniform table reads), or tiles of tile x tile cells when the
                // grid has more cells than bins allowed
                int sub = 2, tile = 1;
                if (const char *v = std::getenv("XRT_MESH_SUB")) sub = std::atoi(v) > 0 ? std::atoi(v) : 2;       // measurement knobs
                if (const char *v = std::getenv("XRT_MESH_TILE")) tile = std::atoi(v) > 0 ? std::atoi(v) : 1;
                const long gx = m.mesh->grid_nx, gy = m.mesh->grid_ny;
                while (sub > 1 && gx * sub * gy * sub > kMeshMaxBins) --sub;
                while (((gx * sub + tile - 1) / tile) * ((gy * sub + tile - 1) / tile) > kMeshMaxBins) ++tile;
                s->mesh_sub = sub;
                s->mesh_tile = tile;
                s->mesh_tiles_x = (int)((gx * sub + tile - 1) / tile);
                s->mesh_bins = s->mesh_tiles_x * (int)((gy * sub + tile - 1) / tile);
            }
            int rc = upload_mesh(s, m.mesh, &m.mesh);
            if (rc != XRT_OK) return rc;
        } else {
            m.mesh = nullptr;
        }
    }
    for (int k = 0; k < d.n_optics; ++k) {
        d.optics[k].cull_t2 = d.optics[k].cull_err = d.optics[k].cull_inv_r = 0.0;
        d.optics[k].mosaic_scan = 0;
        d.optics[k].mosaic_t2 = d.optics[k].mosaic_err = 0.0;
    }
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
    // sorted mesh path: the conditions under which k_trace splits stage A at the coarse mesh (mesh_staged)
    // (any source kind, plasma bundles included: both kernels rebuild the ray from its id; a wavelength that depends on
    // the source direction or the bundle is drawn with the ray in k_mesh_refine)
    s->mesh_sort = (s->mesh_bins > 0 && s->split == 0 && d.optics[0].shape == XRT_SHAPE_MESH &&
                    src.cone != XRT_CONE_ISOTROPIC_XY &&
                    (s->features == FT_MESHLEAN || s->features == FT_FULL)) ? 1 : 0;
    if (s->mesh_sort && src.kind == XRT_SRC_FIXED_AXIS && src.cone == XRT_CONE_ISOTROPIC && src.spatial == XRT_SPATIAL_UNIFORM &&
        src.extent[0] == 0.0 && src.extent[1] == 0.0 && src.extent[2] == 0.0 && src.cone_par[0] > 0.2 &&
        src.cone_par[0] < 1.0 && s->coarse_geom_dev && !getenv("XRT_NO_MESH_DIRGRID")) {
        // direction grid over the cone (xrt_meshsort.cuh): cone frame and source point in the mesh's tracing frame
        const XrtOpticDesc &o = d.optics[0];
        const bool local = (o.flags & XRT_F_TRACE_LOCAL) != 0;
        auto rot = [&](const double *v, double *out) {
            for (int i = 0; i < 3; ++i)
                out[i] = local ? o.orient[3 * i] * v[0] + o.orient[3 * i + 1] * v[1] + o.orient[3 * i + 2] * v[2] : v[i];
        };
        MeshDirGrid &G = s->dirgrid;
        rot(src.axis_basis + 0, G.ex);
        rot(src.axis_basis + 3, G.ey);
        rot(src.axis_basis + 6, G.ez);
        const double cs = src.cone_par[0];
        G.half = std::sqrt(1.0 - cs * cs) / cs * (1.0 + 1e-9);
        G.inv_h = kDirGrid / (2.0 * G.half);
        double rel[3] = {src.origin[0] - (local ? o.origin[0] : 0.0), src.origin[1] - (local ? o.origin[1] : 0.0),
                         src.origin[2] - (local ? o.origin[2] : 0.0)}, ot[3];
        rot(rel, ot);
        void *p = nullptr;
        CU(cudaMallocAsync(&p, kDirGrid * kDirGrid * sizeof(uint32_t), (cudaStream_t)0));
        s->allocs.push_back(p);
        G.mask = (const uint32_t *)p;
        k_mesh_dirgrid<<<(kDirGrid * kDirGrid + kBlock - 1) / kBlock, kBlock, 0, (cudaStream_t)0>>>(
            s->coarse_geom_dev, s->coarse_faces, V3{ot[0], ot[1], ot[2]}, G, (uint32_t *)p);
        CU(cudaGetLastError());
    }
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
        // parameters of the FP32 broad phases (k_cull32 for a Bragg crystal, k_mosaic32 for a mosaic crystal): first optic,
        // isotropic cone from a point / uniform box / plasma voxel, constant or normal line.  Returns the CULL_* source
        // kind, or -1 when the phase does not apply; fills s->cull.
        auto fill_cull32 = [&](double t2) -> int {
            const bool line_ok = src.wave == XRT_WAVE_NORMAL || src.wave == XRT_WAVE_CONST;
            if (!(s->split == 0 && line_ok && src.cone == XRT_CONE_ISOTROPIC && src.spatial == XRT_SPATIAL_UNIFORM &&
                  src.n_sightlines == 0 && std::fabs(src.wave_par[0] * o.inv_two_d) >= 0.1 &&
                  std::getenv("XRT_NO_BROAD32") == nullptr))
                return -1;
            Cull32Par &K = s->cull;
            std::memset(&K, 0, sizeof(K));
            const double r2 = o.radius * o.radius;
            const bool point = (s->known & KN_POINT_SOURCE) != 0;
            const int mode = src.kind == XRT_SRC_BUNDLES ? CULL_BUNDLES
                           : src.kind == XRT_SRC_FOCUSED ? CULL_FOCUSED : (point ? CULL_POINT : CULL_BOX);
            // farthest source point from the centre of curvature (box corners); bundles are tested per ray
            double ll_max = 0.0;
            for (int c8 = 0; c8 < 8; ++c8) {
                double p[3];
                for (int a = 0; a < 3; ++a) {
                    p[a] = src.origin[a];
                    for (int e = 0; e < 3; ++e)
                        p[a] += (((c8 >> e) & 1) ? 0.5 : -0.5) * src.extent[e] * src.orient[3 * e + a];
                }
                const double dx = o.center[0] - p[0], dy = o.center[1] - p[1], dz = o.center[2] - p[2];
                ll_max = std::fmax(ll_max, dx * dx + dy * dy + dz * dz);
            }
            if (mode == CULL_BUNDLES || ll_max <= 4.0 * r2) {
                K.one_m_cos = (float)(1.0 - src.cone_par[0]);
                for (int i = 0; i < 9; ++i) K.basis[i] = (float)src.axis_basis[i];
                for (int i = 0; i < 9; ++i) K.R[i] = (float)src.orient[i];
                for (int i = 0; i < 3; ++i) {
                    K.Lb[i] = (float)(o.center[i] - src.origin[i]);
                    K.Tb[i] = (float)(src.target[i] - src.origin[i]);
                    K.ext[i] = (float)src.extent[i];
                    K.xz[i] = (float)(src.orient[i] + src.orient[6 + i]);
                    K.vel[i] = (float)src.velocity_c[i];
                    K.C[i] = o.center[i];
                    K.T[i] = src.target[i];
                }
                for (int i = 0; i < 3; ++i) {      // point source: basis rows dotted with C - O and with v / c
                    double mi = 0.0, mvi = 0.0;
                    for (int a = 0; a < 3; ++a) {
                        mi += src.axis_basis[3 * i + a] * (o.center[a] - src.origin[a]);
                        mvi += src.axis_basis[3 * i + a] * src.velocity_c[a];
                    }
                    K.m[i] = (float)mi;
                    K.mv[i] = (float)mvi;
                }
                K.ll = (float)ll_max;              // point source: the one value of |C - O|^2
                K.r2 = (float)r2;
                K.inv_r2 = (float)(1.0 / r2);
                K.inv_r = (float)(1.0 / o.radius);
                K.lam0 = (float)src.wave_par[0];
                K.normal_line = src.wave == XRT_WAVE_NORMAL ? 1 : 0;
                K.sig = K.normal_line ? (float)src.wave_par[1] : 0.0f;
                K.inv_two_d = (float)o.inv_two_d;
                K.t2 = (float)t2;
                // approximate-deviate term: 2e-3 sigma / 2d (per bundle for a plasma); rounding of the exact path 1e-9
                const double dev = K.normal_line ? 2e-3 * std::fabs(src.wave_par[1]) * std::fabs(o.inv_two_d) : 0.0;
                K.err_sig = K.normal_line ? (float)(2e-3 * std::fabs(o.inv_two_d)) : 0.0f;
                K.err = (float)((mode == CULL_BUNDLES ? 0.0 : dev) + 1e-9 +
                                (mode == CULL_POINT ? 2e-5 * std::fmax(1.0, ll_max / r2) : 0.0));
                K.moving = (src.velocity_c[0] != 0.0 || src.velocity_c[1] != 0.0 || src.velocity_c[2] != 0.0) ? 1 : 0;
                // second stage: crystal bounds (a ray outside |x| < hx, |y| < hy is lost at the crystal whatever else
                // the optic checks) and the second pre-test level with the rocking-curve uniform
                const uint32_t xy = XRT_F_CHECK_SIZE | XRT_F_HAS_XSIZE | XRT_F_HAS_YSIZE;
                K.bounds_xy = ((o.flags & xy) == xy) ? 1 : 0;
                K.convex = (o.flags & XRT_F_CONVEX) ? 1 : 0;
                K.gauss = o.rocking_type == XRT_ROCK_GAUSS ? 1 : 0;
                for (int i = 0; i < 3; ++i) {
                    K.Ob[i] = (float)(src.origin[i] - o.origin[i]);
                    K.ox[i] = (float)o.orient[i];
                    K.oy[i] = (float)o.orient[3 + i];
                    K.Oc[i] = o.origin[i];
                }
                K.hx = (float)o.half_size[0];
                K.hy = (float)o.half_size[1];
                K.lg_refl = (float)std::log2(o.reflectivity);
                K.two_sigma2 = (float)o.rock_two_sigma2;
                // (a step curve has no second level; the bounds test alone does not pay for the re-pack: measured)
                K.stage2 = (K.gauss && std::getenv("XRT_NO_STAGE2") == nullptr) ? 1 : 0;
                return mode;
            }
            return -1;
        };
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
            // FP32 broad phase (k_cull32, xrt_kernels.cuh): first optic, isotropic cone from a point / uniform box /
            // plasma voxel, constant or normal line.  Its arithmetic error on sin(theta_i) is about 1e-6 for
            // |C - O| ~ R; the margin is 2e-5, scaled with |C - O|^2 / R^2 (per ray where the origin varies), and
            // the phase is left off for a source farther than 2 R from the centre of curvature and for Bragg angles
            // below 6 degrees, where thc -> sI amplifies the error of thc^2 by 1 / (2 sI).
            const int mode32 = fill_cull32(w.cull_t2);
            if (mode32 >= 0) {
                s->cull_mode = mode32;
                d.kn32[0] = 1.0f;       // reported to the caller: the broad phase is in use
            }
            // (mosaic crystals: below)
            // eager normal line (plasma bundles, Doppler shift) with no optic before the crystal: defer the exact deviate
            s->defer_wavelength = (!s->lazy_wavelength && src.wave == XRT_WAVE_NORMAL && s->split == 0) ? 1 : 0;
        }
        // mosaic crystal as split optic: per-layer FP32 pre-test of the crystallite loop (stage_mosaic in xrt_kernels.cuh).
        // Analytic shape traced in global coordinates, Bragg test on, Gaussian or step rocking curve; the layer index
        // travels in the top byte of the queue's id word.
        if (o.interact == XRT_INTERACT_MOSAIC && (o.flags & XRT_F_CHECK_BRAGG) && rock_ok && !(o.flags & XRT_F_TRACE_LOCAL) &&
            o.shape != XRT_SHAPE_MESH && o.mosaic_depth >= 1 && o.mosaic_depth <= 127 && std::isfinite(o.inv_two_d) &&
            std::isfinite(o.mosaic_sin_sigma) && std::getenv("XRT_NO_CULL") == nullptr) {
            XrtOpticDesc &w = d.optics[s->split];
            const double edge = o.rocking_type == XRT_ROCK_GAUSS ? std::sqrt(40.0 / o.rock_inv_two_sigma2) : 0.5 * o.rocking_fwhm;
            const double t = 1.05 * edge + 2e-6;
            w.mosaic_t2 = t * t;
            w.mosaic_err = 2e-6;
            w.mosaic_scan = 1;
            // FP32 broad phase of the crystallite scan (k_mosaic32): concave or convex sphere as first optic, the
            // source conditions of k_cull32 (no plasma bundles), lean mosaic feature set, no cutoff prefilter, wavelength
            // not needed before the crystal
            if (o.shape == XRT_SHAPE_SPHERE && s->features == FT_MOSAICLEAN && o.mosaic_depth <= (1 << kMosaicTagBits) &&
                !(o.flags & XRT_F_MOSAIC_CUTOFF) && s->lazy_wavelength && wave_ok && o.radius > 0.0 &&
                src.kind != XRT_SRC_BUNDLES && std::getenv("XRT_NO_MOSAIC32") == nullptr && fill_cull32(0.0) >= 0) {
                s->mosaic32_mode = fill_cull32(0.0);
                Mosaic32Par &M = s->mosaic32;
                M.sin_sigma = (float)o.mosaic_sin_sigma;
                M.err = (float)w.mosaic_err;
                M.t2 = (float)w.mosaic_t2;
                M.two_sigma2 = (float)o.rock_two_sigma2;
                M.lg_refl = (float)std::log2(o.reflectivity);
                M.depth = o.mosaic_depth;
                M.gauss = o.rocking_type == XRT_ROCK_GAUSS ? 1 : 0;
            }
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

static TraceKernel trace_kernel(const XrtScene *s, bool hist, size_t *smem) {
    switch (s->features) {
    case 0: return trace_kernel_lean(s->split, s->known, hist, smem);
    case FT_MID: return trace_kernel_mid(s->split, s->known, hist, smem);
    case FT_MOSAICLEAN: return trace_kernel_mosaic(s->split, s->known, hist, smem);
    case FT_SRCLEAN: return trace_kernel_src(s->split, s->known, hist, smem);
    case FT_MESHLEAN: return trace_kernel_mesh(s->split, s->known, hist, smem);
    default: return trace_kernel_full(s->split, s->known, hist, smem);
    }
}

static int trace_launch_config(const XrtScene *s, bool hist, TraceKernel *kern, size_t *smem, int *blocks_per_sm, int *regs) {
    *kern = trace_kernel(s, hist, smem);
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

// launches shorter than this run the single-kernel path (the broad phase pays off once its regions fill warps)
static uint64_t cull_min_rays() {
    if (const char *v = std::getenv("XRT_CULL32_MIN_RAYS")) return (uint64_t)std::strtoull(v, nullptr, 10);
    return 1ull << 21;
}

extern "C" int xrt_launch_info(XrtScene *s, int32_t *grid, int32_t *block, int32_t *regs, int32_t *blocks_per_sm) {
    if (!s) return fail(XRT_EINVAL, "null scene");
    TraceKernel kern;
    size_t smem;
    int bps = 0, r = 0;
    int rc = trace_launch_config(s, false, &kern, &smem, &bps, &r);
    if (rc != XRT_OK) return rc;
    if (grid) *grid = s->sm_count * bps;
    if (block) *block = kBlock;
    if (regs) *regs = r;
    if (blocks_per_sm) *blocks_per_sm = bps;
    return XRT_OK;
}

extern "C" int xrt_launch_info_cull(XrtScene *s, int32_t *mode, int32_t *grid, int32_t *regs, int32_t *blocks_per_sm) {
    if (!s) return fail(XRT_EINVAL, "null scene");
    if (s->cull_mode < 0 && s->mesh_sort && !getenv("XRT_NO_MESH_SORT")) {
        // sorted mesh path: mode 4 = k_mesh_coarse in front of k_mesh_refine; with bins = the number of spatial bins
        size_t csmem = 0;
        MeshCoarseKernel mk = s->features == FT_MESHLEAN ? mesh_coarse_kernel_mesh(false, &csmem) : mesh_coarse_kernel_full(false, &csmem);
        MeshRefineKernel rk = s->features == FT_MESHLEAN ? mesh_refine_kernel_mesh(false) : mesh_refine_kernel_full(false);
        cudaFuncAttributes fa, fr;
        CU(cudaFuncGetAttributes(&fa, mk));
        CU(cudaFuncGetAttributes(&fr, rk));
        int bps = 0;
        CU(cudaFuncSetAttribute(mk, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)csmem));
        CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&bps, mk, kBlock, csmem));
        if (mode) *mode = 4;
        if (grid) *grid = s->sm_count * bps;
        if (regs) *regs = fa.numRegs | (fr.numRegs << 16);     // refinement kernel's registers in the high half
        if (blocks_per_sm) *blocks_per_sm = bps | (s->mesh_bins << 8);
        return XRT_OK;
    }
    if (s->cull_mode < 0 && s->mosaic32_mode >= 0) {
        // mode 5: the mosaic broad phase k_mosaic32 in front of k_trace
        Mosaic32Kernel mk = mosaic32_kernel(s->mosaic32_mode, false);
        cudaFuncAttributes fa;
        CU(cudaFuncGetAttributes(&fa, mk));
        int bps = 0;
        CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&bps, mk, kBlock, 0));
        if (mode) *mode = 5;
        if (grid) *grid = s->sm_count * bps;
        if (regs) *regs = fa.numRegs;
        if (blocks_per_sm) *blocks_per_sm = bps;
        return XRT_OK;
    }
    if (mode) *mode = s->cull_mode;
    if (s->cull_mode < 0) {
        if (grid) *grid = 0;
        if (regs) *regs = 0;
        if (blocks_per_sm) *blocks_per_sm = 0;
        return XRT_OK;
    }
    CullKernel ck = cull_kernel(s->cull_mode, false);
    cudaFuncAttributes fa;
    CU(cudaFuncGetAttributes(&fa, ck));
    int bps = 0;
    CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&bps, ck, kBlock, 0));
    if (grid) *grid = s->sm_count * bps;
    if (regs) *regs = fa.numRegs;
    if (blocks_per_sm) *blocks_per_sm = bps;
    return XRT_OK;
}

static int ensure_list(XrtScene *s, uint64_t n_ids, uint64_t n_regions, cudaStream_t st) {
    if (!s->list_next) {
        void *p = nullptr;
        CU(cudaMallocAsync(&p, 2 * sizeof(unsigned int), st));
        CU(cudaMemsetAsync(p, 0, 2 * sizeof(unsigned int), st));
        s->list_next = (unsigned int *)p;
        s->list_phase = 0;
    }
    // a plasma draws a new Poisson total every iteration: headroom, so that the list is not re-allocated (gigabytes
    // from the pool) whenever the total grows by a few rays
    if (s->list_ids_cap < n_ids) {
        n_ids += n_ids / 16 + (1u << 20);
        n_regions += n_regions / 16 + 1024;
        if (s->list_ids) CU(cudaFreeAsync(s->list_ids, st));
        s->list_ids = nullptr;
        s->list_ids_cap = 0;
        void *p = nullptr;
        CU(cudaMallocAsync(&p, n_ids * sizeof(uint32_t), st));
        s->list_ids = (uint32_t *)p;
        s->list_ids_cap = n_ids;
    }
    if (s->list_counts_cap < n_regions) {
        if (s->list_counts) CU(cudaFreeAsync(s->list_counts, st));
        s->list_counts = nullptr;
        s->list_counts_cap = 0;
        void *p = nullptr;
        CU(cudaMallocAsync(&p, n_regions * sizeof(uint32_t), st));
        s->list_counts = (uint32_t *)p;
        s->list_counts_cap = n_regions;
    }
    return XRT_OK;
}

static uint64_t mesh_sort_min_rays() {
    if (const char *v = std::getenv("XRT_MESH_SORT_MIN_RAYS")) return (uint64_t)std::strtoull(v, nullptr, 10);
    return 1ull << 21;
}

static int ensure_mesh_sort(XrtScene *s, uint64_t n_ids, uint64_t n_regions, cudaStream_t st) {
    if (!s->list_next) {
        void *p = nullptr;
        CU(cudaMallocAsync(&p, 2 * sizeof(unsigned int), st));
        CU(cudaMemsetAsync(p, 0, 2 * sizeof(unsigned int), st));
        s->list_next = (unsigned int *)p;
        s->list_phase = 0;
    }
    if (!s->ms_hist) {
        void *p = nullptr;
        CU(cudaMallocAsync(&p, (2 * (size_t)kMeshMaxBins + 1) * sizeof(unsigned int), st));
        CU(cudaMemsetAsync(p, 0, (2 * (size_t)kMeshMaxBins + 1) * sizeof(unsigned int), st));
        s->ms_hist = (unsigned int *)p;
        s->ms_cursor = s->ms_hist + kMeshMaxBins;
        s->ms_total = (uint32_t *)(s->ms_hist + 2 * kMeshMaxBins);
    }
    if (s->ms_cap < n_ids) {
        for (void *p : {(void *)s->ms_entries, (void *)s->ms_sorted, (void *)s->ms_bins})
            if (p) CU(cudaFreeAsync(p, st));
        s->ms_entries = s->ms_sorted = nullptr;
        s->ms_bins = nullptr;
        s->ms_cap = 0;
        void *p = nullptr;
        CU(cudaMallocAsync(&p, n_ids * sizeof(uint32_t), st));
        s->ms_entries = (uint32_t *)p;
        CU(cudaMallocAsync(&p, n_ids * sizeof(uint32_t), st));
        s->ms_sorted = (uint32_t *)p;
        CU(cudaMallocAsync(&p, n_ids * sizeof(uint16_t), st));
        s->ms_bins = (uint16_t *)p;
        s->ms_cap = n_ids;
    }
    if (s->ms_counts_cap < n_regions) {
        if (s->ms_counts) CU(cudaFreeAsync(s->ms_counts, st));
        s->ms_counts = nullptr;
        s->ms_counts_cap = 0;
        void *p = nullptr;
        CU(cudaMallocAsync(&p, n_regions * sizeof(uint32_t), st));
        s->ms_counts = (uint32_t *)p;
        s->ms_counts_cap = n_regions;
    }
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
    const bool hist = out->found_count != nullptr || out->lost_count != nullptr || out->found_bits != nullptr ||
                      out->lost_bits != nullptr;
    TraceKernel kern;
    size_t smem;
    int bps = 0;
    int rc = trace_launch_config(s, hist, &kern, &smem, &bps, nullptr);
    if (rc != XRT_OK) return rc;
    // one resident wave of blocks, each warp strides over the id regions
    if (const char *lim = getenv("XRT_BLOCKS_PER_SM")) {      // measurement knob: fewer resident blocks
        int v = atoi(lim);
        if (v >= 1 && v < bps) bps = v;
    }
    cudaStream_t st = (cudaStream_t)stream;
    PhiloxKeys pk;
    philox_round_keys(seed, stream_id, pk);
    const int lazy_bits = s->lazy_wavelength | (s->need_wavelength << 1) | (s->defer_wavelength << 2);
    const uint64_t cap_blocks = (uint64_t)s->sm_count * (uint64_t)bps;

    const bool mosaic32 = s->mosaic32_mode >= 0 && ray_count >= cull_min_rays();
    const bool two_kernels = (s->cull_mode >= 0 && ray_count >= cull_min_rays()) || mosaic32;
    const bool mesh_sorted = !two_kernels && s->mesh_sort && ray_count >= mesh_sort_min_rays() && !getenv("XRT_NO_MESH_SORT");
    // a launch covers at most 2^30 ids (32-bit offsets in the id list, 4 GB of list at most); 2^27 on the sorted mesh
    // path (27-bit offsets beside the face tag)
    const uint64_t max_launch = mosaic32 ? (1ull << (32 - kMosaicTagBits)) : two_kernels ? (1ull << 30) : (mesh_sorted ? kMeshMaxLaunch : ~0ull);
    for (uint64_t done = 0; done < ray_count; done += max_launch) {
        const uint64_t n = ray_count - done < max_launch ? ray_count - done : max_launch;
        const uint64_t begin = ray_begin + done;
        IdList list;
        if (two_kernels) {
            // regions of consecutive ids: a multiple of 32 ids each, about 384 regions per SM (a multiple of the
            // warps per SM of both kernels for 1, 2, 3, 4, 6 or 8 resident blocks), at least 32 warp passes each
            const uint64_t n_groups = (n + 31) / 32;
            uint64_t gpr = (n_groups + (uint64_t)s->sm_count * 384 - 1) / ((uint64_t)s->sm_count * 384);
            if (gpr < 32) gpr = 32;
            if (mosaic32 && gpr < 128) gpr = 128;   // the scan of a region is drained at its end: long regions
            gpr = (gpr + 3) & ~3ull;            // a multiple of the broad phase's groups per pass: only the last region is ragged
            const uint32_t cap = (uint32_t)(gpr * 32);
            const uint32_t n_regions = (uint32_t)((n + cap - 1) / cap);
            rc = ensure_list(s, (uint64_t)n_regions * cap, n_regions, st);
            if (rc != XRT_OK) return rc;
            if (mosaic32) {
                Mosaic32Kernel mk = mosaic32_kernel(s->mosaic32_mode, hist);
                int mbps = 0;
                CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&mbps, mk, kBlock, 0));
                if (mbps < 1) return fail(XRT_ECUDA, "mosaic broad-phase kernel does not fit on an SM");
                const uint64_t mwant = ((uint64_t)n_regions + kBlock / 32 - 1) / (kBlock / 32);
                const uint64_t mcap = (uint64_t)s->sm_count * (uint64_t)mbps;
                Cull32Out lst = {s->list_ids, s->list_counts, n_regions, cap, s->list_next + s->list_phase,
                                 s->list_next + (s->list_phase ^ 1)};
                s->list_phase ^= 1;
                mk<<<(int)(mwant < mcap ? mwant : mcap), kBlock, 0, st>>>(s->cull, s->mosaic32, s->dev.source, pk, stream_id, begin, n, lst, *out);
                CU(cudaGetLastError());
                list = {s->list_ids, s->list_counts, n_regions, cap, (uint32_t)kMosaicTagBits};
                const uint64_t want = ((uint64_t)list.n_regions + kBlock / 32 - 1) / (kBlock / 32);
                kern<<<(int)(want < cap_blocks ? want : cap_blocks), kBlock, smem, st>>>(s->dev, pk, stream_id, begin, n, *out, list, s->split, lazy_bits);
                CU(cudaGetLastError());
                continue;
            }
            CullKernel ck = cull_kernel(s->cull_mode, hist);
            int cbps = 0;
            CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&cbps, ck, kBlock, 0));
            if (cbps < 1) return fail(XRT_ECUDA, "broad-phase kernel does not fit on an SM");
            const uint64_t want = ((uint64_t)n_regions + kBlock / 32 - 1) / (kBlock / 32);
            const uint64_t ccap = (uint64_t)s->sm_count * (uint64_t)cbps;
            // two counters used alternately: this launch claims regions from one and zeroes the other for the next launch
            Cull32Out lst = {s->list_ids, s->list_counts, n_regions, cap, s->list_next + s->list_phase,
                             s->list_next + (s->list_phase ^ 1)};
            s->list_phase ^= 1;
            ck<<<(int)(want < ccap ? want : ccap), kBlock, 0, st>>>(s->cull, s->dev.source, pk, stream_id, begin, n, lst, *out);
            CU(cudaGetLastError());
            list = {s->list_ids, s->list_counts, n_regions, cap, 0u};
        } else if (mesh_sorted) {
            // regions of consecutive ids for k_mesh_coarse, as for the broad phase
            const uint64_t n_groups = (n + 31) / 32;
            uint64_t gpr = (n_groups + (uint64_t)s->sm_count * 384 - 1) / ((uint64_t)s->sm_count * 384);
            if (gpr < 32) gpr = 32;
            const uint32_t cap = (uint32_t)(gpr * 32);
            const uint32_t n_regions = (uint32_t)((n + cap - 1) / cap);
            rc = ensure_mesh_sort(s, (uint64_t)n_regions * cap, n_regions, st);
            if (rc != XRT_OK) return rc;
            size_t csmem = 0;
            MeshCoarseKernel mk = s->features == FT_MESHLEAN ? mesh_coarse_kernel_mesh(hist, &csmem) : mesh_coarse_kernel_full(hist, &csmem);
            int cbps = 0;
            CU(cudaFuncSetAttribute(mk, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)csmem));
            CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&cbps, mk, kBlock, csmem));
            if (cbps < 1) return fail(XRT_ECUDA, "coarse-mesh kernel does not fit on an SM");
            const uint64_t cwant = ((uint64_t)n_regions + kBlock / 32 - 1) / (kBlock / 32);
            const uint64_t ccap = (uint64_t)s->sm_count * (uint64_t)cbps;
            MeshSortOut lst = {s->ms_entries, s->ms_bins, s->ms_counts, n_regions, cap, s->list_next + s->list_phase,
                               s->list_next + (s->list_phase ^ 1), s->ms_hist, s->mesh_bins, s->mesh_tile, s->mesh_tiles_x, s->mesh_sub, s->dirgrid};
            s->list_phase ^= 1;
            mk<<<(int)(cwant < ccap ? cwant : ccap), kBlock, csmem, st>>>(s->dev, pk, stream_id, begin, n, lst, *out);
            CU(cudaGetLastError());
            k_mesh_scan<<<1, kBlock, 0, st>>>(s->ms_hist, s->ms_cursor, s->ms_total, s->mesh_bins);
            CU(cudaGetLastError());
            const uint64_t sgrid = (uint64_t)s->sm_count * 8, n_batches = (n_regions + kScatterBatch - 1) / kScatterBatch;
            CU(cudaFuncSetAttribute(k_mesh_scatter, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(2 * kMeshMaxBins * sizeof(unsigned int))));
            k_mesh_scatter<<<(int)(n_batches < sgrid ? n_batches : sgrid), kBlock, 2 * (size_t)s->mesh_bins * sizeof(unsigned int), st>>>(
                s->ms_entries, s->ms_bins, s->ms_counts, n_regions, cap, s->ms_cursor, s->ms_sorted, s->mesh_bins);
            CU(cudaGetLastError());
            // the number of hits is known on the device only: one resident wave of the refinement kernel reads *total
            MeshRefineKernel rk = s->features == FT_MESHLEAN ? mesh_refine_kernel_mesh(hist) : mesh_refine_kernel_full(hist);
            int rbps = 0;
            // 4 kB of static shared memory per block: a small carve-out, the rest of the SM's 256 kB is L1
            CU(cudaFuncSetAttribute(rk, cudaFuncAttributePreferredSharedMemoryCarveout, 12));
            CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&rbps, rk, kRefineBlock, 0));
            if (rbps < 1) return fail(XRT_ECUDA, "mesh refinement kernel does not fit on an SM");
            rk<<<s->sm_count * rbps, kRefineBlock, 0, st>>>(s->dev, pk, stream_id, begin, s->ms_sorted, s->ms_total, *out, lazy_bits);
            CU(cudaGetLastError());
            continue;
        } else {
            const uint64_t n_groups = (n + 31) / 32;
            if (n_groups > 0xffffffffull) return fail(XRT_EINVAL, "ray_count too large for one launch");
            list = {nullptr, nullptr, (uint32_t)n_groups, 32u, 0u};
        }
        const uint64_t want = ((uint64_t)list.n_regions + kBlock / 32 - 1) / (kBlock / 32);
        const int grid = (int)(want < cap_blocks ? want : cap_blocks);
        kern<<<grid, kBlock, smem, st>>>(s->dev, pk, stream_id, begin, n, *out, list, s->split, lazy_bits);
        CU(cudaGetLastError());
    }
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
    RecordLaunch a;
    a.sc = &s->dev;
    philox_round_keys(seed, stream_id, a.pk);
    a.stream_id = stream_id;
    a.ids = ids;
    a.ray_begin = ray_begin;
    a.n = n;
    a.in = in;
    a.inj = inj;
    a.out = out;
    a.hist = hist;
    a.split = s->split;
    a.grid = (int)(want < cap ? want : cap);
    a.st = (cudaStream_t)stream;
    switch (s->features) {
    case 0: record_launch_lean(MODE, s->known, a); break;
    case FT_MID: record_launch_mid(MODE, s->known, a); break;
    case FT_MOSAICLEAN: record_launch_mosaic(MODE, s->known, a); break;
    case FT_SRCLEAN: record_launch_src(MODE, s->known, a); break;
    case FT_MESHLEAN: record_launch_mesh(MODE, s->known, a); break;
    default: record_launch_full(MODE, s->known, a); break;
    }
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

extern "C" int xrt_bits_to_ids(const uint32_t *bits_dev, uint64_t n_bits, uint64_t id_begin, uint64_t *ids_dev,
                               uint64_t capacity, uint64_t *count_dev, void *stream) {
    if (!bits_dev || (!ids_dev && capacity) || !count_dev) return fail(XRT_EINVAL, "null argument");
    cudaStream_t st = (cudaStream_t)stream;
    if (n_bits == 0) {
        CU(cudaMemsetAsync(count_dev, 0, sizeof(uint64_t), st));
        return XRT_OK;
    }
    const uint64_t n_words = (n_bits + 31) / 32;
    const uint64_t n_blocks = (n_words + kSelWordsPerBlock - 1) / kSelWordsPerBlock;
    if (n_blocks > 0x7fffffffull) return fail(XRT_EINVAL, "bitmap too large");
    void *scratch = nullptr;
    CU(cudaMallocAsync(&scratch, n_blocks * (sizeof(uint32_t) + sizeof(unsigned long long)) + 16, st));
    unsigned long long *offsets = (unsigned long long *)scratch;
    uint32_t *sums = (uint32_t *)(offsets + n_blocks);
    k_bits_count<<<(unsigned)n_blocks, kSelBlock, 0, st>>>(bits_dev, n_words, sums);
    k_bits_scan<<<1, 1024, 0, st>>>(sums, (uint32_t)n_blocks, offsets, (unsigned long long *)count_dev);
    k_bits_emit<<<(unsigned)n_blocks, kSelBlock, 0, st>>>(bits_dev, n_words, id_begin, offsets, ids_dev, capacity);
    CU(cudaGetLastError());
    CU(cudaFreeAsync(scratch, st));
    return XRT_OK;
}

extern "C" int xrt_lost_select(uint64_t seed, uint64_t stream_id, const uint64_t *ids_dev, uint64_t n, uint64_t m,
                               uint64_t *out_dev, uint64_t *count_dev, void *stream) {
    if (!count_dev || (n && (!ids_dev || !out_dev))) return fail(XRT_EINVAL, "null argument");
    cudaStream_t st = (cudaStream_t)stream;
    if (n == 0 || m == 0) {
        CU(cudaMemsetAsync(count_dev, 0, sizeof(uint64_t), st));
        return XRT_OK;
    }
    void *keys = nullptr;
    CU(cudaMallocAsync(&keys, n * sizeof(uint64_t), st));
    PhiloxKeys pk;
    philox_round_keys(seed, stream_id, pk);
    const uint64_t want = (n + 255) / 256;
    k_lost_keys<<<(unsigned)(want < 4096 ? want : 4096), 256, 0, st>>>(pk, stream_id, ids_dev, n, (uint64_t *)keys);
    k_select_smallest<<<1, 1024, 0, st>>>(ids_dev, (const uint64_t *)keys, n, m, out_dev, (unsigned long long *)count_dev);
    CU(cudaGetLastError());
    CU(cudaFreeAsync(keys, st));
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
