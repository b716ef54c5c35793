/*
 * xrt.h -- C ABI of the B200 photon-raytrace path (libxrt.so).
 *
 * The reference (XICSRT 0.8.13) is pure Python and has no FFI of its own, so
 * these entry points are cut at the one place its driver hands the rays to the
 * numerics: the body of `_raytrace_iter` (xicsrt/xicsrt_raytrace.py:178-226),
 * i.e. `sources.generate_rays()` (:192) followed by `optics.trace()` (:194),
 * plus the per-element bookkeeping `Dispatcher.trace` does around each optic
 * (xicsrt/objects/_Dispatcher.py:166-196: num_out, history copy, make_image).
 * INTEGRATION.md shows the ctypes binding a maintainer of the reference adds.
 *
 * Conventions
 *   - plain C structs, pointers and sizes; no C++ or torch types;
 *   - every function returns 0 on success, a negative XRT_E* code on failure;
 *     xrt_last_error() returns the message for the calling thread's last error;
 *   - "dev" pointers are CUDA device pointers owned by the caller (the Python
 *     side allocates them as torch tensors); the library never frees them;
 *   - "host" pointers inside the *Desc structs are only read during
 *     xrt_scene_create, which uploads what it needs;
 *   - all work is enqueued on the cudaStream_t passed as `stream` (void*, 0 =
 *     legacy default stream) and is asynchronous with respect to the host;
 *   - all ray quantities are IEEE binary64, as in the reference
 *     (xicsrt/objects/_RayArray.py:82-86).
 */
#ifndef XRT_H
#define XRT_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define XRT_VERSION 1
#define XRT_MAX_OPTICS 16
#define XRT_MAX_SIGHTLINES 4

/* error codes */
#define XRT_OK 0
#define XRT_EINVAL (-1)      /* bad argument / descriptor                      */
#define XRT_EUNSUPPORTED (-2)/* class / option the kernels do not implement    */
#define XRT_ECUDA (-3)       /* a CUDA runtime call failed (see last error)    */
#define XRT_ENOMEM (-4)

/* ---- enumerations ------------------------------------------------------ */

/* surface shapes: xicsrt/optics/_Shape{Plane,Sphere,Cylinder,Torus,Mesh}.py */
enum { XRT_SHAPE_PLANE = 0, XRT_SHAPE_SPHERE = 1, XRT_SHAPE_CYLINDER = 2,
       XRT_SHAPE_TORUS = 3, XRT_SHAPE_MESH = 4 };

/* interactions: xicsrt/optics/_Interact{None,Mirror,Crystal,MosaicCrystal}.py */
enum { XRT_INTERACT_NONE = 0, XRT_INTERACT_MIRROR = 1, XRT_INTERACT_CRYSTAL = 2,
       XRT_INTERACT_MOSAIC = 3 };

/* rocking curves: xicsrt/optics/_InteractCrystal.py:139-184 */
enum { XRT_ROCK_STEP = 0, XRT_ROCK_GAUSS = 1, XRT_ROCK_TABLE = 2 };

/* optic flags */
enum { XRT_F_TRACE_LOCAL = 1 << 0, XRT_F_CHECK_SIZE = 1 << 1,
       XRT_F_CHECK_APERTURE = 1 << 2, XRT_F_CHECK_BRAGG = 1 << 3,
       XRT_F_CONVEX = 1 << 4, XRT_F_HAS_XSIZE = 1 << 5, XRT_F_HAS_YSIZE = 1 << 6,
       XRT_F_HAS_ZSIZE = 1 << 7, XRT_F_IMAGE = 1 << 8,
       XRT_F_MOSAIC_CUTOFF = 1 << 9, XRT_F_MESH_REFINE = 1 << 10,
       XRT_F_MESH_INTERP = 1 << 11,
       XRT_F_MESH_LOSSLESS = 1 << 12 };  /* refining mesh: find the fine face by the full test (through the face grid)
                                            instead of the reference's lossy coarse -> nearest-vertex pre-selection */

/* aperture shapes and logic: xicsrt/tools/xicsrt_aperture.py:13-204 */
enum { XRT_AP_NONE = 0, XRT_AP_CIRCLE = 1, XRT_AP_SQUARE = 2, XRT_AP_RECTANGLE = 3,
       XRT_AP_ELLIPSE = 4, XRT_AP_TRIANGLE = 5 };
enum { XRT_LOGIC_AND = 0, XRT_LOGIC_NOT = 1, XRT_LOGIC_OR = 2, XRT_LOGIC_NAND = 3,
       XRT_LOGIC_NOR = 4, XRT_LOGIC_XOR = 5, XRT_LOGIC_XNOR = 6 };

/* sources: xicsrt/sources/_XicsrtSource{Generic,Directed,Focused}.py, _XicsrtPlasma*.py */
enum { XRT_SRC_FIXED_AXIS = 0,   /* Generic (zaxis) and Directed (direction)   */
       XRT_SRC_FOCUSED = 1,      /* cone axis = normalize(target - origin)      */
       XRT_SRC_BUNDLES = 2 };    /* plasma: table of focused voxel sources      */
enum { XRT_SPATIAL_UNIFORM = 0, XRT_SPATIAL_GAUSSIAN = 1 };
/* cone distributions: xicsrt/tools/xicsrt_spread.py */
enum { XRT_CONE_ISOTROPIC = 0, XRT_CONE_ISOTROPIC_XY = 1, XRT_CONE_FLAT = 2,
       XRT_CONE_FLAT_XY = 3 };
/* wavelength models: xicsrt/sources/_XicsrtSourceGeneric.py:295-367 */
enum { XRT_WAVE_CONST = 0, XRT_WAVE_UNIFORM = 1, XRT_WAVE_NORMAL = 2, XRT_WAVE_TABLE = 3 };

/* ---- descriptors ------------------------------------------------------- */

typedef struct XrtAperture {
    int32_t shape;          /* XRT_AP_*    */
    int32_t logic;          /* XRT_LOGIC_* */
    double origin[2];
    double size[2];         /* circle: [r, -]; square: [w, -]; rectangle/ellipse: [sx, sy] */
    double vert[6];         /* triangle: x0 y0 x1 y1 x2 y2 (origin already added)          */
} XrtAperture;

/* triangle mesh tables of one ShapeMesh optic (xicsrt/optics/_ShapeMesh.py:198-287).
   All tables are built on the host at setup time (xicsrt_b200/mesh.py); the Clough-Tocher
   interpolators of the reference (scipy objects) become plain coefficient tables. */
typedef struct XrtMesh {
    int32_t n_points, n_faces;
    const double *points;        /* [n_points][3]                                        */
    const int32_t *faces;        /* [n_faces][3]                                         */
    const double *face_normals;  /* [n_faces][3]                                         */
    const double *face_geom;     /* [n_faces][9]: p0, p1 - p0, p2 - p0 (Moeller-Trumbore) */
    const double *face_area;     /* [n_faces]: |(p0 - p1) x (p0 - p2)| (area-sum inside test) */
    const double *face_rec;      /* [n_faces][16]: p0, e1 x n, e2 x n, unit normal n, area, (index), (e1 x e2).n, pad --
                                    one 128-byte record per face for the candidate-face test  */
    int32_t n_coarse_points, n_coarse_faces;   /* 0 when there is no coarse mesh         */
    const double *coarse_points;
    const int32_t *coarse_faces;
    const double *coarse_geom;   /* [n_coarse_faces][9]                                  */
    const int32_t *point_faces;  /* [8][n_points] faces around each fine point           */
    const uint8_t *point_faces_mask; /* [8][n_points]                                    */
    const int32_t *vertex_faces; /* [n_points][8] the same, vertex-major, -1 = no face   */
    /* Clough-Tocher interpolation of z and the normal over the xy Delaunay triangulation */
    int32_t n_tri;               /* 0 when mesh_interpolate is off                       */
    int32_t pad0;
    const double *ct_coef;       /* [n_tri][19][4] Bezier control coefficients of the cubic
                                    macro element, the fields z, nx, ny, nz side by side */
    const double *tri_transform; /* [n_tri][3][2] barycentric transform (scipy layout:
                                    rows 0,1 = inverse edge matrix, row 2 = third vertex) */
    /* uniform xy grid over the mesh footprint                                           */
    int32_t grid_nx, grid_ny;
    double grid_x0, grid_y0, grid_inv_dx, grid_inv_dy;
    const int32_t *grid_start;   /* [grid_nx*grid_ny + 1] triangles touching each cell   */
    const int32_t *grid_items;   /* triangle ids                                         */
    const int32_t *vgrid_start;  /* [grid_nx*grid_ny + 1] vertices in each cell (nearest-
                                    vertex query, replaces the reference's kd-tree)      */
    const int32_t *vgrid_items;
    const double *vgrid_xyz;     /* [items][4]: x, y, z of vgrid_items[k] (cell-ordered), pad */
    /* the same lookups with the data inline, cell- / vertex-ordered (fewer dependent loads per ray):       */
    const int32_t *nb_start;     /* [grid_nx*grid_ny + 1] vertices of the 3 x 3 block of cells around each cell */
    const double *nb_rec;        /* [items][4]: x, y, z, vertex index                                   */
    const double *tri_rec;       /* [grid_items][8]: barycentric transform (6), triangle index, pad      */
    const double *vertex_face_rec; /* [n_points][8][16]: face_rec of the faces around each vertex, [12] < 0 = no face,
                                      [13] = face index                                                  */
    /* uniform xy grid of the FINE faces (un-refined meshes with many faces, and XRT_F_MESH_LOSSLESS): a ray is tested
       against the faces of the cells its xy track crosses inside [z_min, z_max] only; NULL = test every face      */
    int32_t fgrid_nx, fgrid_ny;
    double fgrid_x0, fgrid_y0, fgrid_inv_dx, fgrid_inv_dy, fgrid_z_min, fgrid_z_max;
    const int32_t *fgrid_start;  /* [fgrid_nx*fgrid_ny + 1]                                               */
    const int32_t *fgrid_items;  /* face ids, ascending inside a cell                                     */
} XrtMesh;

typedef struct XrtOpticDesc {
    int32_t shape;           /* XRT_SHAPE_*    */
    int32_t interact;        /* XRT_INTERACT_* */
    int32_t rocking_type;    /* XRT_ROCK_*     */
    uint32_t flags;          /* XRT_F_*        */
    double origin[3];
    double orient[9];        /* rows: xaxis, yaxis = z cross x, zaxis (_GeometryObject.py:88-94) */
    double half_size[3];     /* xsize/2, ysize/2, zsize/2 (used when the HAS_* flag is set)      */
    double center[3];        /* sphere / cylinder / torus centre (global)                        */
    double radius;           /* sphere / cylinder                                                */
    double torus_major, torus_minor;   /* geometric radii (_ShapeTorus.py:70-87)                 */
    int32_t root_idx;        /* quartic solver slot (_ShapeTorus.py:72-85)                       */
    int32_t mosaic_depth;
    double two_d;            /* 2 * crystal_spacing                                              */
    double inv_two_d;        /* 1 / two_d (sin(theta_B) = wavelength * inv_two_d)                */
    double reflectivity;
    double rocking_fwhm;
    double rock_two_sigma2;  /* gaussian: 2 sigma^2 with sigma = fwhm / (2 sqrt(2 ln 2))         */
    double rock_inv_two_sigma2; /* 1 / rock_two_sigma2                                            */
    double rocking_mix;
    double mosaic_spread;    /* fwhm of crystallite normals [rad]                                */
    double mosaic_sin_sigma; /* sin(sigma) of the crystallite (x, y) offsets, sigma = hwhm/sqrt(2 ln 2) */
    double mosaic_angle_cut; /* angle cutoff derived from mosaic_cutoff (when XRT_F_MOSAIC_CUTOFF) */
    int32_t n_aperture;
    int32_t n_rock;
    const XrtAperture *apertures;     /* [n_aperture]                                            */
    const double *rock_dtheta;        /* [n_rock] radians, ascending                             */
    const double *rock_s;             /* [n_rock] sigma reflectivity                             */
    const double *rock_p;             /* [n_rock] pi reflectivity                                */
    const XrtMesh *mesh;              /* when shape == XRT_SHAPE_MESH                            */
    int32_t npix[2];         /* pixel_xsize, pixel_ysize (when XRT_F_IMAGE)                      */
    double pixel_size;
    uint64_t image_offset;   /* element offset of this optic's image in XrtOutputs.images        */
    /* Bragg pre-test of the fused kernel; filled in by xrt_scene_create (input values are ignored).
       cull_t2 > 0 enables it: a ray with (sin theta_B - sin theta_i)^2 > cull_t2 (1 - min(sin)^2)
       cannot pass the rocking curve (see bragg_cull in csrc/xrt_trace.cuh)                      */
    double cull_t2;          /* (1.05 angle beyond which the rocking curve is 0 or < 2^-57 + 2e-6)^2 */
    double cull_err;         /* bound on the error of the approximate sin theta_B                 */
    double cull_inv_r;       /* 1 / radius (sphere: n = (center - X) / radius)                    */
    /* mosaic crystal as split optic: per-layer FP32 pre-test of the crystallite loop (stage S of the fused kernel);
       filled in by xrt_scene_create: enable flag, first-level threshold (as cull_t2) and error margin on sin(theta). */
    int32_t mosaic_scan;
    int32_t pad2;
    double mosaic_t2;
    double mosaic_err;
} XrtOpticDesc;

typedef struct XrtSightline {    /* xicsrt/filters/_XicsrtBundleFilterSightline.py:31-56 */
    double origin[3];
    double axis[3];
    double radius;
} XrtSightline;

/* one plasma bundle = one focused voxel source (_XicsrtPlasmaGeneric.py:286-345) */
typedef struct XrtBundle {
    double origin[3];
    double cos_spread;       /* cone parameter of this bundle: cos(spread) for the isotropic cone,
                                tan(spread) for flat / flat_xy, sin(spread) for isotropic_xy
                                (a scalar spread s means [-s, s, -s, s], xicsrt_spread.py:352-366) */
    double wave_sigma;       /* Doppler sigma of this bundle [A] (XRT_WAVE_NORMAL)               */
    double velocity_c[3];    /* velocity / c                                                     */
} XrtBundle;

/* Per-iteration bundle table of a plasma source, built on the device: setup_bundles,
   bundle_filter, bundle_generate and the per-bundle ray counts of create_sources
   (xicsrt/sources/_XicsrtPlasmaGeneric.py:176-345; _XicsrtPlasmaCubic.py:23-35;
   _XicsrtPlasmaToroidal.py:34-78; _XicsrtPlasmaToroidalDatafile.py:30-45;
   xicsrt/filters/_XicsrtBundleFilterSightline.py:31-56). */
enum { XRT_PLASMA_GENERIC = 0, XRT_PLASMA_CUBIC = 1, XRT_PLASMA_TOROIDAL = 2, XRT_PLASMA_DATAFILE = 3 };

typedef struct XrtPlasmaDesc {
    int32_t kind;            /* XRT_PLASMA_*                                                     */
    int32_t use_poisson;     /* ray count = Poisson(intensity), else trunc(intensity)            */
    int32_t use_spread_radius; /* spread = atan(spread_radius / |origin - target|)               */
    int32_t n_sightlines;
    int32_t n_profile_t, n_profile_e;   /* DATAFILE: lengths of the rho -> value tables          */
    int32_t thermal_line;    /* 1: wave_sigma = sqrt(T) * sigma_factor (Doppler-broadened line);
                                2: the same with a natural linewidth: a bundle at T == 0 gets 1 eV
                                (_XicsrtSourceGeneric.py:333-339)                                   */
    int32_t cone;            /* XRT_CONE_* of the per-bundle sources (angular_dist)              */
    double origin[3];
    double orient[9];
    double size[3];          /* xsize, ysize, zsize of the plasma box                            */
    double target[3];
    double spread, spread_radius;
    double temperature, emissivity;     /* CUBIC / TOROIDAL constants                            */
    double velocity[3];                 /* TOROIDAL constant                                     */
    double temperature_scale, emissivity_scale, velocity_scale;
    double major_radius, minor_radius, torus_origin[3];
    double intensity_factor; /* time_resolution * bundle_volume / (4 pi) * volume / (bundle_count * bundle_volume) */
    double sigma_factor;     /* sqrt(1 / mass_number / amu / c^2 * eV) * wavelength              */
    double inv_c;            /* 1 / speed of light                                               */
    const double *profile_t_rho, *profile_t_val;   /* DEVICE pointers, caller-owned              */
    const double *profile_e_rho, *profile_e_val;
    const double *inject_u;  /* optional DEVICE [3][n] U[0,1) for the bundle centres (parity tests) */
    XrtSightline sightlines[XRT_MAX_SIGHTLINES];
} XrtPlasmaDesc;

typedef struct XrtSourceDesc {
    int32_t kind;            /* XRT_SRC_*     */
    int32_t spatial;         /* XRT_SPATIAL_* */
    int32_t cone;            /* XRT_CONE_*    */
    int32_t wave;            /* XRT_WAVE_*    */
    double origin[3];
    double orient[9];
    double extent[3];        /* uniform: full box sizes; gaussian: sigmas (fwhm / 2.3548)        */
    double axis_basis[9];    /* FIXED_AXIS: rows o_2, o_1, axis of the cone basis, precomputed
                                on the host exactly as _XicsrtSourceGeneric.py:282-288 does      */
    double target[3];        /* FOCUSED / BUNDLES                                                */
    double cone_par[4];      /* isotropic: [cos(spread)]; flat: [tan(spread)];
                                flat_xy: tan of [xmin,xmax,ymin,ymax];
                                isotropic_xy: sin of [xmin,xmax,ymin,ymax]                       */
    double cone_cos_max;     /* isotropic_xy: cos of the enclosing circular cone                 */
    double wave_par[4];      /* const: [lam]; uniform: [lo, hi]; normal: [lam0, sigma];
                                table: [lam0, cdf_min, cdf_max]                                  */
    double velocity_c[3];    /* velocity / c (all zero = no Doppler shift)                       */
    int32_t n_table;
    int32_t n_sightlines;
    const double *table_cdf; /* [n_table] ascending                                              */
    const double *table_x;   /* [n_table]                                                        */
    XrtSightline sightlines[XRT_MAX_SIGHTLINES];
    /* plasma */
    uint64_t n_bundles;
    const XrtBundle *bundles;        /* [n_bundles]                                              */
    const uint64_t *bundle_end;      /* [n_bundles] inclusive prefix sum of rays per bundle      */
    double voxel_size;
    /* plasma with a natural linewidth: one inverse-CDF table per bundle (XRT_WAVE_TABLE), rows of
       n_table entries; a row is valid where the bundle emits rays (xrt_bundle_voigt_tables)     */
    const double *bundle_x;          /* [n_bundles][n_table]                                     */
    const double *bundle_cdf;        /* [n_bundles][n_table]                                     */
    /* ray id -> bundle lookup hint, built by the library (input values are ignored): entry i is the
       bundle of ray id (i << bundle_hint_shift), so a ray's bundle lies in [hint[j], hint[j + 1]]
       with j = id >> shift and the binary search over bundle_end starts from that bracket         */
    const uint32_t *bundle_hint;
    int32_t bundle_hint_shift;
    int32_t pad1;
} XrtSourceDesc;

typedef struct XrtSceneDesc {
    int32_t version;         /* XRT_VERSION */
    int32_t n_optics;
    XrtSourceDesc source;
    XrtOpticDesc optics[XRT_MAX_OPTICS];
    /* single-precision constants of the spectrometer variant's FP32 broad phase (k_trace stage A32);
       filled in by xrt_scene_create, input values are ignored.  kn32[0] > 0 enables it. */
    float kn32[32];
} XrtSceneDesc;

typedef struct XrtScene XrtScene;   /* opaque; one per device context */

/* ---- per-call buffers (device pointers, caller-owned) ------------------ */

typedef struct XrtOutputs {
    uint64_t *counts;        /* [1 + n_optics] += rays alive after the source / each optic
                                (meta['num_out'], _Dispatcher.py:157-159,182-184)                */
    uint64_t *images;        /* concatenated per-optic pixel counts, [npix_x][npix_y] row-major,
                                += 1 per hit (_TraceObject.py:234-293); may be NULL              */
    uint64_t *found_ids;     /* optional: global ids of rays alive after the last optic          */
    uint64_t *found_count;   /* [1] number of found rays (may exceed found_capacity)             */
    uint64_t found_capacity;
    uint64_t *lost_ids;      /* optional: sample of lost rays (key < lost_threshold)             */
    uint64_t *lost_keys;     /* their 64-bit sampling keys (sort ascending, keep the first m)    */
    uint64_t *lost_count;    /* [1]                                                              */
    uint64_t lost_capacity;
    uint64_t lost_threshold; /* keep a lost ray when its key < threshold (2^64-1 = keep all)     */
    /* the same selections as bitmaps indexed by ray id - bits_begin (caller-zeroed, one bit per ray of the launch):
       no list capacity to guess, and xrt_bits_to_ids returns the ids in ascending (= the reference's ray) order  */
    uint32_t *found_bits;    /* optional: bit set for every ray alive after the last optic       */
    uint32_t *lost_bits;     /* optional: bit set for every lost ray whose key < lost_threshold  */
    uint64_t bits_begin;
} XrtOutputs;

/* per-element history, struct-of-arrays: 7 double planes + 1 byte plane per element
   (_Dispatcher.py:161-162,186-187 deepcopy of the ray dict after each element) */
enum { XRT_HIST_PLANES = 0,  /* struct-of-arrays planes: slot i of plane p at rays[(e * 7 + p) * capacity + i]         */
       XRT_HIST_ROWS = 1 };  /* the reference's row arrays inside the same 7 * capacity doubles per element: origin
                                rows [capacity][3], direction rows [capacity][3], wavelength [capacity]           */

typedef struct XrtHistory {
    double *rays;            /* [1 + n_optics][7][capacity]: ox oy oz dx dy dz wavelength (XRT_HIST_PLANES)  */
    uint8_t *mask;           /* [1 + n_optics][capacity]                                         */
    uint64_t capacity;
    int32_t layout;          /* XRT_HIST_*                                                       */
    int32_t pad0;
} XrtHistory;

typedef struct XrtRaysIn {   /* array-of-rows input, as the reference holds rays                */
    const double *origin;    /* [n][3] */
    const double *direction; /* [n][3] */
    const double *wavelength;/* [n]    */
    const uint8_t *mask;     /* [n]    */
} XrtRaysIn;

/* injected random draws, one entry per optic (NULL where the optic draws nothing).
   u[k]:  [depth_k][n]    U[0,1) for the rocking-curve test (_InteractCrystal.py:189)
   xy[k]: [depth_k][2][n] mosaic (x, y) offsets (xicsrt_spread.py:332)                           */
typedef struct XrtInject {
    const double *u[XRT_MAX_OPTICS];
    const double *xy[XRT_MAX_OPTICS];
} XrtInject;

/* injected raw draws for source generation (parity of _XicsrtSourceGeneric.py:198-393) */
typedef struct XrtSourceInject {
    const double *origin;    /* [3][n]: U[0,1) (uniform box) or final offsets (gaussian)        */
    const double *cone;      /* [2][n]: U[0,1)                                                   */
    const double *wave;      /* [n]: U[0,1) (uniform / table) or standard normal (normal)        */
} XrtSourceInject;

/* ---- entry points ------------------------------------------------------ */

int xrt_version(void);
const char *xrt_last_error(void);

/* Upload a scene (source + optic train + tables) to the current CUDA device. */
int xrt_scene_create(const XrtSceneDesc *desc, XrtScene **scene);
int xrt_scene_destroy(XrtScene *scene);

/* Fused generate -> trace -> bin for rays [ray_begin, ray_begin + ray_count) of random
   stream (seed, stream_id).  Replaces one `_raytrace_iter` call with history off.
   Random numbers are Philox4x32-10 keyed by (seed, stream_id) and counted by global ray
   id, so any partition of the id range over calls / GPUs gives identical results. */
int xrt_trace(XrtScene *scene, uint64_t seed, uint64_t stream_id, uint64_t ray_begin,
              uint64_t ray_count, const XrtOutputs *out, void *stream);

/* Re-trace the listed global ray ids (device array) and store every element's ray state
   into `hist` slot i for ids[i].  With xrt_trace's found/lost lists this yields the
   reference's found/lost histories (xicsrt_raytrace.py:229-278) without an N-sized copy.
   ids == NULL means ids[i] = ray_begin + i. */
int xrt_trace_history(XrtScene *scene, uint64_t seed, uint64_t stream_id,
                      const uint64_t *ids, uint64_t ray_begin, uint64_t n,
                      const XrtHistory *hist, void *stream);

/* Parity entry: trace caller-supplied rays with caller-supplied random draws through the
   optic train; fills counts / images (out, optional) and history (hist, optional).
   Element 0 of the history is the input ray set. */
int xrt_trace_injected(XrtScene *scene, const XrtRaysIn *rays, const XrtInject *draws,
                       uint64_t n, const XrtOutputs *out, const XrtHistory *hist, void *stream);

/* Parity entry for the source: generate n rays from caller-supplied raw draws into
   history element 0 of `hist` (box sources; fixed draw count distributions). */
int xrt_source_injected(XrtScene *scene, const XrtSourceInject *draws, uint64_t n,
                        const XrtHistory *hist, void *stream);

/* Generate rays [ray_begin, +n) of the Philox stream into history element 0 only. */
int xrt_source_generate(XrtScene *scene, uint64_t seed, uint64_t stream_id,
                        uint64_t ray_begin, uint64_t n, const XrtHistory *hist, void *stream);

/* Build the bundle table of one iteration on the device (all pointers are device pointers owned
   by the caller): table[i] for every bundle, intensity[i] = expected number of photons or -1 when
   the bundle is filtered out, counts[i] = rays the bundle emits.  Random numbers: Philox keyed by
   (seed, stream_id), counted by bundle index. */
int xrt_bundles_generate(const XrtPlasmaDesc *desc, uint64_t seed, uint64_t stream_id, uint64_t n_bundles,
                         XrtBundle *table_dev, double *intensity_dev, int64_t *counts_dev, void *stream);

/* Point a plasma scene (source.kind == XRT_SRC_BUNDLES) at a device-resident bundle table and its
   inclusive prefix sum of counts; bundles with a zero count are skipped by the lookup.  n_rays =
   end_dev[n_bundles - 1] as known to the host (0 = unknown: no lookup hint is built).  The hint
   table is built on the legacy default stream, which orders it after work on blocking streams. */
int xrt_scene_set_bundles(XrtScene *scene, const XrtBundle *table_dev, const uint64_t *end_dev, uint64_t n_bundles,
                          uint64_t n_rays);

/* Per-bundle Voigt inverse-CDF tables (xicsrt/tools/xicsrt_voigt.py:30-92, built by every
   per-bundle XicsrtSourceFocused in _XicsrtSourceGeneric.py:319-354): for each bundle b with
   counts_dev[b] > 0, sigma = table_dev[b].wave_sigma and the given Lorentzian gamma [A], writes
   the right bin edges x_dev[b][0..n_table) and the cumulative sums cdf_dev[b][0..n_table) of
   pdf * dx on the reference's stretched grid.  The Voigt profile is evaluated on the device
   (Faddeeva function by the exponentially convergent trapezoid rule with pole correction). */
int xrt_bundle_voigt_tables(const XrtBundle *table_dev, const int64_t *counts_dev, uint64_t n_bundles,
                            double gamma, int32_t n_table, double *x_dev, double *cdf_dev, void *stream);

/* Attach per-bundle wavelength tables to a plasma scene (after xrt_scene_set_bundles). */
int xrt_scene_set_bundle_tables(XrtScene *scene, const double *x_dev, const double *cdf_dev, int32_t n_table);

/* Ascending list of the ray ids whose bit is set in a bitmap written by xrt_trace (found_bits / lost_bits):
   ids_dev[k] = id_begin + index of the k-th set bit; *count_dev = number of set bits (may exceed capacity).
   Count / scan / emit kernels, no sort. */
int xrt_bits_to_ids(const uint32_t *bits_dev, uint64_t n_bits, uint64_t id_begin, uint64_t *ids_dev,
                    uint64_t capacity, uint64_t *count_dev, void *stream);

/* The lost sample of the reference (`_sort_raytrace`, xicsrt_raytrace.py:262-266: a shuffle, then the first
   max_lost): of the n candidate ids, the m with the smallest Philox sampling keys of (seed, stream_id) -- a
   uniform random subset -- written to out_dev in ascending id order; *count_dev = min(m, n). */
int xrt_lost_select(uint64_t seed, uint64_t stream_id, const uint64_t *ids_dev, uint64_t n, uint64_t m,
                    uint64_t *out_dev, uint64_t *count_dev, void *stream);

/* FP64 FMA-chain microbenchmark: runs `iters` dependent-chain DFMA steps per thread on a
   full grid and reports the number of FP64 flops issued; the caller times it with CUDA
   events to obtain the roofline denominator.  out_dev: [1] double (sink). */
int xrt_fp64_burn(uint64_t iters, double *out_dev, double *flops, void *stream);

/* Launch geometry used by xrt_trace (for reporting). */
int xrt_launch_info(XrtScene *scene, int32_t *grid, int32_t *block, int32_t *regs,
                    int32_t *blocks_per_sm);

/* The FP32 broad phase in front of xrt_trace (k_cull32): mode = -1 when it does not apply to the scene, else the
   source kind it was built for (0 point, 1 box, 2 focused box, 3 plasma bundles); launch geometry as above.
   mode = 4: the scene takes the sorted mesh path instead (k_mesh_coarse -> counting sort by hit location ->
   k_mesh_refine, csrc/xrt_meshsort.cuh): grid / low 16 bits of regs / low 8 bits of blocks_per_sm describe
   k_mesh_coarse, the high 16 bits of regs are the registers of k_mesh_refine and blocks_per_sm >> 8 is the
   number of spatial bins.  mode = 5: the broad phase of a mosaic crystal's crystallite scan (k_mosaic32). */
int xrt_launch_info_cull(XrtScene *scene, int32_t *mode, int32_t *grid, int32_t *regs, int32_t *blocks_per_sm);

#ifdef __cplusplus
}
#endif
#endif /* XRT_H */
