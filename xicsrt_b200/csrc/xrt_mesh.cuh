// xrt_mesh.cuh -- triangle-mesh optics (xicsrt/optics/_ShapeMesh.py).
//
// The reference intersects a mesh in up to four steps, each restated here per ray:
//   1. Moeller-Trumbore against every face of the mesh -- of the coarse mesh when
//      mesh_refine is set (:289-348).  No test on t; a later face overwrites an earlier hit.
//   2. (refine) nearest fine vertex of the coarse hit point (cKDTree.query, :464-475) and the
//      <= 8 faces around it.
//   3. (refine) ray/plane point for each candidate face, inside test by area sum with a 1e-10
//      tolerance, t >= 0, first passing candidate wins (:350-426).
//   4. (interpolate) z and the normal from Clough-Tocher interpolators over the xy Delaunay
//      triangulation (:172-196); otherwise the flat normal of the face that was hit (:428-432).
//
// Tables come from xicsrt_b200/mesh.py.  All lanes of a warp walk the face list in the same
// order, so the face operands are warp-uniform loads (one L1 transaction per load).
#pragma once
#include "../../include/xrt.h"
#include "xrt_math.cuh"

namespace xrt {

__device__ __forceinline__ V3 ld3(const double *p) { return v3(__ldg(p), __ldg(p + 1), __ldg(p + 2)); }

// step 1.  geom: [n_faces][9] = p0, edge1, edge2 (global memory, or the block's shared-memory
// copy when STAGED).  Returns the index of the last face hit or -1.
//
// The reference divides first (f = 1/a, u = f (s.h), v = f (D.q)) and then rejects u < 0, u > 1,
// v < 0, u + v > 1.  Almost every face is rejected by u, so the sign of u (= sign(a) sign(s.h),
// exact) and |s.h| > |a| are tested before the division is paid; the two forms can differ only
// for a hit within rounding of a triangle edge.
//
// Two passes per group of 32 faces: the cheap rejection marks the few faces a lane can hit in a
// bit mask, then each lane completes only its own candidates (in ascending face order, so the
// last face still wins).  A warp therefore pays the full test a couple of times per ray instead
// of once per face, although every face is hit by some lane of the warp.
template <bool STAGED>
__device__ __forceinline__ void mesh_face_operands(const double *__restrict__ g, V3 &p0, V3 &e1, V3 &e2) {
    if constexpr (STAGED) { p0 = v3(g); e1 = v3(g + 3); e2 = v3(g + 6); }
    else { p0 = ld3(g); e1 = ld3(g + 3); e2 = ld3(g + 6); }
}

template <bool STAGED>
__device__ __forceinline__ int mesh_all_faces(const double *__restrict__ geom, int n_faces, V3 o, V3 d, V3 &X) {
    const double eps = 1e-15;
    int hit = -1;
    for (int base = 0; base < n_faces; base += 32) {
        const int cnt = min(32, n_faces - base);
        unsigned cand = 0u;
        for (int j = 0; j < cnt; ++j) {
            V3 p0, e1, e2;
            mesh_face_operands<STAGED>(geom + 9 * (base + j), p0, e1, e2);
            const V3 h = cross(d, e2);
            const double a = dot(e1, h);
            const double sh = dot(o - p0, h);
            const bool degenerate = (a > -eps && a < eps);
            const bool u_neg = ((sh < 0.0) != (a < 0.0)) && sh != 0.0;       // u < 0
            const bool u_big = fabs(sh) > fabs(a);                           // u > 1
            if (!(degenerate || u_neg || u_big)) cand |= 1u << j;
        }
        while (cand) {
            const int j = __ffs(cand) - 1;
            cand &= cand - 1u;
            V3 p0, e1, e2;
            mesh_face_operands<STAGED>(geom + 9 * (base + j), p0, e1, e2);
            const V3 h = cross(d, e2);
            const double inv = 1.0 / dot(e1, h);
            const V3 s = o - p0;
            const double u = inv * dot(s, h);
            if (u < 0.0 || u > 1.0) continue;
            const V3 q = cross(s, e1);
            const double v = inv * dot(d, q);
            if (v < 0.0 || u + v > 1.0) continue;
            const double t = inv * dot(e2, q);
            hit = base + j;
            X = v3(o.x + t * d.x, o.y + t * d.y, o.z + t * d.z);
        }
    }
    return hit;
}

// step 1 for rays that all start at one point o (a point source in front of the first optic): with s = o - p0 fixed,
// the three triple products of Moeller-Trumbore are dot products of the direction with per-face constants,
//     a = e1.(d x e2) = d.N,  N = e2 x e1;    s.(d x e2) = d.A,  A = e2 x s;    d.(s x e1) = d.Q,  Q = s x e1,
// nine multiply-adds per face instead of two cross products.  They are not the reference's operation order, so they
// only PRE-SELECT (1e-9 relative slack on every edge, degenerate faces kept): the faces that survive are tested with
// the reference's own arithmetic from the global operands, in ascending order, and the last hit wins as before.
// naq: [n_faces][10] = N, A, Q, slack in shared memory (built by the block from face_geom and o).
constexpr int kPointRec = 10;
__device__ __forceinline__ void mesh_point_constants(const double *__restrict__ g, V3 o, double *__restrict__ out) {
    const V3 p0 = ld3(g), e1 = ld3(g + 3), e2 = ld3(g + 6);
    const V3 s = o - p0;
    const V3 N = cross(e2, e1), A = cross(e2, s), Q = cross(s, e1);
    out[0] = N.x; out[1] = N.y; out[2] = N.z;
    out[3] = A.x; out[4] = A.y; out[5] = A.z;
    out[6] = Q.x; out[7] = Q.y; out[8] = Q.z;
    // absolute slack of the pre-selection: 1e-12 of the operand scale (rounding is 1e-16 of it), for grazing rays
    // whose |a| is itself of the order of the rounding
    out[9] = 1e-12 * (sqrt(dot(N, N)) + sqrt(dot(A, A)) + sqrt(dot(Q, Q)));
}

__device__ __forceinline__ int mesh_all_faces_point(const double *__restrict__ naq, const double *__restrict__ geom, int n_faces,
                                                    V3 o, V3 d, V3 &X) {
    const double eps = 1e-15, tol = 1e-9;
    int hit = -1;
    for (int base = 0; base < n_faces; base += 32) {
        const int cnt = min(32, n_faces - base);
        unsigned cand = 0u;
        for (int j = 0; j < cnt; ++j) {
            const double *c = naq + kPointRec * (base + j);
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
        while (cand) {
            const int j = __ffs(cand) - 1;
            cand &= cand - 1u;
            V3 p0, e1, e2;
            mesh_face_operands<false>(geom + 9 * (base + j), p0, e1, e2);
            const V3 h = cross(d, e2);
            const double a = dot(e1, h);
            if (a > -eps && a < eps) continue;
            const double inv = 1.0 / a;
            const V3 s = o - p0;
            const double u = inv * dot(s, h);
            if (u < 0.0 || u > 1.0) continue;
            const V3 q = cross(s, e1);
            const double v = inv * dot(d, q);
            if (v < 0.0 || u + v > 1.0) continue;
            const double t = inv * dot(e2, q);
            hit = base + j;
            X = v3(o.x + t * d.x, o.y + t * d.y, o.z + t * d.z);
        }
    }
    return hit;
}

// step 1 through the face grid (un-refined meshes with many faces, XRT_F_MESH_LOSSLESS): the reference's loop tests
// every face and the last one hit wins.  A hit point lies on its face, i.e. inside the mesh's z range and inside the xy
// bounding box of that face, so only the faces registered in the cells that the ray's xy track crosses while it is
// inside [z_min, z_max] can be hit; they are tested with the same arithmetic and the highest face index among the hits
// wins -- the result of the full loop.  Rays nearly parallel to the xy plane, whose track covers more than 16 cells, take
// the full loop.
__device__ __forceinline__ bool mesh_test_face_mt(const double *__restrict__ g, V3 o, V3 d, V3 &X) {
    const double eps = 1e-15;
    V3 p0, e1, e2;
    mesh_face_operands<false>(g, p0, e1, e2);
    const V3 h = cross(d, e2);
    const double a = dot(e1, h);
    if (a > -eps && a < eps) return false;
    const double inv = 1.0 / a;
    const V3 s = o - p0;
    const double u = inv * dot(s, h);
    if (u < 0.0 || u > 1.0) return false;
    const V3 q = cross(s, e1);
    const double v = inv * dot(d, q);
    if (v < 0.0 || u + v > 1.0) return false;
    const double t = inv * dot(e2, q);
    X = v3(o.x + t * d.x, o.y + t * d.y, o.z + t * d.z);
    return true;
}

__device__ __forceinline__ int mesh_grid_faces(const XrtMesh &m, V3 o, V3 d, V3 &X) {
    // parameter range of the ray inside the z slab
    double t0 = (m.fgrid_z_min - o.z) / d.z, t1 = (m.fgrid_z_max - o.z) / d.z;
    bool wide = !(d.z != 0.0) || !(t0 == t0) || !(t1 == t1) || isinf(t0) || isinf(t1);
    int cx0 = 0, cx1 = 0, cy0 = 0, cy1 = 0;
    if (!wide) {
        const double xa = fma(t0, d.x, o.x), xb = fma(t1, d.x, o.x), ya = fma(t0, d.y, o.y), yb = fma(t1, d.y, o.y);
        const double dx = 1.0 / m.fgrid_inv_dx, dy = 1.0 / m.fgrid_inv_dy;
        // one cell of slack on each side: rounding of the slab end points, faces registered with a padded box
        const double fx0 = floor((fmin(xa, xb) - m.fgrid_x0) * m.fgrid_inv_dx - 1e-6), fx1 = floor((fmax(xa, xb) - m.fgrid_x0) * m.fgrid_inv_dx + 1e-6);
        const double fy0 = floor((fmin(ya, yb) - m.fgrid_y0) * m.fgrid_inv_dy - 1e-6), fy1 = floor((fmax(ya, yb) - m.fgrid_y0) * m.fgrid_inv_dy + 1e-6);
        (void)dx; (void)dy;
        if (fx1 < 0.0 || fy1 < 0.0 || fx0 >= (double)m.fgrid_nx || fy0 >= (double)m.fgrid_ny) return -1;   // track misses the footprint
        cx0 = (int)fmax(fx0, 0.0); cx1 = (int)fmin(fx1, (double)(m.fgrid_nx - 1));
        cy0 = (int)fmax(fy0, 0.0); cy1 = (int)fmin(fy1, (double)(m.fgrid_ny - 1));
        wide = (cx1 - cx0 + 1) * (cy1 - cy0 + 1) > 16;
    }
    int hit = -1;
    if (wide) {
        for (int f = 0; f < m.n_faces; ++f) {
            V3 P;
            if (mesh_test_face_mt(m.face_geom + 9 * (size_t)f, o, d, P)) { hit = f; X = P; }
        }
        return hit;
    }
    for (int cy = cy0; cy <= cy1; ++cy) {
        for (int cx = cx0; cx <= cx1; ++cx) {
            const int c = cy * m.fgrid_nx + cx;
            const int b = __ldg(m.fgrid_start + c), e = __ldg(m.fgrid_start + c + 1);
            for (int k = b; k < e; ++k) {
                const int f = __ldg(m.fgrid_items + k);
                if (f <= hit) continue;          // a lower face cannot win; also skips faces met in an earlier cell
                V3 P;
                if (mesh_test_face_mt(m.face_geom + 9 * (size_t)f, o, d, P)) { hit = f; X = P; }
            }
        }
    }
    return hit;
}

// step 2: exact nearest vertex through the uniform xy grid -- rings of cells around the query
// cell until no unvisited cell can hold a closer vertex (the xy distance to the ring bounds
// the 3-D distance from below).
__device__ __forceinline__ int mesh_nearest_vertex(const XrtMesh &m, V3 q) {
    const int nx = m.grid_nx, ny = m.grid_ny;
    const double dx = 1.0 / m.grid_inv_dx, dy = 1.0 / m.grid_inv_dy;
    int cx = (int)floor((q.x - m.grid_x0) * m.grid_inv_dx);
    int cy = (int)floor((q.y - m.grid_y0) * m.grid_inv_dy);
    cx = min(max(cx, 0), nx - 1);
    cy = min(max(cy, 0), ny - 1);
    if (m.nb_start) {
        // One contiguous list holds the vertices of the 3 x 3 block of cells around (cx, cy), coordinates and index
        // inline: two dependent loads instead of a walk over nine cells.  The winner is exact when it is closer than
        // the rim of the block (sides on the edge of the grid have nothing beyond them); otherwise the ring search
        // below decides.
        const int c = cy * nx + cx;
        const int b = __ldg(m.nb_start + c), e = __ldg(m.nb_start + c + 1);
        int bi = -1;
        double bd2 = CUDART_INF;
        for (int k = b; k < e; ++k) {
            const double2 xy = __ldg((const double2 *)(m.nb_rec + 4 * (size_t)k));
            const double2 zi = __ldg((const double2 *)(m.nb_rec + 4 * (size_t)k + 2));
            const V3 p = v3(xy.x, xy.y, zi.x) - q;
            const double d2 = dot(p, p);
            if (d2 < bd2) { bd2 = d2; bi = (int)zi.y; }
        }
        const double inf = CUDART_INF;
        const double rx0 = (cx > 0) ? q.x - (m.grid_x0 + (cx - 1) * dx) : inf;
        const double rx1 = (cx < nx - 1) ? (m.grid_x0 + (cx + 2) * dx) - q.x : inf;
        const double ry0 = (cy > 0) ? q.y - (m.grid_y0 + (cy - 1) * dy) : inf;
        const double ry1 = (cy < ny - 1) ? (m.grid_y0 + (cy + 2) * dy) - q.y : inf;
        const double rim = fmax(fmin(fmin(rx0, rx1), fmin(ry0, ry1)), 0.0);
        if (bi >= 0 && bd2 < rim * rim) return bi;
    }
    int best = -1;
    double best_d2 = CUDART_INF;
    const int max_ring = max(nx, ny);
    for (int ring = 0; ring <= max_ring; ++ring) {
        if (ring > 0) {
            const double ox = fmin(q.x - (m.grid_x0 + (cx - ring + 1) * dx), (m.grid_x0 + (cx + ring) * dx) - q.x);
            const double oy = fmin(q.y - (m.grid_y0 + (cy - ring + 1) * dy), (m.grid_y0 + (cy + ring) * dy) - q.y);
            const double bound = fmax(fmin(ox, oy), 0.0);
            if (bound * bound > best_d2) break;
        }
        for (int yy = cy - ring; yy <= cy + ring; ++yy) {
            if (yy < 0 || yy >= ny) continue;
            const bool edge_row = (yy == cy - ring) || (yy == cy + ring);
            const int step = edge_row ? 1 : 2 * ring;       // interior rows: only the two end cells
            for (int xx = cx - ring; xx <= cx + ring; xx += (step > 0 ? step : 1)) {
                if (xx < 0 || xx >= nx) continue;
                const int c = yy * nx + xx;
                const int b = __ldg(m.vgrid_start + c), e = __ldg(m.vgrid_start + c + 1);
                for (int k = b; k < e; ++k) {       // coordinates stored cell-ordered: no index hop
                    const double2 xy = __ldg((const double2 *)(m.vgrid_xyz + 4 * (size_t)k));
                    const double z = __ldg(m.vgrid_xyz + 4 * (size_t)k + 2);
                    const V3 p = v3(xy.x, xy.y, z) - q;
                    const double d2 = dot(p, p);
                    if (d2 < best_d2) { best_d2 = d2; best = k; }
                }
            }
        }
    }
    return best >= 0 ? __ldg(m.vgrid_items + best) : -1;
}

// step 3: the faces around vertex `vert`.  face_geom gives p0 and the two edges, face_area the
// constant |(p0 - p1) x (p0 - p2)| of the reference's area-sum test.
struct FaceRec { double2 g0, g1, g2, g3, g4, g5, g6, g7; };

__device__ __forceinline__ FaceRec load_face_rec(const double *rec) {
    const double2 *g = (const double2 *)rec;
    FaceRec r;
    r.g0 = __ldg(g); r.g1 = __ldg(g + 1); r.g2 = __ldg(g + 2); r.g3 = __ldg(g + 3); r.g4 = __ldg(g + 4);
    r.g5 = __ldg(g + 5); r.g6 = __ldg(g + 6); r.g7 = __ldg(g + 7);
    return r;
}

// one candidate face (_ShapeMesh.py:350-426): ray / plane point, inside test by the area sum, distance >= 0.
// Record (xicsrt_b200/mesh.py:device_tables): p0, m1 = e1 x n, m2 = e2 x n, n, area, face index, A0 = (e1 x e2) . n.
__device__ __forceinline__ bool mesh_test_face(const FaceRec &r, V3 o, V3 d, V3 &X) {
    const V3 p0 = v3(r.g0.x, r.g0.y, r.g1.x), m1 = v3(r.g1.y, r.g2.x, r.g2.y), m2 = v3(r.g3.x, r.g3.y, r.g4.x);
    const V3 n = v3(r.g4.y, r.g5.x, r.g5.y);
    const double area = r.g6.x, A0 = r.g7.x;
    const double dist = dot(p0 - o, n) / dot(d, n);
    if (!(dist >= 0.0)) return false;
    const V3 P = v3(d.x * dist + o.x, d.y * dist + o.y, d.z * dist + o.z);
    const V3 a = P - p0;
    // The reference sums the areas |b x c|, |c x a|, |a x b| of the three sub-triangles (b = a - e1, c = a - e2).  P lies
    // in the face plane, so each cross product is parallel to the unit normal and its length is |(.) . n|; with
    // (a x e) . n = a . (e x n):  (a x b) . n = -a . m1,  (c x a) . n = a . m2,  (b x c) . n = A0 + a . m1 - a . m2.
    const double s1 = dot(a, m1), s2 = dot(a, m2);
    const double diff = fabs(A0 + s1 - s2) + fabs(s2) + fabs(s1) - area;
    if (diff < 1e-10) {
        X = P;
        return true;
    }
    return false;
}

__device__ __forceinline__ int mesh_candidate_faces(const XrtMesh &m, int vert, V3 o, V3 d, V3 &X) {
    if (m.vertex_face_rec) {
        // the vertex's <= 8 face records sit back to back (no hop through the face index); a rolled loop: the mesh
        // variants are bound by instruction fetch as much as by loads (ncu: no_instruction 4.4 stall cycles per issue
        // with this loop unrolled)
        const double *base = m.vertex_face_rec + 128 * (size_t)vert;
#pragma unroll 1
        for (int k = 0; k < 8; ++k) {
            const FaceRec cur = load_face_rec(base + 16 * k);
            if (cur.g6.x < 0.0) break;          // records are packed: the first empty slot ends the list
            if (mesh_test_face(cur, o, d, X)) return (int)cur.g6.y;
        }
        return -1;
    }
    // the <= 8 face ids of the vertex in two 16-byte loads, then one 128-byte record per face
    const int4 fa = __ldg((const int4 *)(m.vertex_faces + 8 * (size_t)vert));
    const int4 fb = __ldg((const int4 *)(m.vertex_faces + 8 * (size_t)vert + 4));
    const int ids[8] = {fa.x, fa.y, fa.z, fa.w, fb.x, fb.y, fb.z, fb.w};
#pragma unroll 1
    for (int k = 0; k < 8; ++k) {
        const int f = ids[k];
        if (f < 0) continue;
        const FaceRec r = load_face_rec(m.face_rec + 16 * (size_t)f);
        if (mesh_test_face(r, o, d, X)) return f;
    }
    return -1;
}

// step 4: triangle of the xy Delaunay triangulation that holds (x, y): scipy's containment
// rule (all barycentric coordinates within [-eps, 1 + eps], eps = 100 DBL_EPSILON) on the
// triangles registered in the point's grid cell.  -1 = outside the hull (scipy returns NaN).
__device__ __forceinline__ int mesh_find_triangle(const XrtMesh &m, double x, double y, double &b0, double &b1, double &b2) {
    if (!(x == x) || !(y == y)) return -1;
    int cx = (int)floor((x - m.grid_x0) * m.grid_inv_dx);
    int cy = (int)floor((y - m.grid_y0) * m.grid_inv_dy);
    cx = min(max(cx, 0), m.grid_nx - 1);
    cy = min(max(cy, 0), m.grid_ny - 1);
    const int c = cy * m.grid_nx + cx;
    const int b = __ldg(m.grid_start + c), e = __ldg(m.grid_start + c + 1);
    const double eps = 100.0 * 2.220446049250313e-16;
    if (m.tri_rec) {
        // transform and triangle index inline in the cell's list: no hop through the triangle id
        for (int k = b; k < e; ++k) {
            const double2 *T = (const double2 *)(m.tri_rec + 8 * (size_t)k);
            const double2 t01 = __ldg(T), t23 = __ldg(T + 1), t45 = __ldg(T + 2), ti = __ldg(T + 3);
            const double ddx = x - t45.x, ddy = y - t45.y;
            b0 = t01.x * ddx + t01.y * ddy;
            b1 = t23.x * ddx + t23.y * ddy;
            b2 = 1.0 - b0 - b1;
            if (b0 >= -eps && b0 <= 1.0 + eps && b1 >= -eps && b1 <= 1.0 + eps && b2 >= -eps && b2 <= 1.0 + eps) return (int)ti.x;
        }
        return -1;
    }
    for (int k = b; k < e; ++k) {
        const int t = __ldg(m.grid_items + k);
        const double *T = m.tri_transform + 6 * (size_t)t;
        const double ddx = x - __ldg(T + 4), ddy = y - __ldg(T + 5);
        b0 = __ldg(T + 0) * ddx + __ldg(T + 1) * ddy;
        b1 = __ldg(T + 2) * ddx + __ldg(T + 3) * ddy;
        b2 = 1.0 - b0 - b1;
        if (b0 >= -eps && b0 <= 1.0 + eps && b1 >= -eps && b1 <= 1.0 + eps && b2 >= -eps && b2 <= 1.0 + eps) return t;
    }
    return -1;
}

// Clough-Tocher cubic of the four fields (z, nx, ny, nz) at once.  Coefficient order of xicsrt_b200/mesh.py CT_NAMES:
//  0 c3000  1 c0300  2 c0030  3 c0003  4 c2100  5 c2010  6 c2001  7 c1200  8 c0210  9 c0201
// 10 c1020 11 c0120 12 c0021 13 c1002 14 c0102 15 c0012 16 c1101 17 c1011 18 c0111
// Table layout [n_tri][19][4]: the 19 monomials (with their factors 1, 3, 6) are formed once and each multiplies the four
// field coefficients that sit side by side (two 16-byte loads).
__device__ __forceinline__ void ct_cubic4(const double *__restrict__ c, double b1, double b2, double b3, double b4, double w[4]) {
    const double b11 = b1 * b1, b22 = b2 * b2, b33 = b3 * b3, b44 = b4 * b4;
    const double t1 = 3.0 * b11, t2 = 3.0 * b22, t3 = 3.0 * b33, t4 = 3.0 * b44;
    const double s14 = 6.0 * b1 * b4, s234 = 6.0 * b2 * b3 * b4;
    const double mono[19] = {b11 * b1, b22 * b2, b33 * b3, b44 * b4,
                             t1 * b2, t1 * b3, t1 * b4, t2 * b1, t2 * b3, t2 * b4,
                             t3 * b1, t3 * b2, t3 * b4, t4 * b1, t4 * b2, t4 * b3,
                             s14 * b2, s14 * b3, s234};
    const double2 *c2 = (const double2 *)c;
    double w0 = 0.0, w1 = 0.0, w2 = 0.0, w3 = 0.0;
#pragma unroll
    for (int i = 0; i < 19; ++i) {
        const double2 ca = __ldg(c2 + 2 * i), cb = __ldg(c2 + 2 * i + 1);
        w0 = fma(mono[i], ca.x, w0);
        w1 = fma(mono[i], ca.y, w1);
        w2 = fma(mono[i], cb.x, w2);
        w3 = fma(mono[i], cb.y, w3);
    }
    w[0] = w0; w[1] = w1; w[2] = w2; w[3] = w3;
}

// number of faces step 1 walks (the coarse mesh when refining) and their operands
__device__ __forceinline__ int mesh_stage1_faces(const XrtOpticDesc &op, const double *&geom) {
    const XrtMesh &m = *op.mesh;
    if (op.flags & XRT_F_MESH_REFINE) { geom = m.coarse_geom; return m.n_coarse_faces; }
    geom = m.face_geom;
    return m.n_faces;
}

// ShapeMesh.intersect (:135-170).  o, d in the optic's tracing frame.  `staged` (optional) is a
// shared-memory copy of the step-1 face operands.
// The coarse step alone (step 1 of a refining mesh): true = some coarse face is hit, Xc = the hit point.
// The fused kernel runs it for every ray, re-packs the ~half that hit and resumes mesh_intersect with Xc.
// Returns the coarse face that was hit (the last one in face order) or -1.
__device__ __forceinline__ int mesh_coarse_face(const XrtOpticDesc &op, V3 o, V3 d, V3 &Xc, const double *staged = nullptr,
                                                bool point_constants = false) {
    const double *geom;
    const int n1 = mesh_stage1_faces(op, geom);
    Xc = nan3();
    if (staged && point_constants) return mesh_all_faces_point(staged, geom, n1, o, d, Xc);
    return staged ? mesh_all_faces<true>(staged, n1, o, d, Xc) : mesh_all_faces<false>(geom, n1, o, d, Xc);
}

__device__ __forceinline__ bool mesh_coarse_hit(const XrtOpticDesc &op, V3 o, V3 d, V3 &Xc, const double *staged = nullptr,
                                                bool point_constants = false) {
    return mesh_coarse_face(op, o, d, Xc, staged, point_constants) >= 0;
}

// The point where the ray meets the plane of one step-1 face: the arithmetic of the hit branch of mesh_all_faces for
// a face that is known to be hit (sorted mesh path: k_mesh_coarse found the face, k_trace needs the point again).
__device__ __forceinline__ V3 mesh_face_point(const double *__restrict__ g, V3 o, V3 d) {
    V3 p0, e1, e2;
    mesh_face_operands<false>(g, p0, e1, e2);
    const V3 h = cross(d, e2);
    const double inv = 1.0 / dot(e1, h);
    const V3 s = o - p0;
    const V3 q = cross(s, e1);
    const double t = inv * dot(e2, q);
    return v3(o.x + t * d.x, o.y + t * d.y, o.z + t * d.z);
}

// Spatial bin of a coarse hit point: the cells of the vertex grid (the cell arithmetic of mesh_nearest_vertex) cut in
// sub x sub, grouped in tiles of tile x tile, row-major.  Rays of one bin read the same few kB of the refinement tables.
__device__ __forceinline__ int mesh_bin_of(const XrtMesh &m, V3 q, int sub, int tile, int tiles_x) {
    int cx = (int)floor((q.x - m.grid_x0) * m.grid_inv_dx * (double)sub);
    int cy = (int)floor((q.y - m.grid_y0) * m.grid_inv_dy * (double)sub);
    cx = min(max(cx, 0), m.grid_nx * sub - 1);
    cy = min(max(cy, 0), m.grid_ny * sub - 1);
    if (tile == 1) return cy * tiles_x + cx;         // the default: no integer divisions (40 instructions per warp pass)
    return (cy / tile) * tiles_x + cx / tile;
}

// mesh_intersect (below) is the out-of-line copy: the fused kernel reaches it from four places (optics before / at /
// after the split optic, stage A2) and four inlined copies made the mesh variants 250 kB of code -- instruction-cache
// stalls of 4.5 cycles per issue.  k_mesh_refine, which has one call site on its hot path, inlines it (no call, no
// stack traffic for X and n, table loads scheduled with the caller's arithmetic).
__device__ __forceinline__ bool mesh_intersect_inline(const XrtOpticDesc &op, V3 o, V3 d, V3 &X, V3 &n,
                                                      const double *staged = nullptr, const V3 *resume_Xc = nullptr) {
    const XrtMesh &m = *op.mesh;
    X = nan3();
    n = nan3();
    const double *geom;
    const int n1 = mesh_stage1_faces(op, geom);
    int face;
    if (!(op.flags & XRT_F_MESH_REFINE) || (op.flags & XRT_F_MESH_LOSSLESS)) {
        if (m.fgrid_start) face = mesh_grid_faces(m, o, d, X);
        else if (op.flags & XRT_F_MESH_REFINE) face = mesh_all_faces<false>(m.face_geom, m.n_faces, o, d, X);
        else face = staged ? mesh_all_faces<true>(staged, n1, o, d, X) : mesh_all_faces<false>(geom, n1, o, d, X);
    } else {
        V3 Xc = nan3();
        if (resume_Xc) {        // coarse step already done (mesh_coarse_hit): a coarse face was hit at *resume_Xc
            Xc = *resume_Xc;
            face = 0;
        } else {
            face = staged ? mesh_all_faces<true>(staged, n1, o, d, Xc) : mesh_all_faces<false>(geom, n1, o, d, Xc);
        }
        if (face >= 0) {
            const int vert = mesh_nearest_vertex(m, Xc);
            face = (vert >= 0) ? mesh_candidate_faces(m, vert, o, d, X) : -1;
        }
    }
    if (face < 0) return false;

    if (op.flags & XRT_F_MESH_INTERP) {
        // outside the triangulation's hull scipy returns NaN for z and the normal; the ray
        // stays "hit" and carries the NaNs on, exactly as in the reference
        double b0, b1, b2;
        const int t = mesh_find_triangle(m, X.x, X.y, b0, b1, b2);
        if (t < 0) {
            X.z = CUDART_NAN;
            n = nan3();
        } else {
            const double mn = fmin(b0, fmin(b1, b2));
            const double e1 = b0 - mn, e2 = b1 - mn, e3 = b2 - mn, e4 = 3.0 * mn;
            double w[4];
            ct_cubic4(m.ct_coef + (size_t)t * 76, e1, e2, e3, e4, w);
            X.z = w[0];
            V3 nn = v3(w[1], w[2], w[3]);
            n = nn * rsqrt(dot(nn, nn));
        }
    } else {
        n = ld3(m.face_normals + 3 * face);
    }
    return true;
}

static __device__ __noinline__ bool mesh_intersect(const XrtOpticDesc &op, V3 o, V3 d, V3 &X, V3 &n,
                                            const double *staged = nullptr, const V3 *resume_Xc = nullptr) {
    return mesh_intersect_inline(op, o, d, X, n, staged, resume_Xc);
}

}  // namespace xrt
