# -*- coding: utf-8 -*-
"""
Triangle-mesh optics: setup-time tables (host side).

Three layers, all setup-time (nothing here runs per ray):

1. mesh generators of the built-in classes -- the (x, y) grid of ``XicsrtOpticMeshSphericalCrystal``
   (``xicsrt/optics/_ShapeMeshSphere.py:75-98``) and the angle grids of the cylindrical and toroidal
   ones (``_ShapeMeshCylinder.py:94-192``, ``_ShapeMeshTorus.py:84-268``).  The per-point formulas
   are evaluated point by point in the reference's operation order: the grids are symmetric, so
   the Delaunay triangulation of their (x, y) projection has exactly co-circular cells whose
   diagonal is decided by the last bit of the coordinates.
2. ``_mesh_precalc`` (``_ShapeMesh.py:198-264``): xy-Delaunay, face normals, vertex -> faces table,
   nearest-vertex tree, scipy Clough-Tocher interpolators for z and the normal components.
3. device tables (``fill_mesh``): what the kernel needs instead of scipy objects -- per-triangle
   barycentric transforms and the 19 Bezier control coefficients of the Clough-Tocher macro
   element for each of the 4 interpolated fields (scipy's own vertex gradients, so the cubic
   is the one scipy evaluates), plus uniform xy grids for point-in-triangulation and
   nearest-vertex queries.
"""
import ctypes as C
import logging

import numpy as np

from . import _lib as L

log = logging.getLogger('xicsrt_b200')


# ---------------------------------------------------------------------------
# 1. generators

def _rotate(a, b, theta):
    """xicsrt/tools/xicsrt_math.py:72-99, 1-D branch."""
    a = np.asarray(a)
    b = np.asarray(b)
    b_hat = b / np.linalg.norm(b)
    u = b_hat * np.dot(a, b_hat)
    v = a - u
    w = np.cross(b_hat, v)
    return u + v * np.cos(theta) + w * np.sin(theta)


def _sphere_mesh(param, meshsize):
    from scipy.spatial import Delaunay
    xsize, ysize, radius = param['xsize'], param['ysize'], param['radius']
    x = np.linspace(-xsize / 2, xsize / 2, meshsize[0])
    y = np.linspace(-ysize / 2, ysize / 2, meshsize[1])
    center = np.array([0.0, 0.0, radius])
    xx, yy = np.meshgrid(x, y)
    zz = radius - np.sqrt(radius**2 - xx**2 - yy**2)
    points = np.stack((xx.flatten(), yy.flatten(), zz.flatten())).T
    norm = center - points
    norm = norm / np.expand_dims(np.linalg.norm(norm, axis=1), 1)
    return points, norm, Delaunay(points[:, 0:2]).simplices


def _cylinder_point(param, x, angle):
    z0 = np.array([0.0, 0.0, 1.0])
    x0 = np.array([1.0, 0.0, 0.0])
    radius = param['radius']
    o = np.array([0.0, 0.0, 0.0]) + radius * z0 + np.array([x, 0.0, 0.0])
    n = _rotate(z0, x0, angle)
    return o - radius * n, n


def _torus_point(param, a, b):
    z0 = np.asarray([0.0, 0.0, 1.0])
    x0 = np.asarray([1.0, 0.0, 0.0])
    s_maj, s_min = param['torus_sign_major'], param['torus_sign_minor']
    r_maj, r_min = param['radius_major'], param['radius_minor']
    y0 = np.cross(z0, x0)
    center = r_maj * z0 * s_maj
    c_norm = _rotate(z0, y0, a)
    c = center - r_maj * c_norm * s_maj
    q = c + r_min * c_norm * s_min
    axis = np.cross(c_norm * s_min, y0)
    x_norm = _rotate(c_norm * s_min, axis, b)
    return q - x_norm * r_min, x_norm


def _torus_point_fd(param, a, b, delta=1e-8):
    xyz, _ = _torus_point(param, a, b)
    xyz1, _ = _torus_point(param, a + delta, b)
    xyz2, _ = _torus_point(param, a, b + delta)
    n = np.cross(xyz1 - xyz, xyz2 - xyz)
    return xyz, n / np.linalg.norm(n)


def _angle_mesh(point_fn, a_range, b_range, mesh_size):
    from scipy.spatial import Delaunay
    a = np.linspace(a_range[0], a_range[1], mesh_size[0])
    b = np.linspace(b_range[0], b_range[1], mesh_size[1])
    aa, bb = np.meshgrid(a, b, indexing='ij')
    pts = np.empty((len(a), len(b), 3))
    nrm = np.empty((len(a), len(b), 3))
    for i in range(len(a)):
        for j in range(len(b)):
            pts[i, j], nrm[i, j] = point_fn(aa[i, j], bb[i, j])
    angles_2d = np.stack((aa.flatten(), bb.flatten()), axis=0).T
    return pts.reshape(-1, 3), nrm.reshape(-1, 3), Delaunay(angles_2d).simplices


def setup_mesh(param):
    """The ``setup()`` of the built-in mesh classes: fill mesh_* / mesh_coarse_* from the shape parameters."""
    shape = param['_shape']
    if shape == 'mesh':
        for key in ('mesh_points', 'mesh_normals', 'mesh_faces', 'mesh_coarse_points', 'mesh_coarse_normals',
                    'mesh_coarse_faces'):
            if param[key] is not None:
                param[key] = np.asarray(param[key])
        return
    if shape == 'mesh_sphere':
        gen = lambda size: _sphere_mesh(param, size)
    elif shape == 'mesh_cylinder':
        xsize = param['xsize'] if param['mesh_xsize'] is None else param['mesh_xsize']
        ysize = param['ysize'] if param['mesh_ysize'] is None else param['mesh_ysize']
        param['x_range'] = [-1 * xsize / 2, xsize / 2]
        half = np.arcsin(ysize / 2 / param['radius'])
        param['angle_range'] = [-1 * half, half]
        gen = lambda size: _angle_mesh(lambda a, b: _cylinder_point(param, a, b), param['x_range'],
                                       param['angle_range'], size)
    elif shape == 'mesh_torus':
        cvx = tuple(bool(v) for v in np.asarray(param['convex']).ravel())
        if len(cvx) != 2:
            raise Exception(f"Cannot be parse convex config option: {param['convex']}")
        param['torus_sign_major'] = -1 if cvx[0] else 1
        param['torus_sign_minor'] = -1 if cvx[1] else 1
        xsize = param['xsize'] if param['mesh_xsize'] is None else param['mesh_xsize']
        ysize = param['ysize'] if param['mesh_ysize'] is None else param['mesh_ysize']
        half_major = np.arcsin(xsize / 2 / param['radius_major'])
        half_minor = np.arcsin(ysize / 2 / param['radius_minor'])
        param['angle_major'] = [-1 * half_major, half_major]
        param['angle_minor'] = [-1 * half_minor, half_minor]
        method = param['normal_method']
        if method == 'analytic':
            fn = lambda a, b: _torus_point(param, a, b)
        elif method == 'fd':
            fn = lambda a, b: _torus_point_fd(param, a, b)
        elif method == 'jax':
            raise NotImplementedError()
        else:
            raise Exception(f"normal_method {method} unknown.")
        gen = lambda size: _angle_mesh(fn, param['angle_major'], param['angle_minor'], size)
    else:
        raise KeyError(shape)
    param['mesh_points'], param['mesh_normals'], param['mesh_faces'] = gen(param['mesh_size'])
    param['mesh_coarse_points'], param['mesh_coarse_normals'], param['mesh_coarse_faces'] = gen(param['mesh_coarse_size'])


# ---------------------------------------------------------------------------
# 2. _mesh_precalc

def point_faces_table(n_points, faces):
    """
    Faces around each point, in ascending face order: (8, P) indices + mask
    (_ShapeMesh.py:446-462; more than 8 faces at a vertex is an error there too).
    """
    f_idx = np.repeat(np.arange(len(faces)), 3)
    p_idx = np.asarray(faces).ravel()
    order = np.argsort(p_idx, kind='stable')
    p_sorted, f_sorted = p_idx[order], f_idx[order]
    start = np.searchsorted(p_sorted, np.arange(n_points))
    count = np.searchsorted(p_sorted, np.arange(n_points), side='right') - start
    if count.max(initial=0) > 8:
        raise ValueError('could not broadcast: a mesh vertex belongs to more than 8 faces')
    idx = np.zeros((8, n_points), dtype=np.int32)
    mask = np.zeros((8, n_points), dtype=np.bool_)
    for k in range(8):
        sel = count > k
        idx[k, sel] = f_sorted[start[sel] + k]
        mask[k, sel] = True
    return idx, mask


def precalc(param, points, normals, faces):
    from scipy.interpolate import CloughTocher2DInterpolator
    from scipy.spatial import Delaunay, cKDTree
    out = {'points': points, 'normals': normals, 'faces': faces}
    delaunay = Delaunay(points[:, 0:2])
    if faces is None:
        faces = delaunay.simplices
        out['faces'] = faces
    if param['mesh_interpolate']:
        out['interp'] = {
            'z': CloughTocher2DInterpolator(delaunay, points[:, 2].flatten()),
            'normal_x': CloughTocher2DInterpolator(delaunay, normals[:, 0].flatten()),
            'normal_y': CloughTocher2DInterpolator(delaunay, normals[:, 1].flatten()),
            'normal_z': CloughTocher2DInterpolator(delaunay, normals[:, 2].flatten()),
        }
    p0, p1, p2 = points[faces[..., 0], :], points[faces[..., 1], :], points[faces[..., 2], :]
    out['faces_center'] = np.mean(np.array([p0, p1, p2]), 0)
    fn = np.cross((p0 - p1), (p2 - p1))
    fn /= np.linalg.norm(fn, axis=1)[:, None]
    out['faces_normal'] = fn
    out['points_tree'] = cKDTree(points)
    out['p_faces_idx'], out['p_faces_mask'] = point_faces_table(len(points), faces)
    return out


def initialize_mesh(param):
    """check_param + mesh_initialize of ShapeMesh (_ShapeMesh.py:111-133, 266-287)."""
    if param['mesh_points'] is None:
        raise Exception('A mesh optic needs mesh_points.')
    if param['mesh_interpolate'] is None:
        param['mesh_interpolate'] = (param['mesh_normals'] is not None)
    elif param['mesh_interpolate']:
        if param['mesh_normals'] is None:
            raise Exception('Surface normal vectors must be defined in order to use mesh interpolation.')
    if param['mesh_refine'] is None:
        if param['mesh_coarse_points'] is not None:
            param['mesh_refine'] = True
    pts = param['mesh_points']
    spread = [np.max(pts[:, i]) - np.min(pts[:, i]) for i in range(3)]
    if spread[2] > spread[0] or spread[2] > spread[1]:
        log.warning('Mesh is not oriented with the surface normals near the local z direction.\n'
                    'This may lead to unexpected and incorrect results.')
    param['mesh'] = precalc(param, param['mesh_points'], param['mesh_normals'], param['mesh_faces'])
    if param['mesh_coarse_points'] is not None:
        param['mesh_coarse'] = precalc(param, param['mesh_coarse_points'], param['mesh_coarse_normals'],
                                       param['mesh_coarse_faces'])


# ---------------------------------------------------------------------------
# 3. Clough-Tocher macro element as plain tables

# order of the 19 Bezier control coefficients in the device table
CT_NAMES = ('c3000', 'c0300', 'c0030', 'c0003', 'c2100', 'c2010', 'c2001', 'c1200', 'c0210', 'c0201',
            'c1020', 'c0120', 'c0021', 'c1002', 'c0102', 'c0012', 'c1101', 'c1011', 'c0111')


def barycentric_transforms(points, simplices):
    """
    (n_tri, 3, 2) in scipy's Delaunay.transform layout: rows 0, 1 = inverse of the matrix whose
    columns are p0 - p2 and p1 - p2, row 2 = p2.  Closed-form 2x2 inverse (scipy's lazily computed
    property costs ~1 s for a few thousand triangles; results agree to rounding, checked in
    tests/test_mesh_tables.py).
    """
    p0, p1, p2 = points[simplices[:, 0]], points[simplices[:, 1]], points[simplices[:, 2]]
    a, b = p0[:, 0] - p2[:, 0], p1[:, 0] - p2[:, 0]
    c, d = p0[:, 1] - p2[:, 1], p1[:, 1] - p2[:, 1]
    det = a * d - b * c
    T = np.empty((len(simplices), 3, 2))
    T[:, 0, 0], T[:, 0, 1] = d / det, -b / det
    T[:, 1, 0], T[:, 1, 1] = -c / det, a / det
    T[:, 2, :] = p2
    return T


def ct_coefficients(tri, values, grad, transform=None):
    """
    Control coefficients of scipy's Clough-Tocher cubic for every triangle (vectorised restatement of
    scipy/interpolate/interpnd.pyx ``_clough_tocher_2d_single``): vertex values and scipy's estimated
    vertex gradients along the edges; the cross-edge derivative is made linear using the centroid of
    the neighbouring triangle (g = -1/2 on hull edges).

    tri: scipy Delaunay; values: (P,); grad: (P, 2).  Returns (n_tri, 19) in CT_NAMES order.
    """
    pts = tri.points
    simp = tri.simplices
    p0, p1, p2 = pts[simp[:, 0]], pts[simp[:, 1]], pts[simp[:, 2]]
    e12, e23, e31 = p1 - p0, p2 - p1, p0 - p2
    f1, f2, f3 = values[simp[:, 0]], values[simp[:, 1]], values[simp[:, 2]]
    g1, g2, g3 = grad[simp[:, 0]], grad[simp[:, 1]], grad[simp[:, 2]]
    dot = lambda g, e: g[:, 0] * e[:, 0] + g[:, 1] * e[:, 1]
    df12, df21 = +dot(g1, e12), -dot(g2, e12)
    df23, df32 = +dot(g2, e23), -dot(g3, e23)
    df31, df13 = +dot(g3, e31), -dot(g1, e31)

    c = {}
    c['c3000'] = f1
    c['c2100'] = (df12 + 3 * c['c3000']) / 3
    c['c2010'] = (df13 + 3 * c['c3000']) / 3
    c['c0300'] = f2
    c['c1200'] = (df21 + 3 * c['c0300']) / 3
    c['c0210'] = (df23 + 3 * c['c0300']) / 3
    c['c0030'] = f3
    c['c1020'] = (df31 + 3 * c['c0030']) / 3
    c['c0120'] = (df32 + 3 * c['c0030']) / 3
    c['c2001'] = (c['c2100'] + c['c2010'] + c['c3000']) / 3
    c['c0201'] = (c['c1200'] + c['c0300'] + c['c0210']) / 3
    c['c0021'] = (c['c1020'] + c['c0120'] + c['c0030']) / 3

    g = np.full((len(simp), 3), -0.5)
    T = barycentric_transforms(pts, simp) if transform is None else transform
    for k in range(3):
        nb = tri.neighbors[:, k]
        has = nb != -1
        cen = (pts[simp[nb[has], 0]] + pts[simp[nb[has], 1]] + pts[simp[nb[has], 2]]) / 3
        d = cen - T[has, 2, :]
        b0 = T[has, 0, 0] * d[:, 0] + T[has, 0, 1] * d[:, 1]
        b1 = T[has, 1, 0] * d[:, 0] + T[has, 1, 1] * d[:, 1]
        bc = np.stack([b0, b1, 1.0 - b0 - b1], axis=1)
        if k == 0:
            g[has, k] = (2 * bc[:, 2] + bc[:, 1] - 1) / (2 - 3 * bc[:, 2] - 3 * bc[:, 1])
        elif k == 1:
            g[has, k] = (2 * bc[:, 0] + bc[:, 2] - 1) / (2 - 3 * bc[:, 0] - 3 * bc[:, 2])
        else:
            g[has, k] = (2 * bc[:, 1] + bc[:, 0] - 1) / (2 - 3 * bc[:, 1] - 3 * bc[:, 0])

    c['c0111'] = (g[:, 0] * (-c['c0300'] + 3 * c['c0210'] - 3 * c['c0120'] + c['c0030'])
                  + (-c['c0300'] + 2 * c['c0210'] - c['c0120'] + c['c0021'] + c['c0201'])) / 2
    c['c1011'] = (g[:, 1] * (-c['c0030'] + 3 * c['c1020'] - 3 * c['c2010'] + c['c3000'])
                  + (-c['c0030'] + 2 * c['c1020'] - c['c2010'] + c['c2001'] + c['c0021'])) / 2
    c['c1101'] = (g[:, 2] * (-c['c3000'] + 3 * c['c2100'] - 3 * c['c1200'] + c['c0300'])
                  + (-c['c3000'] + 2 * c['c2100'] - c['c1200'] + c['c2001'] + c['c0201'])) / 2
    c['c1002'] = (c['c1101'] + c['c1011'] + c['c2001']) / 3
    c['c0102'] = (c['c1101'] + c['c0111'] + c['c0201']) / 3
    c['c0012'] = (c['c1011'] + c['c0111'] + c['c0021']) / 3
    c['c0003'] = (c['c1002'] + c['c0102'] + c['c0012']) / 3
    return np.stack([c[name] for name in CT_NAMES], axis=1)


def ct_evaluate(coef, b):
    """The cubic at barycentric coordinates b (n, 3) of triangles with coefficients coef (n, 19)."""
    c = {name: coef[:, i] for i, name in enumerate(CT_NAMES)}
    minval = np.min(b, axis=1)
    b1, b2, b3, b4 = b[:, 0] - minval, b[:, 1] - minval, b[:, 2] - minval, 3 * minval
    return (b1**3 * c['c3000'] + 3 * b1**2 * b2 * c['c2100'] + 3 * b1**2 * b3 * c['c2010'] + 3 * b1**2 * b4 * c['c2001']
            + 3 * b1 * b2**2 * c['c1200'] + 6 * b1 * b2 * b4 * c['c1101'] + 3 * b1 * b3**2 * c['c1020']
            + 6 * b1 * b3 * b4 * c['c1011'] + 3 * b1 * b4**2 * c['c1002'] + b2**3 * c['c0300']
            + 3 * b2**2 * b3 * c['c0210'] + 3 * b2**2 * b4 * c['c0201'] + 3 * b2 * b3**2 * c['c0120']
            + 6 * b2 * b3 * b4 * c['c0111'] + 3 * b2 * b4**2 * c['c0102'] + b3**3 * c['c0030']
            + 3 * b3**2 * b4 * c['c0021'] + 3 * b3 * b4**2 * c['c0012'] + b4**3 * c['c0003'])


def barycentric(transform, xy):
    """scipy's _barycentric_coordinates for matching rows of transform (n, 3, 2) and xy (n, 2)."""
    d = xy - transform[:, 2, :]
    b0 = transform[:, 0, 0] * d[:, 0] + transform[:, 0, 1] * d[:, 1]
    b1 = transform[:, 1, 0] * d[:, 0] + transform[:, 1, 1] * d[:, 1]
    return np.stack([b0, b1, 1.0 - b0 - b1], axis=1)


# containment tolerance of scipy's find_simplex (qhull.pyx: eps = 100 * DBL_EPSILON)
SIMPLEX_EPS = 100 * np.finfo(np.float64).eps


def _cell_lists(nx, ny, cell_of_item, n_items_hint=None):
    """CSR lists from (item, cell) pairs."""
    items, cells = cell_of_item
    order = np.argsort(cells, kind='stable')
    cells_sorted = cells[order]
    start = np.searchsorted(cells_sorted, np.arange(nx * ny + 1)).astype(np.int32)
    return start, items[order].astype(np.int32)


def lookup_grids(points_xy, tri_simplices, cells_per_axis=None):
    """
    Uniform xy grid over the mesh footprint with, per cell, the triangles whose bounding box
    touches the cell (point-in-triangulation) and the vertices inside it (nearest vertex).
    """
    lo = points_xy.min(axis=0)
    hi = points_xy.max(axis=0)
    n = len(points_xy)
    if cells_per_axis is None:
        cells_per_axis = int(np.clip(np.sqrt(n), 4, 1024))
    nx = ny = cells_per_axis
    span = np.maximum(hi - lo, 1e-300)
    inv = np.array([nx, ny]) / span

    def cell_xy(xy):
        c = np.floor((xy - lo) * inv).astype(np.int64)
        c[:, 0] = np.clip(c[:, 0], 0, nx - 1)
        c[:, 1] = np.clip(c[:, 1], 0, ny - 1)
        return c

    # vertices
    vc = cell_xy(points_xy)
    vstart, vitems = _cell_lists(nx, ny, (np.arange(n), vc[:, 1] * nx + vc[:, 0]))
    # triangles: every cell overlapped by the (slightly grown) bounding box
    tp = points_xy[tri_simplices]
    pad = 1e-9 * span
    c_lo = cell_xy(tp.min(axis=1) - pad)
    c_hi = cell_xy(tp.max(axis=1) + pad)
    items, cells = [], []
    wmax = int((c_hi - c_lo).max()) + 1 if len(tri_simplices) else 0     # no triangulation: mesh_refine alone
    for dx in range(wmax):
        for dy in range(wmax):
            cx, cy = c_lo[:, 0] + dx, c_lo[:, 1] + dy
            ok = (cx <= c_hi[:, 0]) & (cy <= c_hi[:, 1])
            items.append(np.flatnonzero(ok))
            cells.append((cy * nx + cx)[ok])
    if not items:
        items, cells = [np.zeros(0, dtype=np.int64)], [np.zeros(0, dtype=np.int64)]
    tstart, titems = _cell_lists(nx, ny, (np.concatenate(items), np.concatenate(cells)))
    # vertices of the 3 x 3 block of cells around each cell, cell-ordered: the nearest-vertex query reads one
    # contiguous list instead of walking nine cells
    items, cells = [], []
    vcx, vcy = vc[:, 0], vc[:, 1]
    for dx in (-1, 0, 1):
        for dy in (-1, 0, 1):
            cx, cy = vcx + dx, vcy + dy                  # the cell (cx, cy) has this vertex in its neighbourhood
            ok = (cx >= 0) & (cx < nx) & (cy >= 0) & (cy < ny)
            items.append(np.flatnonzero(ok))
            cells.append((cy * nx + cx)[ok])
    nstart, nitems = _cell_lists(nx, ny, (np.concatenate(items), np.concatenate(cells)))
    return {'nx': nx, 'ny': ny, 'x0': float(lo[0]), 'y0': float(lo[1]), 'inv_dx': float(inv[0]), 'inv_dy': float(inv[1]),
            'tri_start': tstart, 'tri_items': titems, 'vert_start': vstart, 'vert_items': vitems,
            'nb_start': nstart, 'nb_items': nitems}


def face_grid(points, faces, cells_per_axis=None):
    """
    Uniform xy grid over the mesh footprint with, per cell, the faces whose (slightly grown) xy bounding box touches the
    cell, plus the z range of the mesh: a ray can only hit a face inside the cells its xy track crosses while it is
    within that z range, so the full Moeller-Trumbore loop over every face (_ShapeMesh.py:289-348) can be restricted
    to those lists without changing its result.
    """
    pts = np.asarray(points, dtype=np.float64)
    faces = np.asarray(faces)
    lo, hi = pts[:, 0:2].min(axis=0), pts[:, 0:2].max(axis=0)
    if cells_per_axis is None:
        cells_per_axis = int(np.clip(np.sqrt(len(faces) / 2.0), 2, 1024))
    nx = ny = cells_per_axis
    span = np.maximum(hi - lo, 1e-300)
    inv = np.array([nx, ny]) / span

    def cell_xy(xy):
        c = np.floor((xy - lo) * inv).astype(np.int64)
        c[:, 0] = np.clip(c[:, 0], 0, nx - 1)
        c[:, 1] = np.clip(c[:, 1], 0, ny - 1)
        return c
    fp = pts[faces][:, :, 0:2]
    pad = 1e-9 * span
    c_lo, c_hi = cell_xy(fp.min(axis=1) - pad), cell_xy(fp.max(axis=1) + pad)
    items, cells = [], []
    wmax = int((c_hi - c_lo).max()) + 1
    for dx in range(wmax):
        for dy in range(wmax):
            cx, cy = c_lo[:, 0] + dx, c_lo[:, 1] + dy
            ok = (cx <= c_hi[:, 0]) & (cy <= c_hi[:, 1])
            items.append(np.flatnonzero(ok))
            cells.append((cy * nx + cx)[ok])
    start, fitems = _cell_lists(nx, ny, (np.concatenate(items), np.concatenate(cells)))
    zpad = 1e-9 * max(float(pts[:, 2].max() - pts[:, 2].min()), float(span.max()))
    return {'nx': nx, 'ny': ny, 'x0': float(lo[0]), 'y0': float(lo[1]), 'inv_dx': float(inv[0]), 'inv_dy': float(inv[1]),
            'start': start, 'items': fitems, 'z_min': float(pts[:, 2].min() - zpad), 'z_max': float(pts[:, 2].max() + zpad)}


def face_geometry(points, faces):
    """Moeller-Trumbore operands per face: p0, p1 - p0, p2 - p0 (_ShapeMesh.py:301-316)."""
    p0, p1, p2 = points[faces[:, 0]], points[faces[:, 1]], points[faces[:, 2]]
    return np.ascontiguousarray(np.concatenate([p0, p1 - p0, p2 - p0], axis=1), dtype=np.float64)


def face_area(points, faces):
    """|(p0 - p1) x (p0 - p2)| per face: the constant of the area-sum inside test (_ShapeMesh.py:405-409)."""
    p0, p1, p2 = points[faces[:, 0]], points[faces[:, 1]], points[faces[:, 2]]
    return np.ascontiguousarray(np.linalg.norm(np.cross(p0 - p1, p0 - p2), axis=1), dtype=np.float64)


def device_tables(param):
    """Everything fill_mesh uploads, as numpy arrays (also used by the CPU tests of the tables)."""
    m = param['mesh']
    t = {'points': np.ascontiguousarray(m['points'], dtype=np.float64),
         'face_geom': face_geometry(np.asarray(m['points'], dtype=np.float64), np.asarray(m['faces'])),
         'face_area': face_area(np.asarray(m['points'], dtype=np.float64), np.asarray(m['faces'])),
         'faces': np.ascontiguousarray(m['faces'], dtype=np.int32),
         'face_normals': np.ascontiguousarray(m['faces_normal'], dtype=np.float64),
         'point_faces': np.ascontiguousarray(m['p_faces_idx'], dtype=np.int32),
         'point_faces_mask': np.ascontiguousarray(m['p_faces_mask'], dtype=np.uint8)}
    t['vertex_faces'] = np.ascontiguousarray(np.where(m['p_faces_mask'], m['p_faces_idx'], -1).T, dtype=np.int32)
    # record of the candidate-face test (csrc/xrt_mesh.cuh:mesh_test_face): p0, m1 = e1 x n, m2 = e2 x n, unit normal n,
    # area |e1 x e2|, (face index, filled per vertex), A0 = (e1 x e2) . n.  With a = P - p0 the three sub-triangle areas of
    # the reference's inside test (_ShapeMesh.py:350-426) are |a . m1|, |a . m2| and |A0 + a . m1 - a . m2|.
    rec = np.zeros((len(t['faces']), 16))
    e1, e2, nrm = t['face_geom'][:, 3:6], t['face_geom'][:, 6:9], t['face_normals']
    rec[:, 0:3], rec[:, 3:6], rec[:, 6:9] = t['face_geom'][:, 0:3], np.cross(e1, nrm), np.cross(e2, nrm)
    rec[:, 9:12], rec[:, 12] = nrm, t['face_area']
    rec[:, 14] = np.einsum('ij,ij->i', np.cross(e1, e2), nrm)
    t['face_rec'] = rec
    refine = bool(param['mesh_refine'])
    if refine:
        mc = param['mesh_coarse']
        t['coarse_points'] = np.ascontiguousarray(mc['points'], dtype=np.float64)
        t['coarse_faces'] = np.ascontiguousarray(mc['faces'], dtype=np.int32)
        t['coarse_geom'] = face_geometry(t['coarse_points'], t['coarse_faces'])
    interp = bool(param['mesh_interpolate'])
    tri = None
    if interp:
        tri = m['interp']['z'].tri
        transform = barycentric_transforms(tri.points, tri.simplices)
        coef = []
        for key in ('z', 'normal_x', 'normal_y', 'normal_z'):
            ip = m['interp'][key]
            coef.append(ct_coefficients(ip.tri, np.asarray(ip.values)[:, 0], np.asarray(ip.grad)[:, 0, :], transform))
        t['ct_coef'] = np.ascontiguousarray(np.stack(coef, axis=1), dtype=np.float64)      # (n_tri, 4, 19)
        t['tri_transform'] = np.ascontiguousarray(transform, dtype=np.float64)             # (n_tri, 3, 2)
        simplices = tri.simplices
    else:
        simplices = np.zeros((0, 3), dtype=np.int32)
    if interp or refine:
        t['grid'] = lookup_grids(t['points'][:, 0:2], simplices)
        xyz = np.zeros((len(t['grid']['vert_items']), 4))
        xyz[:, 0:3] = t['points'][t['grid']['vert_items']]
        t['grid']['vert_xyz'] = xyz
        # 3 x 3 neighbourhood lists with the coordinates and the vertex index inline (32-byte records)
        nb = np.zeros((len(t['grid']['nb_items']), 4))
        nb[:, 0:3] = t['points'][t['grid']['nb_items']]
        nb[:, 3] = t['grid']['nb_items']
        t['grid']['nb_rec'] = nb
        # triangle records per cell: barycentric transform (6) + triangle index inline (64-byte records)
        if interp:
            tr = np.zeros((len(t['grid']['tri_items']), 8))
            tr[:, 0:6] = t['tri_transform'].reshape(-1, 6)[t['grid']['tri_items']]
            tr[:, 6] = t['grid']['tri_items']
            t['grid']['tri_rec'] = tr
    lossless = bool(param.get('mesh_lossless'))
    if (not refine or lossless) and len(t['faces']) > 64:
        t['face_grid'] = face_grid(t['points'], t['faces'])
    t['lossless'] = lossless
    if refine:
        # the <= 8 candidate faces of every vertex as consecutive 128-byte records (area < 0: no face), so that the
        # candidate loop needs no index hop and can fetch the next record while it tests the current one
        vf = t['vertex_faces']
        rec8 = np.zeros((len(vf), 8, 16))
        rec8[:, :, 12] = -1.0
        ok = vf >= 0
        rec8[ok] = t['face_rec'][vf[ok]]
        rec8[:, :, 13] = vf
        t['vertex_face_rec'] = rec8
    return t


def fill_mesh(param, keep):
    """-> (POINTER(XrtMesh), optic flag bits).  The tables are built once per prepared optic."""
    t = param.get('_device_tables')
    if t is None:
        t = param['_device_tables'] = device_tables(param)
    m = L.XrtMesh()
    flags = 0
    m.n_points, m.n_faces = len(t['points']), len(t['faces'])
    m.points = keep.f64(t['points'])
    m.faces = keep.arr(t['faces'], np.int32, C.c_int32)
    m.face_normals = keep.f64(t['face_normals'])
    m.face_geom = keep.f64(t['face_geom'])
    m.face_area = keep.f64(t['face_area'])
    m.face_rec = keep.f64(t['face_rec'])
    m.vertex_faces = keep.arr(t['vertex_faces'], np.int32, C.c_int32)
    m.point_faces = keep.arr(t['point_faces'], np.int32, C.c_int32)
    m.point_faces_mask = keep.arr(t['point_faces_mask'], np.uint8, C.c_uint8)
    if t.get('lossless'):
        flags |= L.F_MESH_LOSSLESS
    if 'face_grid' in t:
        fg = t['face_grid']
        m.fgrid_nx, m.fgrid_ny = fg['nx'], fg['ny']
        m.fgrid_x0, m.fgrid_y0, m.fgrid_inv_dx, m.fgrid_inv_dy = fg['x0'], fg['y0'], fg['inv_dx'], fg['inv_dy']
        m.fgrid_z_min, m.fgrid_z_max = fg['z_min'], fg['z_max']
        m.fgrid_start = keep.arr(fg['start'], np.int32, C.c_int32)
        m.fgrid_items = keep.arr(fg['items'], np.int32, C.c_int32)
    if 'coarse_points' in t:
        flags |= L.F_MESH_REFINE
        m.n_coarse_points, m.n_coarse_faces = len(t['coarse_points']), len(t['coarse_faces'])
        m.coarse_points = keep.f64(t['coarse_points'])
        m.coarse_faces = keep.arr(t['coarse_faces'], np.int32, C.c_int32)
        m.coarse_geom = keep.f64(t['coarse_geom'])
    if 'ct_coef' in t:
        flags |= L.F_MESH_INTERP
        m.n_tri = len(t['ct_coef'])
        # device layout: coefficient-major, the four fields (z, nx, ny, nz) of one coefficient side by side
        m.ct_coef = keep.f64(np.ascontiguousarray(np.transpose(t['ct_coef'], (0, 2, 1))))
        m.tri_transform = keep.f64(t['tri_transform'])
    if 'grid' in t:
        g = t['grid']
        m.grid_nx, m.grid_ny = g['nx'], g['ny']
        m.grid_x0, m.grid_y0, m.grid_inv_dx, m.grid_inv_dy = g['x0'], g['y0'], g['inv_dx'], g['inv_dy']
        m.grid_start = keep.arr(g['tri_start'], np.int32, C.c_int32)
        m.grid_items = keep.arr(g['tri_items'], np.int32, C.c_int32)
        m.vgrid_start = keep.arr(g['vert_start'], np.int32, C.c_int32)
        m.vgrid_items = keep.arr(g['vert_items'], np.int32, C.c_int32)
        m.vgrid_xyz = keep.f64(g['vert_xyz'])
        m.nb_start = keep.arr(g['nb_start'], np.int32, C.c_int32)
        m.nb_rec = keep.f64(g['nb_rec'])
        if 'tri_rec' in g:
            m.tri_rec = keep.f64(g['tri_rec'])
    if 'vertex_face_rec' in t:
        m.vertex_face_rec = keep.f64(t['vertex_face_rec'])
    keep.obj(m)
    return C.pointer(m), flags
