# -*- coding: utf-8 -*-
"""
Mesh device tables on the CPU: the Clough-Tocher control coefficients, barycentric transforms
and lookup grids that xicsrt_b200.mesh.device_tables uploads, evaluated here with a numpy
transcription of the device algorithm and compared with scipy's own interpolator objects and
kd-tree (the objects the reference calls, _ShapeMesh.py:172-196, 464-475).
"""
import numpy as np
import pytest

from oracle import scenes
from xicsrt_b200 import config as xconfig, mesh as xmesh, scene as xscene


def optic_param(name, elem='crystal'):
    cfg = xconfig.get_config(xconfig.to_numpy(scenes.get(name)))
    return xscene.prepare(cfg)[4][elem]


def grid_cell(g, xy):
    cx = np.clip(np.floor((xy[:, 0] - g['x0']) * g['inv_dx']).astype(int), 0, g['nx'] - 1)
    cy = np.clip(np.floor((xy[:, 1] - g['y0']) * g['inv_dy']).astype(int), 0, g['ny'] - 1)
    return cx, cy


def find_triangle(t, xy):
    """Device algorithm: candidates of the point's cell, first one with all barycentrics in [-eps, 1+eps]."""
    g = t['grid']
    cx, cy = grid_cell(g, xy)
    out = np.full(len(xy), -1)
    eps = xmesh.SIMPLEX_EPS
    for i in range(len(xy)):
        if not (np.isfinite(xy[i]).all()):
            continue
        c = cy[i] * g['nx'] + cx[i]
        for tri in g['tri_items'][g['tri_start'][c]:g['tri_start'][c + 1]]:
            b = xmesh.barycentric(t['tri_transform'][tri:tri + 1], xy[i:i + 1])[0]
            if np.all(b >= -eps) and np.all(b <= 1 + eps):
                out[i] = tri
                break
    return out


@pytest.mark.parametrize('name', ['mesh_torus', 'mesh_sphere', 'mesh_cylinder', 'mesh_user_interp'])
def test_clough_tocher_tables_reproduce_scipy(name):
    param = optic_param(name)
    t = xmesh.device_tables(param)
    pts = t['points']
    rng = np.random.default_rng(3)
    lo, hi = pts[:, 0:2].min(axis=0), pts[:, 0:2].max(axis=0)
    xy = lo + (hi - lo) * (rng.random((3000, 2)) * 1.04 - 0.02)      # a little beyond the hull too
    xy = np.concatenate([xy, pts[::7, 0:2]])                          # and exactly on vertices
    tri = find_triangle(t, xy)
    inside = tri >= 0
    interp = param['mesh']['interp']
    for f, key in enumerate(('z', 'normal_x', 'normal_y', 'normal_z')):
        ref = interp[key](xy[:, 0], xy[:, 1])
        assert np.array_equal(np.isnan(ref), ~inside), f'{name}/{key}: hull membership differs'
        b = xmesh.barycentric(t['tri_transform'][tri[inside]], xy[inside])
        got = xmesh.ct_evaluate(t['ct_coef'][tri[inside], f, :], b)
        err = np.max(np.abs(got - ref[inside]))
        assert err < 5e-15, f'{name}/{key}: {err:.2e}'


@pytest.mark.parametrize('name', ['mesh_torus', 'mesh_cylinder', 'mesh_sphere'])
def test_nearest_vertex_grid_matches_kdtree(name):
    """Device algorithm: ring search over the vertex grid until no closer vertex can exist."""
    param = optic_param(name)
    t = xmesh.device_tables(param)
    g, pts = t['grid'], t['points']
    rng = np.random.default_rng(5)
    lo, hi = pts.min(axis=0), pts.max(axis=0)
    q = lo + (hi - lo) * (rng.random((1500, 3)) * 1.2 - 0.1)
    ref = param['mesh']['points_tree'].query(q)[1]
    cx, cy = grid_cell(g, q[:, 0:2])
    dx, dy = 1.0 / g['inv_dx'], 1.0 / g['inv_dy']
    for i in range(len(q)):
        best, best_d2 = -1, np.inf
        for ring in range(max(g['nx'], g['ny']) + 1):
            if ring > 0:
                # everything in this ring and beyond is at least this far away in the xy plane
                ox = min(q[i, 0] - (g['x0'] + (cx[i] - ring + 1) * dx), (g['x0'] + (cx[i] + ring) * dx) - q[i, 0])
                oy = min(q[i, 1] - (g['y0'] + (cy[i] - ring + 1) * dy), (g['y0'] + (cy[i] + ring) * dy) - q[i, 1])
                bound = max(min(ox, oy), 0.0)
                if bound * bound > best_d2:
                    break
            for yy in range(cy[i] - ring, cy[i] + ring + 1):
                for xx in range(cx[i] - ring, cx[i] + ring + 1):
                    if max(abs(xx - cx[i]), abs(yy - cy[i])) != ring or not (0 <= xx < g['nx'] and 0 <= yy < g['ny']):
                        continue
                    c = yy * g['nx'] + xx
                    for v in g['vert_items'][g['vert_start'][c]:g['vert_start'][c + 1]]:
                        d2 = np.sum((pts[v] - q[i])**2)
                        if d2 < best_d2:
                            best, best_d2 = v, d2
        assert best == ref[i], (i, best, ref[i])


def test_point_faces_table_matches_reference_loop():
    param = optic_param('mesh_torus')
    faces = param['mesh']['faces']
    idx, mask = xmesh.point_faces_table(len(param['mesh']['points']), faces)
    for p in range(0, len(param['mesh']['points']), 13):
        ref = np.nonzero(np.equal(faces, p))[0]
        assert mask[:, p].sum() == len(ref)
        assert np.array_equal(idx[:len(ref), p], ref)
        assert not idx[len(ref):, p].any()
