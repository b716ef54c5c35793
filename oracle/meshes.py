# -*- coding: utf-8 -*-
"""
oracle.meshes -- TEST INFRASTRUCTURE (see oracle/__init__.py).

numpy / scipy restatement of the mesh intersection of ``xicsrt/optics/_ShapeMesh.py``:
Moeller-Trumbore over all faces (:289-348), coarse -> fine refinement through the nearest
fine vertex and its <= 8 faces (:464-475, :350-426), Clough-Tocher interpolation of z and the
normal with scipy's own interpolator objects (:172-196), flat face normals otherwise (:428-432).
The setup-time tables (``param['mesh']``, ``param['mesh_coarse']``) come from
``xicsrt_b200.mesh.initialize_mesh`` (shared host logic).
"""
import numpy as np


def moeller_trumbore(mesh, O, D, alive):
    """:289-348 -- every face in turn; no test on t; a later face overwrites an earlier hit."""
    faces, pts = mesh['faces'], mesh['points']
    p0, p1, p2 = pts[faces[..., 0], :], pts[faces[..., 1], :], pts[faces[..., 2], :]
    eps = 1e-15
    n = len(alive)
    X = np.full(D.shape, np.nan, dtype=np.float64)
    hits = np.zeros(n, dtype=np.int64)
    any_hit = np.zeros(n, dtype=bool)
    for ii in range(faces.shape[0]):
        ok = alive.copy()
        edge1 = p1[ii, :] - p0[ii, :]
        edge2 = p2[ii, :] - p0[ii, :]
        h = np.cross(D, edge2)
        f = np.einsum('i,ji->j', edge1, h)
        ok &= ~((f > -eps) & (f < eps))
        if not np.any(ok):
            continue
        with np.errstate(divide='ignore', invalid='ignore'):
            f = 1.0 / f
            s = O - p0[ii, :]
            u = f * np.einsum('ij,ij->i', s, h)
            ok &= ~((u < 0.0) | (u > 1.0))
            if not np.any(ok):
                continue
            q = np.cross(s, edge1)
            v = f * np.einsum('ij,ij->i', D, q)
            ok &= ~((v < 0.0) | (u + v > 1.0))
            if not np.any(ok):
                continue
            t = f * np.einsum('i,ji->j', edge2, q)
        any_hit[ok] = True
        hits[ok] = ii
        X[ok] = O[ok] + t[ok, None] * D[ok, :]
    return X, alive & any_hit, hits


def near_faces(mesh, X, alive):
    """:464-475 -- nearest fine vertex (3-D) of each coarse hit and the faces around it."""
    n = len(alive)
    idx = mesh['points_tree'].query(X[alive])[1]
    faces_idx = np.zeros((8, n), dtype=np.int32)
    faces_mask = np.zeros((8, n), dtype=np.bool_)
    faces_idx[:, alive] = mesh['p_faces_idx'][:, idx]
    faces_mask[:, alive] = mesh['p_faces_mask'][:, idx]
    return faces_idx, faces_mask


def candidate_faces(mesh, O, D, alive, faces_idx, faces_mask):
    """:350-426 -- ray/plane point per candidate face, inside test by area sum, first candidate wins."""
    n = len(alive)
    X = np.full(D.shape, np.nan, dtype=np.float64)
    hits = np.zeros(n, dtype=np.int64)
    faces = mesh['faces'][faces_idx]
    pts = mesh['points']
    p0, p1, p2 = pts[faces[..., 0], :], pts[faces[..., 1], :], pts[faces[..., 2], :]
    nrm = mesh['faces_normal'][faces_idx]
    with np.errstate(divide='ignore', invalid='ignore'):
        t0 = p0 - O[None, :, :]
        t1 = np.einsum('ijk,ijk->ij', t0, nrm)
        t2 = np.einsum('jk,ijk->ij', D, nrm)
        dist = t1 / t2
        inter = np.einsum('jk,ij->ijk', D, dist) + O
        a, b, c = inter - p0, inter - p1, inter - p2
        diff = (np.linalg.norm(np.cross(b, c), axis=2) + np.linalg.norm(np.cross(c, a), axis=2)
                + np.linalg.norm(np.cross(a, b), axis=2) - np.linalg.norm(np.cross((p0 - p1), (p0 - p2)), axis=2))
        test = (diff < 1e-10) & (dist >= 0) & faces_mask
    alive = alive & np.any(test, axis=0)
    which = np.argmax(test[:, alive], axis=0)
    hits[alive] = faces_idx[which, alive]
    X[alive] = inter[which, alive, :]
    return X, alive, hits


def intersect(param, O, D, alive):
    """ShapeMesh.intersect (:135-170).  Returns (X, normals, alive)."""
    mesh = param['mesh']
    if not param['mesh_refine'] or param.get('mesh_lossless'):
        # mesh_lossless (product option, not in the reference): the full test on the fine mesh = mesh_refine off
        X, alive, hits = moeller_trumbore(mesh, O, D, alive.copy())
    else:
        Xc, alive_c, _ = moeller_trumbore(param['mesh_coarse'], O, D, alive.copy())
        fidx, fmask = near_faces(mesh, Xc, alive_c)
        X, alive, hits = candidate_faces(mesh, O, D, alive_c, fidx, fmask)
    if param['mesh_interpolate']:
        ip = mesh['interp']
        X[:, 2] = ip['z'](X[:, 0], X[:, 1])
        nrm = np.empty(X.shape)
        nrm[:, 0] = ip['normal_x'](X[:, 0], X[:, 1])
        nrm[:, 1] = ip['normal_y'](X[:, 0], X[:, 1])
        nrm[:, 2] = ip['normal_z'](X[:, 0], X[:, 1])
        with np.errstate(invalid='ignore'):
            nrm = np.einsum('i,ij->ij', 1.0 / np.linalg.norm(nrm, axis=1), nrm)
    else:
        nrm = np.zeros((len(alive), 3), dtype=np.float64)
        nrm[alive, :] = mesh['faces_normal'][hits[alive], :]
    return X, nrm, alive
