# -*- coding: utf-8 -*-
"""
oracle.optics -- TEST INFRASTRUCTURE (see oracle/__init__.py).

numpy restatement of one optic stage: intersect -> bounds/aperture ->
interaction -> image, for the analytic shapes.  Mesh shapes live in
oracle/meshes.py.  Each function cites the reference lines it follows.
"""
import numpy as np

from oracle import vecs
from oracle import quartic


# ---------------------------------------------------------------------------
# ray/surface distance.  All return (t, alive) with t = NaN where not alive.

def hit_plane(param, O, D, alive):
    """_ShapePlane.py:32-53 -- t = ((o - O).z)/(D.z), keep t >= 0."""
    t = np.full(alive.shape, np.nan)
    if param['trace_local']:
        z = np.array([0.0, 0.0, 1.0])
        t[alive] = np.dot(0 - O[alive], z) / np.dot(D[alive], z)
    else:
        t[alive] = np.dot(param['origin'] - O[alive], param['zaxis']) / np.dot(D[alive], param['zaxis'])
    with np.errstate(invalid='ignore'):
        alive = alive & (t >= 0)
    return t, alive


def _pick_root(t_lo, t_hi, convex):
    if convex:
        return np.where(t_lo < t_hi, t_lo, t_hi)
    return np.where(t_lo > t_hi, t_lo, t_hi)


def hit_sphere(param, O, D, alive):
    """_ShapeSphere.py:53-100 -- geometric solution, far root if concave, no sign test on t."""
    R = param['radius']
    alive = alive.copy()
    t = np.full(alive.shape, np.nan)
    L = param['center'] - O
    tca = vecs.dot(L, D)
    with np.errstate(invalid='ignore'):
        d = np.sqrt(vecs.dot(L[alive], L[alive]) - tca[alive]**2)
        inside = d <= R
    idx = np.flatnonzero(alive)
    alive[idx] = inside
    d = d[inside]
    thc = np.sqrt(R**2 - d**2)
    t[alive] = _pick_root(tca[alive] - thc, tca[alive] + thc, param['convex'])
    return t, alive


def hit_cylinder(param, O, D, alive):
    """_ShapeCylinder.py:52-109 -- quadratic in the components normal to the x axis."""
    pa, va, r = param['center'], param['xaxis'], param['radius']
    alive = alive.copy()
    t = np.full(alive.shape, np.nan)
    dp = O - pa
    A1 = D - np.einsum('ij,j->i', D, va)[:, None] * va[None, :]
    B1 = dp - np.einsum('ij,j->i', dp, va)[:, None] * va[None, :]
    A = vecs.dot(A1, A1)
    B = 2 * vecs.dot(A1, B1)
    C = vecs.dot(B1, B1) - r**2
    dis = B**2 - 4 * A * C
    with np.errstate(invalid='ignore'):
        ok = dis >= 0
    alive &= ok
    sq = np.sqrt(dis[alive])
    t[alive] = _pick_root((-B[alive] - sq) / (2 * A[alive]), (-B[alive] + sq) / (2 * A[alive]), param['convex'])
    return t, alive


def hit_torus(param, O, D, alive):
    """
    _ShapeTorus.py:116-183 -- quartic in torus coordinates (axis = local y),
    solved by the fqs Ferrari/Cardano routine; the root is taken by solver
    *slot* ``root_idx``; complex slots become NaN; keep finite t > 0.
    """
    rmin, rmaj = param['torus_minor'], param['torus_major']
    Rm = param['orientation']
    Ot = vecs.to_local(Rm, O - param['center'])
    Dt = vecs.to_local(Rm, D)

    OO = vecs.dot(Ot, Ot)
    OD = vecs.dot(Ot, Dt)
    rsq = rmaj**2 + rmin**2
    ones = np.ones(len(alive))
    c0 = ones**4
    c1 = 4 * ones**2 * OD
    c2 = 4 * OD**2 + 2 * OO * ones**2 - 2 * rsq * ones**2 + 4 * rmaj**2 * Dt[:, 1]**2
    c3 = 4 * OD * (OO - rsq) + 8 * rmaj**2 * Dt[:, 1] * Ot[:, 1]
    c4 = OO**2 - 2 * rsq * OO + 4 * rmaj**2 * Ot[:, 1]**2 + (rmaj**2 - rmin**2)**2

    with np.errstate(all='ignore'):
        slots = quartic.quartic_slots(c0, c1, c2, c3, c4)
    root = slots[param['root_idx']]
    t = np.where(root.imag != 0, np.nan, root.real)
    alive = alive.copy()
    with np.errstate(invalid='ignore'):
        alive &= np.isfinite(t) & (t > 0.0)
    return t, alive


def point_on_ray(O, D, t, alive):
    """_ShapeObject.py:69-81 -- X = O + D t on alive rays, NaN elsewhere."""
    X = np.full(O.shape, np.nan)
    X[alive] = O[alive] + D[alive] * t[alive, None]
    return X


# ---------------------------------------------------------------------------
# surface normals (NaN / zero rows where not alive, as in the reference)

def normal_plane(param, X, alive):
    """_ShapePlane.py:55-62."""
    n = np.full(X.shape, np.nan)
    n[alive] = param['zaxis']
    return n


def normal_sphere(param, X, alive):
    """_ShapeSphere.py:102-106."""
    n = np.full(X.shape, np.nan)
    n[alive] = vecs.unit(param['center'] - X[alive])
    return n


def normal_cylinder(param, X, alive):
    """_ShapeCylinder.py:111-133 -- toward the axis point at the same x."""
    pa, va = param['center'], param['xaxis']
    n = np.full(X.shape, np.nan)
    s = np.einsum('ij,j->i', pa - X[alive], va)
    foot = pa - s[:, None] * va[None, :]
    n[alive] = vecs.unit(foot - X[alive])
    return n


def normal_torus(param, X, alive):
    """_ShapeTorus.py:186-216 -- from the nearest point of the torus axis circle."""
    C = param['center']
    yaxis = np.cross(param['zaxis'], param['xaxis'])
    n = np.zeros(X.shape)
    p = X[alive] - C
    p = p - np.einsum('i,j->ij', np.einsum('ij,j->i', p, yaxis), yaxis)
    Q = C + param['torus_major'] * vecs.unit(p)
    n[alive] = vecs.unit(X[alive] - Q)
    return n


SHAPES = {
    'plane': (hit_plane, normal_plane),
    'sphere': (hit_sphere, normal_sphere),
    'cylinder': (hit_cylinder, normal_cylinder),
    'torus': (hit_torus, normal_torus),
}


# ---------------------------------------------------------------------------
# bounds and apertures

def tri_inside(pt, p0, p1, p2):
    """xicsrt/tools/xicsrt_math.py:290-307 (barycentric, >= 0)."""
    area = 0.5 * (-p1[1] * p2[0] + p0[1] * (-p1[0] + p2[0]) + p0[0] * (p1[1] - p2[1]) + p1[0] * p2[1])
    a = 1 / (2 * area) * (p0[1] * p2[0] - p0[0] * p2[1] + (p2[1] - p0[1]) * pt[:, 0] + (p0[0] - p2[0]) * pt[:, 1])
    b = 1 / (2 * area) * (p0[0] * p1[1] - p0[1] * p1[0] + (p0[1] - p1[1]) * pt[:, 0] + (p1[0] - p0[0]) * pt[:, 1])
    c = 1 - a - b
    return (a >= 0) & (b >= 0) & (c >= 0)


def aperture_inside(ap, x, y):
    """Shape tests of xicsrt/tools/xicsrt_aperture.py:110-204 on points (x, y)."""
    shape = (ap.get('shape') or 'none').lower()
    org = np.atleast_1d(np.asarray(ap.get('origin') if ap.get('origin') is not None else [0.0, 0.0], dtype=np.float64))
    size = np.atleast_1d(np.asarray(ap['size'], dtype=np.float64)) if 'size' in ap else None
    if shape == 'none':
        return np.ones(x.shape, dtype=bool)
    if shape == 'circle':
        return ((x - org[0])**2 + (y - org[1])**2) < size[0]**2
    if shape == 'square':
        return (np.abs(x - org[0]) < size[0] / 2) & (np.abs(y - org[1]) < size[0] / 2)
    if shape == 'rectangle':
        return (np.abs(x - org[0]) < size[0] / 2) & (np.abs(y - org[1]) < size[1] / 2)
    if shape == 'ellipse':
        # full sizes used as semi-axes, as the reference does (:182)
        return (((x - org[0]) / size[0])**2 + ((y - org[1]) / size[1])**2) < 1
    if shape == 'triangle':
        v = np.asarray(ap['vertices'], dtype=np.float64)
        return tri_inside(np.stack([x, y], axis=1), v[0, 0:2] + org[0:2], v[1, 0:2] + org[0:2], v[2, 0:2] + org[0:2])
    raise Exception(f'Aperture shape: "{shape}" is not implemented.')


def aperture_list(aperture):
    if aperture is None:
        return []
    if isinstance(aperture, dict):
        return [aperture]
    return list(np.atleast_1d(np.asarray(aperture, dtype=object)))


def aperture_fold(Xl, alive, aperture):
    """
    xicsrt_aperture.py:13-49 -- sequential fold of the aperture list; every
    logic op acts only on the rays alive at entry.
    """
    aps = aperture_list(aperture)
    if not aps:
        return alive
    idx = np.flatnonzero(alive)
    x, y = Xl[idx, 0], Xl[idx, 1]
    acc = np.ones(len(idx), dtype=bool)
    for ap in aps:
        test = aperture_inside(ap, x, y)
        logic = (ap.get('logic') or 'and').lower()
        if logic == 'and':
            acc = acc & test
        elif logic == 'not':
            acc = acc & ~test
        elif logic == 'or':
            acc = acc | test
        elif logic == 'nand':
            acc = ~(acc & test)
        elif logic == 'nor':
            acc = ~(acc | test)
        elif logic == 'xor':
            acc = acc ^ test
        elif logic == 'xnor':
            acc = ~(acc ^ test)
        else:
            raise Exception(f'Aperture logic "{logic}" is not known.')
    out = alive.copy()
    out[idx] = acc
    return out


def within_bounds(param, X, alive):
    """_TraceObject.py:180-232 -- strict |x| < size/2 per axis, then apertures."""
    if param['trace_local']:
        Xl = X
    else:
        Xl = np.zeros(X.shape)
        Xl[alive] = vecs.point_to_local(param, X[alive])
    alive = alive.copy()
    if param['check_size']:
        for ax, key in enumerate(('xsize', 'ysize', 'zsize')):
            if param[key] is not None:
                idx = np.flatnonzero(alive)
                alive[idx] = np.abs(Xl[idx, ax]) < param[key] / 2
    if param['check_aperture']:
        alive = aperture_fold(Xl, alive, param['aperture'])
    return alive


# ---------------------------------------------------------------------------
# interactions

def mirror(D, n, sel):
    """_InteractMirror.py:29-42 -- D -= 2 (D.n) n on the selected rays."""
    D[sel] -= 2 * (vecs.dot(D[sel], n[sel])[:, None] * n[sel])


def bragg_angles(param, D, W, n, sel):
    """_InteractCrystal.py:96-116."""
    tb = np.zeros(sel.shape)
    ti = np.zeros(sel.shape)
    tb[sel] = np.arcsin(W[sel] / (2 * param['crystal_spacing']))
    dot = np.abs(vecs.dot(D[sel], -1 * n[sel]))
    ti[sel] = (np.pi / 2) - np.arccos(dot / np.linalg.norm(D[sel], axis=1))
    return tb, ti


def rocking_probability(param, ti, tb):
    """_InteractCrystal.py:139-184 (without the reflectivity factor)."""
    kind = param['rocking_type']
    if 'step' in kind:
        return np.where(np.abs(ti - tb) <= param['rocking_fwhm'] / 2, 1.0, 0.0)
    if 'gauss' in kind:
        sigma = param['rocking_fwhm'] / (2 * np.sqrt(2 * np.log(2)))
        return np.exp(-np.power(ti - tb, 2.) / (2 * sigma**2))
    if 'file' in kind:
        from xicsrt_b200 import rocking
        tab = rocking.load_table(param['rocking_file'], param['rocking_filetype'])
        dth = ti - tb
        s = np.interp(dth, tab['dtheta'], tab['reflect_s'], left=0.0, right=0.0)
        p = np.interp(dth, tab['dtheta'], tab['reflect_p'], left=0.0, right=0.0)
        return param['rocking_mix'] * s + (1 - param['rocking_mix']) * p
    raise Exception('Rocking curve type not understood: {}'.format(kind))


def bragg_filter(param, D, W, n, sel, stream, site):
    """
    _InteractCrystal.py:118-196 -- keep rays whose rocking-curve probability
    is >= a fresh U[0,1) draw; one draw per *selected* ray, in ray order.
    """
    if param['check_bragg'] is False:
        return sel
    tb, ti = bragg_angles(param, D, W, n, sel)
    p = rocking_probability(param, ti[sel], tb[sel])
    p = p * param['reflectivity']
    u = stream.uniform(0.0, 1.0, int(np.sum(sel)), site=site, mask=sel)
    out = sel.copy()
    out[sel] = p >= u
    return out


def mosaic_normal(param, n, sel, stream, site):
    """
    _InteractMosaicCrystal.py:109-139 with xicsrt_spread.py:297-339 --
    (x, y) ~ N(0, sin^2(sigma)) on the z = 1 plane about the nominal normal.
    """
    hwhm = param['mosaic_spread'] / 2.0
    sigma = hwhm / np.sqrt(2 * np.log(2))
    s = np.sin(sigma)
    k = int(np.sum(sel))
    xy = stream.mvn([0, 0], [[s**2, 0], [0, s**2]], k, site=site, mask=sel)
    loc = np.stack([xy[:, 0], xy[:, 1], np.full(k, 1.0)], axis=1)
    loc = loc * (1 / np.linalg.norm(loc, axis=1))[:, None]

    nn = n[sel]
    r0 = np.cross(nn, [1, 0, 0]) + np.cross(nn, [0, 0, 1])
    r0 /= np.linalg.norm(r0, axis=1)[:, None]
    r1 = np.cross(nn, r0)
    r1 /= np.linalg.norm(r1, axis=1)[:, None]
    out = n.copy()
    out[sel] = loc[:, 0:1] * r0 + loc[:, 1:2] * r1 + loc[:, 2:3] * nn
    return out


def interact(param, rays, X, n, alive, stream, tag):
    """
    Dispatch on the interaction kind.  Returns the final alive mask; updates
    rays['origin'] / rays['direction'] in place.

    none   : _InteractObject.py:25-40
    mirror : _InteractMirror.py:24-42
    crystal: _InteractCrystal.py:90-94
    mosaic : _InteractMosaicCrystal.py:53-107
    """
    kind = param['_interact']
    O, D, W = rays['origin'], rays['direction'], rays['wavelength']
    if kind == 'none':
        O[:] = X
        return alive
    if kind == 'mirror':
        O[:] = X
        mirror(D, n, alive)
        return alive
    if kind == 'crystal':
        alive = bragg_filter(param, D, W, n, alive, stream, f'{tag}.u.0')
        O[:] = X
        mirror(D, n, alive)
        return alive
    if kind == 'mosaic':
        alive = alive.copy()
        if param['mosaic_cutoff'] is not None:
            tb, ti = bragg_angles(param, D, W, n, alive)
            sig = param['mosaic_spread'] / (2 * np.sqrt(2 * np.log(2)))
            cut = np.sqrt(-1 * np.log(param['mosaic_cutoff']) * 2 * sig**2)
            alive[alive] = np.abs(tb[alive] - ti[alive]) < cut
        if np.sum(alive) > 0:
            done = np.zeros(alive.shape, dtype=np.bool_)
            for layer in range(param['mosaic_depth']):
                cand = (~done) & alive
                if np.sum(cand) == 0:
                    break
                nm = mosaic_normal(param, n, cand, stream, f'{tag}.xy.{layer}')
                cand = bragg_filter(param, D, W, nm, cand, stream, f'{tag}.u.{layer}')
                O[:] = X
                mirror(D, nm, cand)
                done[cand] = True
            alive &= done
        return alive
    raise KeyError(kind)


# ---------------------------------------------------------------------------
# pixel binning

def bin_image(param, O, alive):
    """
    _TraceObject.py:234-293 -- channel = rint(local/pixel + (npix-1)/2)
    (round-half-even), hits outside the grid are dropped.
    """
    if not param['enable_image']:
        return None
    nx, ny = param['pixel_xsize'], param['pixel_ysize']
    img = np.zeros((nx, ny))
    if np.sum(alive) > 0:
        pix = vecs.point_to_local(param, O[alive]) / param['pixel_size']
        cx = np.round(pix[:, 0] + (nx - 1) / 2).astype(int)
        cy = np.round(pix[:, 1] + (ny - 1) / 2).astype(int)
        ok = (cx >= 0) & (cx < nx) & (cy >= 0) & (cy < ny)
        np.add.at(img, (cx[ok], cy[ok]), 1.0)
    return img


# ---------------------------------------------------------------------------
# one optic, start to end

def trace_optic(param, rays, stream, tag):
    """
    _TraceObject.py:135-178 -- optional to-local, intersect, bounds, interact,
    optional to-external.  ``rays`` is updated in place and returned.
    """
    local = bool(param['trace_local'])
    Rm = param['orientation']
    if local:
        rays['origin'] = vecs.to_local(Rm, rays['origin'] - param['origin'])
        rays['direction'] = vecs.to_local(Rm, rays['direction'])

    O, D = rays['origin'], rays['direction']
    alive = rays['mask'].copy()
    shape = param['_shape']
    if shape.startswith('mesh'):
        from oracle import meshes
        X, n, alive = meshes.intersect(param, O, D, alive)
    else:
        hit, normal = SHAPES[shape]
        t, alive = hit(param, O, D, alive)
        X = point_on_ray(O, D, t, alive)
        n = normal(param, X, alive)

    alive = within_bounds(param, X, alive)
    alive = interact(param, rays, X, n, alive, stream, tag)
    rays['mask'] = alive

    if local:
        rays['origin'] = vecs.to_external(Rm, rays['origin']) + param['origin']
        rays['direction'] = vecs.to_external(Rm, rays['direction'])
    return rays
