# -*- coding: utf-8 -*-
"""
oracle.quartic -- TEST INFRASTRUCTURE (see oracle/__init__.py).

Closed-form quartic roots in the *slot order* the torus intersection depends
on.  Restates the algorithm of ``xicsrt/tools/xicsrt_quartic.py:22-207``
(itself adapted from the MIT-licensed ``fqs`` solver): Ferrari's reduction
through one real root of the resolvent cubic (Cardano / trigonometric
branches), then two quadratics whose four roots come out as

    slot 0, 1 = roots of  x^2 + s x + (z0 + t)      (minus a/4)
    slot 2, 3 = roots of  x^2 - s x + (z0 - t)      (minus a/4)

each pair ordered (-sqrt, +sqrt).  Complex arithmetic is kept exactly as in
the reference because ``_ShapeTorus.py:164-167`` decides "no intersection" by
``imag != 0`` on these complex values.
"""

import numpy as np


def _signed_cbrt(x):
    out = np.zeros_like(x)
    pos = x >= 0
    out[pos] = x[pos] ** (1. / 3.)
    out[~pos] = -(-x[~pos]) ** (1. / 3.)
    return out


def resolvent_root(p, r, c):
    """
    One real root of  z^3 + p z^2 + r z + c = 0  (``multi_cubic(..., all_roots=False)``,
    xicsrt_quartic.py:54-162).
    """
    third = 1. / 3.
    a13 = p * third
    a2 = a13 * a13
    f = third * r - a2
    g = a13 * (2 * a2 - r) + c
    h = 0.25 * g * g + f * f * f

    triple = (f == 0) & (g == 0) & (h == 0)
    three_real = (~triple) & (h <= 0)
    one_real = (~triple) & (~three_real)

    z = np.zeros(len(p))
    z[triple] = -_signed_cbrt(c[triple])

    j = np.sqrt(-f[three_real])
    k = np.arccos(-0.5 * g[three_real] / (j * j * j))
    z[three_real] = 2 * j * np.cos(third * k) - a13[three_real]

    sh = np.sqrt(h[one_real])
    S = _signed_cbrt(-0.5 * g[one_real] + sh)
    U = _signed_cbrt(-0.5 * g[one_real] - sh)
    z[one_real] = (S + U) - a13[one_real]
    return z


def _quadratic_pair(a, b):
    """Roots of x^2 + a x + b (complex), ordered (-sqrt, +sqrt); :22-51."""
    half = -0.5 * a
    sq = np.sqrt(half * half - b + 0j)
    return half - sq, half + sq


def quartic_slots(a0, b0, c0, d0, e0):
    """Four complex root arrays of a0 x^4 + b0 x^3 + c0 x^2 + d0 x + e0; :165-207."""
    a, b, c, d = b0 / a0, c0 / a0, d0 / a0, e0 / a0
    q4 = 0.25 * a
    q42 = q4 * q4

    p = 3 * q42 - 0.5 * b
    q = a * q42 - b * q4 + 0.5 * c
    r = 3 * q42 * q42 - b * q42 + c * q4 - d

    z0 = resolvent_root(p, r, p * r - 0.5 * q * q)

    s = np.sqrt(2 * p + 2 * z0.real + 0j)
    t = np.zeros_like(s)
    flat = (s == 0)
    t[flat] = z0[flat] * z0[flat] + r[flat]
    t[~flat] = -q[~flat] / s[~flat]

    r0, r1 = _quadratic_pair(s, z0 + t)
    r2, r3 = _quadratic_pair(-s, z0 - t)
    return r0 - q4, r1 - q4, r2 - q4, r3 - q4
