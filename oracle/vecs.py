# -*- coding: utf-8 -*-
"""
oracle.vecs -- TEST INFRASTRUCTURE (see oracle/__init__.py).

Frame transforms of ``xicsrt/objects/_GeometryObject.py:113-168``.  ``R`` is
the 3x3 orientation whose rows are the element's x, y, z axes.
"""
import numpy as np


def to_local(R, v):
    """R . v  for each row of v  (reference einsum 'ji,ki->kj')."""
    return np.einsum('ji,ki->kj', R, v)


def to_external(R, v):
    """R^T . v  for each row of v  (reference einsum 'ij,ki->kj')."""
    return np.einsum('ij,ki->kj', R, v)


def point_to_local(param, p):
    return to_local(param['orientation'], p - param['origin'])


def point_to_external(param, p):
    return to_external(param['orientation'], p) + param['origin']


def unit(v):
    return v / np.linalg.norm(v, axis=1)[:, None]


def dot(a, b):
    return np.einsum('ij,ij->i', a, b)
