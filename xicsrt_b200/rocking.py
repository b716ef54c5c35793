# -*- coding: utf-8 -*-
"""
Rocking-curve tables for ``rocking_type='file'``.

Reads an XOP ``diff_pat.dat`` file into the three columns the Bragg test uses
(reference ``xicsrt/tools/xicsrt_bragg.py:19-112`` for the file layout and
``xicsrt/optics/_InteractCrystal.py:151-178`` for the units and the mixing).
The table is uploaded once at scene creation; the per-ray lookup
(``np.interp`` with zero outside the table) runs in the kernel.

(The reference's reader raises NameError as shipped -- it logs through an
undefined ``m_log`` -- so this path cannot run there unpatched; the column
meaning is taken from its ``col_list``.)
"""
import os

import numpy as np

_cache = {}


def guess_filetype(filename):
    rootname, _ = os.path.splitext(os.path.basename(filename))
    return 'xop' if rootname == 'diff_pat' else None


def load_table(filename, filetype=None):
    """Returns {'dtheta' [rad], 'reflect_s', 'reflect_p'} (cached per file + mtime)."""
    if filetype is None:
        filetype = guess_filetype(filename)
    if filetype is None:
        raise Exception('Could not guess the filetype. Please use the filetype keyword.')
    if filetype == 'x0h':
        raise NotImplementedError(f'A reader for filetype {filetype} not yet implemented.')
    if filetype != 'xop':
        raise Exception(f'Filetype {filetype} not recognized.')

    key = (os.path.abspath(filename), os.path.getmtime(filename))
    if key not in _cache:
        data = np.loadtxt(filename, dtype=np.float64)
        # columns: dtheta_in [urad], dtheta_out, phase_p, phase_s, circular, reflect_p, reflect_s
        _cache[key] = {
            'dtheta': np.ascontiguousarray(data[:, 0] * 1e-6),
            'reflect_p': np.ascontiguousarray(data[:, 5]),
            'reflect_s': np.ascontiguousarray(data[:, 6]),
        }
    return _cache[key]
