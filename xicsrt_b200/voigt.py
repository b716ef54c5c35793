# -*- coding: utf-8 -*-
"""
Setup-time Voigt tables for the source wavelength sampler.

The reference samples a Voigt line by inverse-CDF lookup in a 1000-bin table
on a non-uniform grid (``xicsrt/tools/xicsrt_voigt.py:30-92``) and draws
``U(min cdf, max cdf)`` so the clipped tails are never hit
(``xicsrt_voigt.py:119-130``).  The table is built here once per (gamma,
sigma) on the host with scipy's Faddeeva function and uploaded; the per-ray
lookup (binary search + linear interpolation, ``np.interp`` semantics) runs in
the kernel.

Line-width conversions: ``xicsrt/sources/_XicsrtSourceGeneric.py:341-352``.
"""
import numpy as np
import scipy.constants as const
from scipy.special import wofz

C_LIGHT = const.physical_constants['speed of light in vacuum'][0]
AMU_KG = const.physical_constants['atomic mass unit-kilogram relationship'][0]
EV_J = const.physical_constants['electron volt-joule relationship'][0]


def doppler_sigma(temperature, mass_number, wavelength):
    """Gaussian sigma [A] of a thermally broadened line (T in eV, m in amu)."""
    return np.sqrt(temperature / mass_number / AMU_KG / C_LIGHT**2 * EV_J) * wavelength


def natural_gamma(linewidth, wavelength):
    """Lorentzian gamma [A] from a natural linewidth [1/s]."""
    return linewidth * wavelength**2 / (4 * np.pi * C_LIGHT * 1e10)


def voigt_profile(x, sigma, gamma):
    z = (x + 1j * gamma) / np.sqrt(2) / sigma
    return wofz(z).real / np.sqrt(2 * np.pi) / sigma


def cdf_table(gamma, sigma, gridsize=1000, cutoff=1e-5):
    """
    Returns (x, cdf): right bin edges and the cumulative sum of pdf*dx on the
    reference's stretched grid.
    """
    gridsize_min = 100
    fraction = 0.5
    gauss_hw = np.sqrt(2.0 * np.log(1.0 / fraction)) * sigma
    lorentz_hw = gamma * np.sqrt(1.0 / fraction - 1.0)
    hw_max = np.sqrt(gauss_hw**2 + lorentz_hw**2)

    value = gridsize_min / 2 * (hw_max / 5.0)

    lorentz_cut = gamma * np.sqrt(1.0 / cutoff - 1.0)
    gauss_cut = np.sqrt(-1 * sigma**2 * 2 * np.log(cutoff * sigma * np.sqrt(2 * np.pi)))
    base = np.exp(1 / 10 * np.log(max(lorentz_cut, gauss_cut) / value))

    edges = np.linspace(-value, value, gridsize + 1)
    edges = edges * base**np.abs(edges / value * 10)
    mid = (edges[:-1] + edges[1:]) / 2

    pdf = voigt_profile(mid, sigma, gamma)
    cdf = np.cumsum(pdf * (edges[1:] - edges[:-1]))

    if np.sum((cdf > 0.25) & (cdf < 0.75)) < 3:
        raise Exception('Voight CDF calculation does not have enough resolution.')
    if np.max(cdf) < 0.99:
        raise Exception('Voight CDF calculation domain too small.')
    return edges[1:], cdf


def wavelength_model(param):
    """
    Classify a source's wavelength distribution into what the kernel samples.

    Returns a dict with ``mode`` in {'const', 'uniform', 'normal', 'table'} and
    the numbers that mode needs.  Follows the branch order of
    ``_XicsrtSourceGeneric.py:295-354`` including the silent +1 eV when a
    Lorentzian-only line is requested (``:333-339``).
    """
    wtype = str.lower(param['wavelength_dist'])
    lam0 = float(param['wavelength'])
    if wtype == 'monochrome':
        return {'mode': 'const', 'wavelength': lam0}
    if wtype == 'uniform':
        rng = np.asarray(param['wavelength_range'], dtype=np.float64)
        return {'mode': 'uniform', 'lo': float(rng[0]), 'hi': float(rng[1])}
    if wtype != 'voigt':
        raise Exception(f'Wavelength distribution {wtype} unknown')

    linewidth = float(param['linewidth'])
    temperature = float(param['temperature'])
    if linewidth == 0.0 and temperature == 0.0:
        return {'mode': 'const', 'wavelength': lam0}
    if linewidth == 0.0:
        return {'mode': 'normal', 'wavelength': lam0,
                'sigma': float(doppler_sigma(temperature, param['mass_number'], lam0))}
    if temperature == 0.0:
        temperature = 1.0
    gamma = natural_gamma(linewidth, lam0)
    sigma = doppler_sigma(temperature, param['mass_number'], lam0)
    x, cdf = cdf_table(gamma, sigma)
    return {'mode': 'table', 'wavelength': lam0, 'x': x, 'cdf': cdf,
            'gamma': float(gamma), 'sigma': float(sigma)}
