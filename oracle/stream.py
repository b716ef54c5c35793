# -*- coding: utf-8 -*-
"""
Random streams for the oracle (test infrastructure).

The reference seeds the *global* legacy numpy generator once per run
(``xicsrt/xicsrt_raytrace.py:111``) and then every draw site pulls from it in
program order.  ``np.random.RandomState(seed)`` yields the same MT19937
stream, so issuing the same calls in the same order reproduces the
reference's rays exactly.

Two facts this relies on (checked in ``tests/test_oracle_stream.py``):
  ``RandomState.uniform(lo, hi, n) == lo + (hi - lo) * random_sample(n)`` and
  ``RandomState.normal(mu, s, n)   == mu + s * standard_normal(n)``, bit for bit.

Every draw is optionally recorded together with the set of rays it was drawn
for, so the very same numbers can be injected into the CUDA path
(``xrt_trace_injected``), scattered by ray index.
"""
import numpy as np


class LegacyStream:
    def __init__(self, seed=None, record=False):
        self.rs = np.random.RandomState(seed)
        self.record = record
        self.log = []   # list of (site, kind, values, mask-or-None)

    def _note(self, site, kind, values, mask):
        if self.record and site is not None:
            self.log.append((site, kind, np.array(values, copy=True),
                             None if mask is None else np.array(mask, copy=True)))

    def uniform(self, lo, hi, n, site=None, mask=None):
        """U[lo, hi) of length n; records the raw U[0,1) numbers."""
        u = self.rs.random_sample(n)
        self._note(site, 'u01', u, mask)
        return lo + (hi - lo) * u

    def normal(self, mu, sigma, n, site=None, mask=None):
        """N(mu, sigma) of length n; records the standard normals."""
        z = self.rs.standard_normal(n)
        self._note(site, 'z', z, mask)
        return mu + sigma * z

    def mvn(self, mean, cov, n, site=None, mask=None):
        """multivariate_normal; records the *output* (post-covariance) rows."""
        x = self.rs.multivariate_normal(mean, cov, n)
        self._note(site, 'mvn', x, mask)
        return x

    def poisson(self, lam, site=None):
        k = self.rs.poisson(lam)
        self._note(site, 'poisson', np.array([k]), None)
        return k

    def shuffle(self, arr):
        self.rs.shuffle(arr)

    # ------------------------------------------------------------------
    def scattered(self, site, n_total, width=1):
        """
        Full-length array (NaN where no draw happened) of the values recorded
        at ``site``, scattered by the alive-mask that was current at the draw.
        """
        out = np.full((n_total, width) if width > 1 else (n_total,), np.nan)
        for s, kind, values, mask in self.log:
            if s != site:
                continue
            if mask is None:
                out[...] = values.reshape(out.shape)
            else:
                out[mask] = values
        return out
