# -*- coding: utf-8 -*-
"""
The device transcendentals (xicsrt_b200/csrc/xrt_fastmath.cuh) are __host__ __device__: the same
source is compiled for the host here and compared with long double libm over the argument
ranges the ray code uses.  Bar: a few 1e-16 (the parity tolerance of the ray states is 1e-9).
"""
import os
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

PROGRAM = r'''
#include "xrt_fastmath.cuh"
#include <cstdio>
#include <cmath>
#include <random>
using namespace xrt;
int main() {
    std::mt19937_64 g(12345);
    std::uniform_real_distribution<double> U(0.0, 1.0);
    const long double PI = 3.141592653589793238462643383279502884L;
    double es = 0, ec = 0, ec2 = 0, el = 0, ee = 0, ea = 0, et = 0;
    static double tab[2 * kSincosTable];
    for (int i = 0; i < kSincosTable; ++i) sincos_2pi((double)i / kSincosTable, tab[2 * i + 1], tab[2 * i]);
    for (int i = 0; i < 2000000; ++i) {
        double u = U(g);
        if (i < 9) u = i * 0.125;                       // quadrant boundaries
        double s, c;
        sincos_2pi(u, s, c);
        long double rs = sinl(2 * PI * (long double)u), rc = cosl(2 * PI * (long double)u);
        es = fmax(es, fabs((double)(s - rs)));
        ec = fmax(ec, fabs((double)(c - rc)));
        ec2 = fmax(ec2, fabs((double)(cos_2pi(u) - rc)));
        double st, ct;
        sincos_2pi_tab(u, tab, st, ct);
        et = fmax(et, fmax(fabs((double)(st - rs)), fabs((double)(ct - rc))));
        double v = 1.0 - u;
        if (v <= 0) v = 1.1e-16;
        if (i % 3 == 0) v = ldexp(v, -(i % 53));        // down to 2^-53
        long double rl = logl((long double)v);
        if (rl != 0) el = fmax(el, fabs((double)((log_pos(v) - rl) / rl)));
        double x = (i % 2) ? u * 45.0 : u * 700.0;
        long double re = expl(-(long double)x);
        ee = fmax(ee, fabs((double)((exp_neg(x) - re) / re)));
        double w = (u - 0.5) * 0.1;
        if (w != 0) ea = fmax(ea, fabs((double)((asin_small(w) - asinl((long double)w)) / asinl((long double)w))));
    }
    // inverse normal CDF: one Newton step in long double from the value under test (erfcl on the tail that holds u)
    double en = 0, en_tail = 0;
    const long double SQ2 = 1.41421356237309504880168872420969808L, SQ2PI = 2.50662827463100050241576528481104525L;
    for (int i = 0; i < 600000; ++i) {
        double u = U(g);
        if (i % 4 == 1) u = ldexp(u + 0.5, -(1 + i % 53));          // lower tail down to 2^-54
        if (i % 4 == 2) u = 1.0 - ldexp(u + 0.5, -(1 + i % 52));    // upper tail up to 1 - 2^-53
        if (i % 4 == 3) u = 0.5 + (u - 0.5) * ldexp(1.0, -(i % 50)); // around the median
        if (!(u > 0.0 && u < 1.0)) continue;
        const double z = inv_normal_cdf(u);
        const long double zl = z;
        long double f;                                               // Phi(z) - u, formed on the smaller tail
        if (u < 0.5) f = 0.5L * erfcl(-zl / SQ2) - (long double)u;
        else f = ((long double)1.0 - (long double)u) - 0.5L * erfcl(zl / SQ2);
        const long double zr = zl - f * SQ2PI * expl(0.5L * zl * zl);
        if (zr != 0) {
            const double e = fabs((double)((zl - zr) / zr));
            en = fmax(en, e);
            if (u < 1e-6 || u > 1.0 - 1e-6) en_tail = fmax(en_tail, e);
        }
    }
    const double zmid = inv_normal_cdf(0.5), zlo = inv_normal_cdf(ldexp(1.0, -54)), zhi = inv_normal_cdf(1.0 - ldexp(1.0, -53));
    double s0, c0, s1, c1;
    sincos_2pi(0.0, s0, c0);
    sincos_2pi(1.0, s1, c1);
    printf("%.3e %.3e %.3e %.3e %.3e %.3e %g %g %g %g %g %g %.3e %.3e %.3e %.17g %.17g %.17g\n", es, ec, ec2, el, ee, ea,
           s0, c0, s1, c1, log_pos(1.0), exp_neg(0.0), et, en, en_tail, zmid, zlo, zhi);
    return 0;
}
'''


@pytest.mark.timeout(300)
def test_device_math_against_long_double(tmp_path):
    nvcc = shutil.which('nvcc') or '/usr/local/cuda/bin/nvcc'
    if not os.path.exists(nvcc):
        pytest.skip('nvcc not available')
    src = tmp_path / 'fm.cu'
    src.write_text(PROGRAM)
    exe = tmp_path / 'fm'
    subprocess.run([nvcc, '-O2', '-Wno-deprecated-gpu-targets', '-I', os.path.join(ROOT, 'xicsrt_b200', 'csrc'),
                    '-o', str(exe), str(src)], check=True, capture_output=True)
    out = subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout.split()
    es, ec, ec2, el, ee, ea = (float(v) for v in out[:6])
    assert es < 4e-16 and ec < 4e-16 and ec2 < 4e-16, (es, ec, ec2)     # absolute, |value| <= 1
    assert el < 4e-16 and ee < 4e-16 and ea < 4e-16, (el, ee, ea)       # relative
    assert [float(v) for v in out[6:12]] == [0.0, 1.0, 0.0, 1.0, 0.0, 1.0]
    assert float(out[12]) < 6e-16, out[12]                               # table version: absolute
    # inverse normal CDF (the line-shape deviate of the exact kernels): relative, whole range and the 1e-6 tails
    assert float(out[13]) < 1e-15 and float(out[14]) < 1e-15, out[13:15]
    assert float(out[15]) == 0.0 and abs(float(out[16]) + 8.2923610758135947) < 1e-14 and abs(float(out[17]) - 8.2095361516013874) < 1e-14, out[15:18]


def test_pretest_normal_deviate_coefficients_against_scipy():
    """
    normal_approx (csrc/xrt_trace.cuh): the FP32 inverse-normal-CDF of the Bragg pre-test.  The
    constants are read from the source and the same arithmetic is done in numpy float32; its
    error against scipy must stay far inside the 2e-3 the pre-test's margin (cull_err) allows.
    """
    import re
    import numpy as np
    from scipy.stats import norm
    text = open(os.path.join(ROOT, 'xicsrt_b200', 'csrc', 'xrt_trace.cuh')).read()
    body = text[text.index('float normal_approx('):]
    body = body[:body.index('\n}\n')]
    lead = float(re.search(r'float p = ([-0-9.e+]+)f;', body).group(1))
    coef = [float(c) for c in re.findall(r'p = fmaf\(p, w, ([-0-9.e+]+)f\);', body)]
    assert len(coef) == 4 and 'w < 5.0f' in body and '(hi >> 9)) - 3.0f' in body
    f32 = np.float32
    hi = np.random.default_rng(3).integers(0, 2**32, 2000000, dtype=np.uint64)
    k = (hi >> np.uint64(9)).astype(np.float64)
    x = (k * 2.0**-22 - 1.0).astype(f32)                  # exact in float32: the [2, 4) bit pattern minus 3
    w = (f32(-0.6931471805599453) * np.log2((f32(1) - x * x).astype(f32)).astype(f32)).astype(f32)
    usable = w < f32(5)
    w = (w - f32(2.5)).astype(f32)
    p = np.full_like(w, f32(lead))
    for c in coef:
        p = (p * w + f32(c)).astype(f32)
    z = (f32(1.4142135623730951) * p * x).astype(f32)
    u = (hi.astype(np.float64) + 0.5) * 2.0**-32          # any uniform whose top 32 bits are `hi`
    err = np.abs(z.astype(np.float64) - norm.ppf(u))[usable]
    assert usable.mean() > 0.99 and np.abs(z[usable]).max() < 2.95
    assert err.max() < 1e-4, err.max()                    # 20 x inside the 2e-3 of cull_err
