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
    double es = 0, ec = 0, ec2 = 0, el = 0, ee = 0, ea = 0;
    for (int i = 0; i < 2000000; ++i) {
        double u = U(g);
        if (i < 9) u = i * 0.125;                       // quadrant boundaries
        double s, c;
        sincos_2pi(u, s, c);
        long double rs = sinl(2 * PI * (long double)u), rc = cosl(2 * PI * (long double)u);
        es = fmax(es, fabs((double)(s - rs)));
        ec = fmax(ec, fabs((double)(c - rc)));
        ec2 = fmax(ec2, fabs((double)(cos_2pi(u) - rc)));
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
    double s0, c0, s1, c1;
    sincos_2pi(0.0, s0, c0);
    sincos_2pi(1.0, s1, c1);
    printf("%.3e %.3e %.3e %.3e %.3e %.3e %g %g %g %g %g %g\n", es, ec, ec2, el, ee, ea,
           s0, c0, s1, c1, log_pos(1.0), exp_neg(0.0));
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
    assert [float(v) for v in out[6:]] == [0.0, 1.0, 0.0, 1.0, 0.0, 1.0]
