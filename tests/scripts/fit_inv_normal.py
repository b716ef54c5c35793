"""Coefficients of inv_normal_cdf (xicsrt_b200/csrc/xrt_fastmath.cuh): Chebyshev interpolation of erfinv(x) / x in 60-digit
arithmetic, converted to powers of the centred variable.  usage: python tests/scripts/fit_inv_normal.py 24 20 19 -> /tmp/ninv/coef.json"""
import mpmath as mp, numpy as np, sys, json
mp.mp.dps = 60
def g_of_w(w):
    w = mp.mpf(w)
    if w == 0: return mp.sqrt(mp.pi)/2
    x = mp.sqrt(1 - mp.exp(-w))
    return mp.erfinv(x)/x
def cheb_fit(f, a, b, deg):
    n = deg + 1
    nodes = [mp.cos(mp.pi*(k+mp.mpf(1)/2)/n) for k in range(n)]
    fx = [f((a+b)/2 + (b-a)/2*t) for t in nodes]
    c = []
    for j in range(n):
        s = mp.fsum(fx[k]*mp.cos(mp.pi*j*(k+mp.mpf(1)/2)/n) for k in range(n))
        c.append(2*s/n)
    c[0] /= 2
    return c
def cheb_to_mono(c):
    # sum c_j T_j(t) -> sum m_i t^i
    n = len(c)
    T0 = [mp.mpf(1)]; T1 = [mp.mpf(0), mp.mpf(1)]
    mono = [mp.mpf(0)]*n
    def add(coefs, scale):
        for i,v in enumerate(coefs): mono[i] += scale*v
    add(T0, c[0])
    if n > 1: add(T1, c[1])
    for j in range(2, n):
        T2 = [mp.mpf(0)] + [2*v for v in T1]
        for i,v in enumerate(T0): T2[i] -= v
        add(T2, c[j]); T0, T1 = T1, T2
    return mono
def fit_interval(f, a, b, deg, var_center, var_half):
    # polynomial in tau = (v - center) where v in [a,b]; t = (v-center)/half
    c = cheb_fit(f, mp.mpf(a), mp.mpf(b), deg)
    mono_t = cheb_to_mono(c)
    half = (mp.mpf(b)-mp.mpf(a))/2
    return [m/half**i for i,m in enumerate(mono_t)], [abs(v) for v in c[-3:]]
res = {}
for deg in (int(sys.argv[1]),):
    co, tailc = fit_interval(g_of_w, 0, 6.25, deg, 3.125, 3.125)
    print('central deg', deg, 'last cheb', [mp.nstr(v, 3) for v in tailc])
    res['central'] = [float(v) for v in co]
def h_of_s(s):
    w = mp.mpf(s)**2
    x = mp.sqrt(1 - mp.exp(-w))
    return mp.erfinv(x)/x
d2, d3 = int(sys.argv[2]), int(sys.argv[3])
co, tailc = fit_interval(h_of_s, 2.5, 4.0, d2, 3.25, 0.75); print('tail1 deg', d2, [mp.nstr(v,3) for v in tailc]); res['tail1'] = [float(v) for v in co]
co, tailc = fit_interval(h_of_s, 4.0, 6.02, d3, 5.01, 1.01); print('tail2 deg', d3, [mp.nstr(v,3) for v in tailc]); res['tail2'] = [float(v) for v in co]
import os
os.makedirs('/tmp/ninv', exist_ok=True)
json.dump(res, open('/tmp/ninv/coef.json','w'))
for k, v in res.items():
    print(k, len(v), ', '.join(repr(c) for c in v))
