#!/usr/bin/env python
"""Minimax-quality polynomial coefficients for sin(pi h)/h and cos(pi h) on |h| <= 1/2 in T = h^2
(Chebyshev interpolation in 60-digit arithmetic, then rounded to double).  Used by the fp64 sampler."""
import mpmath as mp

mp.mp.dps = 60


def cheb_fit(f, a, b, deg):
    n = deg + 1
    nodes = [(a + b) / 2 + (b - a) / 2 * mp.cos(mp.pi * (2 * k + 1) / (2 * n)) for k in range(n)]
    A = mp.matrix(n, n)
    y = mp.matrix(n, 1)
    for i, x in enumerate(nodes):
        for j in range(n):
            A[i, j] = x ** j
        y[i] = f(x)
    c = mp.lu_solve(A, y)
    return [c[i] for i in range(n)]


def maxerr(f, c, a, b, weight=lambda T: 1, m=4000):
    worst = 0
    for k in range(m + 1):
        T = a + (b - a) * k / m
        p = sum(float(ci) * mp.mpf(T) ** i for i, ci in enumerate(c))   # coefficients rounded to double
        worst = max(worst, abs((p - f(mp.mpf(T))) * weight(mp.mpf(T))))
    return worst


fq = lambda T: mp.pi if T == 0 else mp.sin(mp.pi * mp.sqrt(T)) / mp.sqrt(T)
fc = lambda T: mp.cos(mp.pi * mp.sqrt(T))
for name, f, deg, w in (("Q (sin(pi h)/h)", fq, 8, lambda T: mp.sqrt(T)), ("C (cos(pi h))", fc, 8, lambda T: 1),
                        ("Q deg7", fq, 7, lambda T: mp.sqrt(T)), ("C deg9", fc, 9, lambda T: 1)):
    c = cheb_fit(f, mp.mpf(0), mp.mpf(1) / 4, deg)
    print(name, "deg", deg, "max abs err", mp.nstr(maxerr(f, c, 0, mp.mpf(1) / 4, w), 5))
    print("   ", ", ".join(float(ci).hex() for ci in c))
    print("   ", ", ".join(repr(float(ci)) for ci in c))
