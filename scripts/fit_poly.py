"""Fit the polynomial maps of csrc/rod_math.cuh (Chebyshev-node interpolation in 50-digit arithmetic).

    python scripts/fit_poly.py bendw  <w2max> [degree]   theta'/sin(theta') as a function of w2 = |axial(R - R^T)|^2 = 4 sin^2(theta),
                                                          theta' = acos(cos(theta) - 1e-10)  (the reference's guarded angle)
Prints the ascending coefficients and the max relative error on the range.
"""
import sys

import mpmath as mp

mp.mp.dps = 50


def bendw(w2):
    c = mp.sqrt(1 - w2 / 4) - mp.mpf("1e-10")
    th = mp.acos(c)
    return th / mp.sin(th)


def fit(f, lo, hi, deg):
    n = deg + 1
    xs = [(lo + hi) / 2 + (hi - lo) / 2 * mp.cos(mp.pi * (2 * k + 1) / (2 * n)) for k in range(n)]
    ys = [f(x) for x in xs]
    # solve the Vandermonde system in high precision (degree <= 15: fine at 50 digits)
    A = mp.matrix(n, n)
    for i, x in enumerate(xs):
        for j in range(n):
            A[i, j] = x ** j
    c = mp.lu_solve(A, mp.matrix(ys))
    coef = [float(c[i]) for i in range(n)]
    worst = mp.mpf(0)
    for k in range(2001):
        x = lo + (hi - lo) * mp.mpf(k) / 2000
        p = sum(mp.mpf(coef[j]) * x ** j for j in range(n))
        worst = max(worst, abs(p / f(x) - 1))
    return coef, float(worst)


if __name__ == "__main__":
    kind, hi = sys.argv[1], mp.mpf(sys.argv[2])
    f = {"bendw": bendw}[kind]
    degs = [int(sys.argv[3])] if len(sys.argv) > 3 else range(3, 16)
    for d in degs:
        coef, err = fit(f, mp.mpf(0), hi, d)
        print(d, f"{err:.3e}")
        if err < 1.2e-16 or len(sys.argv) > 3:
            print("{" + ", ".join(repr(c) for c in coef) + "}")
            break
