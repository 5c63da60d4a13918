"""Fits behind common.cuh's GELU epilogues (exact erf GELU is the target in every case).
  tanh form : Phi(x) = 0.5 + 0.5*tanh(x*Q(x^2))            one MUFU.TANH (2^-11 relative) -- measured, not shipped
  exp2 form : 1 - erf(a/sqrt2) = 2^(-a*R(a)), a = |x|        one MUFU.EX2, h = fma(-|x/2|, 2^(-a R), max(x, 0))
"""
import numpy as np
from scipy.special import erf, erfc
from scipy.optimize import least_squares, minimize

x = np.linspace(-8, 8, 400001)
ref = 0.5 * x * (1 + erf(x / np.sqrt(2)))


def horner(c, t):  # c[0] + c[1] t + ...
    r = np.zeros_like(t) + c[-1]
    for k in c[-2::-1]:
        r = r * t + k
    return r


def g_tanh(a, x):
    return 0.5 * x * (1 + np.tanh(x * horner(a, np.minimum(x * x, 36.0))))


def g_exp2(c, x, amax):
    a = np.minimum(np.abs(x), amax)
    e = np.exp2(-a * horner(c, a))
    return np.maximum(x, 0) - np.abs(0.5 * x) * e


def refine(f, a):
    for p in (4, 8, 16, 32):
        a = minimize(lambda a: np.sum(((f(a) - ref) * 1e4) ** p), a, method="Nelder-Mead",
                     options=dict(xatol=1e-13, fatol=1e-13, maxiter=40000)).x
    return a


if __name__ == "__main__":
    a_std = [0.7978845608, 0.7978845608 * 0.044715]
    print("tanh std", np.abs(g_tanh(a_std, x) - ref).max())
    a = least_squares(lambda a: g_tanh(a, x) - ref, np.array(a_std + [0.0]), xtol=1e-15, ftol=1e-15).x
    a = refine(lambda a: g_tanh(a, x), a)
    print("tanh deg2", a.tolist(), np.abs(g_tanh(a, x) - ref).max())
    amax = 6.0
    aa = np.linspace(1e-3, amax, 4000)
    target = -np.log2(erfc(aa / np.sqrt(2))) / aa
    for deg in (3, 4, 5, 6):
        c = np.polynomial.polynomial.polyfit(aa, target, deg)
        c = least_squares(lambda c: (g_exp2(c, x, amax) - ref) * 1e4, c, xtol=1e-15, ftol=1e-15).x
        c = refine(lambda c: g_exp2(c, x, amax), c)
        c32 = c.astype(np.float32).astype(np.float64)
        print("exp2 deg", deg, [float(np.float32(v)) for v in c], "max err", np.abs(g_exp2(c, x, amax) - ref).max(),
              "f32 coeffs", np.abs(g_exp2(c32, x, amax) - ref).max())
