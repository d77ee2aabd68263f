"""Fit the polynomial used by erf_poly() in csrc/ard_common.cuh.

erf(x) = sign(x) * (1 - 2^(-|x| * Q(|x|)))   with Q a degree-N polynomial fitted to  -log2(erfc(t)) / t  on (0, T].
One MUFU (ex2) per evaluation; absolute error is reported for float32 evaluation.
"""
import numpy as np
from scipy.special import erfc, erf

def fit(N=7, T=4.2):
    k = np.arange(4000)
    t = 0.5 * T * (1 - np.cos(np.pi * (k + 0.5) / 4000))   # Chebyshev nodes on (0,T)
    t = t[t > 1e-6]
    y = -np.log2(erfc(t)) / t
    # weight so that the ABSOLUTE error of erf is equalised: d erf = erfc * ln2 * t * dQ
    w = erfc(t) * t
    V = np.vander(t, N + 1, increasing=True)
    c, *_ = np.linalg.lstsq(V * w[:, None], y * w, rcond=None)
    return c

def check(c, T=4.2):
    x = np.linspace(-6, 6, 2000001).astype(np.float32)
    ax = np.minimum(np.abs(x), np.float32(T)).astype(np.float32)
    p = np.float32(c[-1]) * np.ones_like(ax)
    for ck in c[-2::-1]:
        p = (p * ax + np.float32(ck)).astype(np.float32)
    r = (np.float32(1) - np.exp2(-(p * ax).astype(np.float32)).astype(np.float32)).astype(np.float32)
    r = np.copysign(r, x)
    err = np.abs(r.astype(np.float64) - erf(x.astype(np.float64)))
    g = 0.5 * x.astype(np.float64) * (1 + r.astype(np.float64))
    gt = 0.5 * x.astype(np.float64) * (1 + erf(x.astype(np.float64)))
    return err.max(), np.abs(g - gt).max()

if __name__ == "__main__":
    for N in (5, 6, 7, 8):
        for T in (3.9, 4.2):
            c = fit(N, T)
            print(N, T, "max|erf err| %.2e  max|gelu err| %.2e" % check(c, T))
    c = fit(5, 3.9)
    print("coeffs (c0..cN):")
    for i, ck in enumerate(c):
        print(f"#define ERF_Q{i} {ck:.9e}f")
