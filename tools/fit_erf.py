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


# ------------------------------------------------------------------------------------------------ packed-fp16 GELU (--tanh)
def fit_tanh(deg=2, clampv=16.0):
    """gelu_erf_f16x2 in csrc/ard_common.cuh: erf(x / sqrt 2) = tanh(x P(min(x^2, clampv))), P of degree `deg` in x^2,
    least squares on the absolute GELU error 0.5 x (model - erf)."""
    from scipy.optimize import least_squares
    x = np.linspace(0, 6, 60001)
    tgt = erf(x / np.sqrt(2))

    def model(c):
        x2 = np.minimum(x * x, clampv)
        p = np.zeros_like(x) + c[-1]
        for ck in c[-2::-1]:
            p = p * x2 + ck
        return np.tanh(x * p)
    c0 = np.zeros(deg + 1)
    c0[0] = np.sqrt(2 / np.pi)
    r = least_squares(lambda c: (model(c) - tgt) * (0.5 * np.maximum(x, 0.3)), c0, xtol=1e-15, ftol=1e-15)
    e = np.abs(model(r.x) - tgt)
    return r.x, e.max(), np.abs(0.5 * x * e).max()


def emulate_f16(c, mufu_noise=2.0 ** -11, n=2_000_000, seed=0):
    """Rel. l2 error of the fp16 evaluation (every op rounded to half) against the float64 erf GELU on N(0, 1.5) inputs;
    `mufu_noise` is an assumed absolute error of MUFU.TANH.F16."""
    def r(v):
        return np.asarray(v, dtype=np.float64).astype(np.float16).astype(np.float64)
    rng = np.random.default_rng(seed)
    x32 = (rng.standard_normal(n) * 1.5).astype(np.float32)
    gt = 0.5 * x32.astype(np.float64) * (1 + erf(x32.astype(np.float64) / np.sqrt(2)))
    x = r(x32)
    x2 = r(np.minimum(r(x * x), 16.0))
    p = r(r(c[2]) * x2 + r(c[1]))
    p = r(p * x2 + r(c[0]))
    t = r(np.tanh(r(x * p)) + mufu_noise * rng.uniform(-1, 1, x.shape))
    hx = r(0.5 * x)
    g = r(hx * t + hx)
    return np.linalg.norm(g - gt) / np.linalg.norm(gt)


if __name__ == "__main__" and "--tanh" in __import__("sys").argv:
    c, e_erf, e_gelu = fit_tanh()
    print("tanh form: coeffs", c, "max|erf err| %.2e max|gelu err| %.2e" % (e_erf, e_gelu))
    print("fp16 evaluation rel l2 error: %.3e (exact tanh), %.3e (MUFU noise 2^-11)" % (emulate_f16(c, 0.0), emulate_f16(c)))
