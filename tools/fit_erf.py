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


# ------------------------------------------------------------------------------------------------ packed-fp16 gelu' (--grad)
def fit_grad():
    """gelu_erf_grad_bf16x2 in csrc/ard_common.cuh:  gelu'(x) = 1/2 + t/2 + (1 - t^2) x (d0 + d1 x^2) / 2,  t = tanh(x (c0 + c1 x^2)),
    the four constants fitted (iteratively re-weighted least squares -> near-minimax) to the exact Phi(x) + x phi(x)."""
    from scipy.optimize import least_squares
    x = np.linspace(-8, 8, 400001)
    exact = 0.5 * (1 + erf(x / np.sqrt(2))) + x * np.exp(-x * x / 2) / np.sqrt(2 * np.pi)

    def model(c):
        x2 = x * x
        t = np.tanh(x * (c[0] + c[1] * x2))
        return 0.5 + 0.5 * t + 0.5 * (1 - t * t) * x * (c[2] + c[3] * x2)
    c = np.array([0.80015708, 0.03470089, 0.80015708, 3 * 0.03470089])
    w = np.ones_like(x)
    for _ in range(30):
        c = least_squares(lambda cc: (model(cc) - exact) * w, c, xtol=1e-15, ftol=1e-15).x
        e = np.abs(model(c) - exact)
        w = w * (1 + 2 * e / e.max())
        w /= w.mean()
    return c, np.abs(model(c) - exact).max()


def emulate_grad_f16(c, n=2_000_000, sigma=1.5):
    import torch
    h = np.float16
    xs = np.random.default_rng(0).normal(0, sigma, n)
    xb = torch.tensor(xs, dtype=torch.float32).bfloat16().double().numpy()   # hpre is stored as bf16
    ex = lambda v: 0.5 * (1 + erf(v / np.sqrt(2))) + v * np.exp(-v * v / 2) / np.sqrt(2 * np.pi)   # noqa: E731
    x = xb.astype(h)
    x2 = np.minimum((x * x).astype(h), h(64))
    u = (x * (h(c[1]) * x2 + h(c[0])).astype(h)).astype(h)
    t = np.tanh(u.astype(np.float64)).astype(h)
    wv = (x * (h(0.5 * c[3]) * x2 + h(0.5 * c[2])).astype(h)).astype(h)
    s = (h(1) - (t * t).astype(h)).astype(h)
    r = (t * h(0.5) + h(0.5)).astype(h)
    g = ((s * wv).astype(h) + r).astype(h).astype(np.float64)
    return (np.linalg.norm(g - ex(xb)) / np.linalg.norm(ex(xb)), np.abs(g - ex(xb)).max(),
            np.linalg.norm(ex(xb) - ex(xs)) / np.linalg.norm(ex(xs)))


if __name__ == "__main__" and "--grad" in __import__("sys").argv:
    c, e = fit_grad()
    print("gelu' tanh-derivative form: c0 c1 d0 d1 =", c, "max |error| %.2e" % e)
    print("fp16 evaluation on bf16 inputs: rel l2 %.3e, max abs %.2e; bf16 rounding of the input alone: rel l2 %.3e" % emulate_grad_f16(c))
