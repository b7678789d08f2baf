"""Fit and check the sigmoid-form evaluation of the exact-erf GELU used by the GEMM epilogues (csrc/crf_ptx.cuh).

    Phi(x) = 1 / (1 + 2^(-x R(t))),  t = min(x^2, 25),  R(t) = c0 + c1 t + c2 t^2

`python tools/fit_gelu.py` refits the three coefficients (minimax on gelu over |x| <= 9, Nelder-Mead) and reports the
fp32 error of gelu and gelu' for the coefficients compiled into the library.  Development tool (needs scipy)."""
import numpy as np
from scipy.optimize import minimize
from scipy.special import erf, erfc

COMPILED = np.float32([2.3011212, 0.10677572, -0.0010142630])


def fit(deg=2, xm=5.0):
    xs = np.linspace(1e-4, 9, 60001)
    gp_t = xs * 0.5 * (1 + erf(xs / np.sqrt(2)))
    gn_t = -xs * 0.5 * erfc(xs / np.sqrt(2))

    def cost(c):
        t = np.minimum(xs * xs, xm ** 2)
        r = np.zeros_like(t)
        for k in range(deg, -1, -1):
            r = r * t + c[k]
        w = xs * r
        return max(np.abs(xs / (1 + np.exp2(-w)) - gp_t).max(), np.abs(-xs / (1 + np.exp2(w)) - gn_t).max())

    t = np.linspace(0, xm ** 2, 4000)[1:]
    x = np.sqrt(t)
    w_true = (np.log1p(erf(x / np.sqrt(2))) - np.log(erfc(x / np.sqrt(2)))) / np.log(2)
    c0 = np.polyfit(t, w_true / x, deg)[::-1]
    best = minimize(cost, c0, method="Nelder-Mead", options={"xatol": 1e-12, "fatol": 1e-12, "maxiter": 40000})
    return best.x, best.fun


def check(c):
    c = np.float32(c)
    x = np.linspace(-10, 10, 2000001).astype(np.float32)
    t = np.minimum(x * x, np.float32(25))
    r = (c[2] * t + c[1]) * t + c[0]
    cdf = (np.float32(1) / (np.float32(1) + np.exp2(-(x * r)).astype(np.float32))).astype(np.float32)
    xd = x.astype(np.float64)
    g_true = xd * 0.5 * (1 + erf(xd / np.sqrt(2)))
    dg_true = 0.5 * (1 + erf(xd / np.sqrt(2))) + xd * np.exp(-xd * xd / 2) / np.sqrt(2 * np.pi)
    ln2 = np.float32(np.log(2.0))
    wp = (np.float32(5) * c[2] * t + np.float32(3) * c[1]) * t + c[0]          # w'(x) of the fitted exponent
    dg = cdf + x * cdf * (np.float32(1) - cdf) * wp * ln2                     # exact derivative of the sigmoid form
    return float(np.abs(x * cdf - g_true).max()), float(np.abs(dg - dg_true).max())


if __name__ == "__main__":
    c, f = fit()
    print("refit coefficients", [float(np.float32(v)) for v in c], "minimax |gelu err|", f)
    print("compiled coefficients: max |gelu err| %.3g, max |gelu' err| %.3g (fp32 evaluation)" % check(COMPILED))
