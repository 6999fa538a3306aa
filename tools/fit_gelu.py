"""Fit and check of the single-transcendental forward GELU used by the GEMM epilogue (csrc/common.cuh gelu_erf):
0.5 * erfc(a / sqrt2) ~= 2^-(1 + r(a)), r = degree-5 polynomial without constant term.  Needs scipy (CPU only)."""
import numpy as np
from scipy.optimize import least_squares
from scipy.special import erfc

a = np.linspace(0, 7, 14001)
true_tail = 0.5 * erfc(a / np.sqrt(2))
V = np.stack([a ** k for k in range(1, 6)], 1)
w = erfc(a / np.sqrt(2)) * np.maximum(a, 0.3)
c0 = np.linalg.lstsq(V * w[:, None], -np.log2(erfc(a / np.sqrt(2))) * w, rcond=None)[0]


def resid(c):
    return (0.5 * np.exp2(-(V @ c)) - true_tail) * np.maximum(a, 0.5)


c = least_squares(lambda c: np.sign(resid(c)) * np.abs(resid(c)) ** 4 * 1e12, c0, xtol=1e-15, ftol=1e-15, max_nfev=4000).x
print("coefficients a^1..a^5:", list(c))
x = np.linspace(-12, 12, 2400001).astype(np.float32)
ax = np.abs(x)
p = np.float32(-c[4])
for k in (3, 2, 1, 0):
    p = np.float32(p * ax + np.float32(-c[k]))
t = np.exp2(np.float32(p * ax + np.float32(-1.0)).astype(np.float64)).astype(np.float32)
g = np.maximum(x, 0) - ax * t
ref = x.astype(np.float64) * 0.5 * erfc(-x.astype(np.float64) / np.sqrt(2))
print("max |gelu - exact| in fp32:", float(np.abs(g - ref).max()))

# derivative used by the GELU' epilogue (csrc/common.cuh gelu_erf_grad): Phi(x) + x * phi(x) from the same polynomial
half_minus = np.float32(0.5) - t
cdf = np.float32(0.5) + np.copysign(half_minus, x)
gauss = np.exp2(x.astype(np.float64) ** 2 * (-0.7213475204444817) - 1.3257480647361595).astype(np.float32)
X = x.astype(np.float64)
ref_d = 0.5 * erfc(-X / np.sqrt(2)) + X * np.exp(-X * X / 2) / np.sqrt(2 * np.pi)
print("max |gelu' - exact| in fp32:", float(np.abs(cdf + x * gauss - ref_d).max()))
