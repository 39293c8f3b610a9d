"""Dev tool: coefficients of common.cuh:gelu_sig — exact GELU in logistic form x / (1 + exp(-2 x q(x^2))), with q fitted
(weighted Lawson minimax on the GELU output error) to atanh(erf(x / sqrt 2)) / x on |x| <= 5.5."""
import numpy as np
from scipy.special import log_ndtr, ndtr

X, DEG = 5.5, 3
x = np.linspace(1e-4, X, 200001)
phi = ndtr(x)
target = 0.5 * (log_ndtr(x) - log_ndtr(-x)) / x
gelu = x * phi
sens = x * 2 * phi * (1 - phi) * x            # d gelu / d q
A = np.stack([x ** (2 * k) for k in range(DEG)], 1)
w, best = np.ones_like(x), None
for _ in range(200):
    c, *_ = np.linalg.lstsq(A * (w * sens)[:, None], target * w * sens, rcond=None)
    q = A @ c
    e = np.maximum(np.abs(x / (1 + np.exp(-2 * x * q)) - gelu), np.abs(-x / (1 + np.exp(2 * x * q)) + x * ndtr(-x)))
    if best is None or e.max() < best[0]:
        best = (e.max(), c.copy())
    w = w * (1 + 2 * e / e.max())
    w /= w.mean()
print("max abs error", best[0], "coefficients (x^0, x^2, x^4 of q):", list(best[1]))
