"""Fit of the GELU used by the GEMM / fused-MLP epilogues (svb_convnext_kernels.cuh: gelu_pack2, gelu_fast).

    GELU(x) = relu(x) - |x| * E(|x|),   E(u) = erfc(u / sqrt2) / 2 = 2 ** (u * p(u) - 1)

p is a polynomial fitted by iteratively re-weighted least squares so that the max ABSOLUTE error of GELU is minimal.
Degree 3 (what gelu_pack2 uses, evaluated in v = -u): 8.6e-6; degree 4 (gelu_fast): 5.7e-7.  Run: python scripts/fit_gelu.py
"""
import numpy as np
from scipy.special import erfc

u = np.linspace(1e-6, 7.5, 40000)
target = np.log2(erfc(u / np.sqrt(2)) / 2) + 1  # = u * p(u)


def fit(deg):
    w = u * erfc(u / np.sqrt(2)) / 2 * np.log(2) * u
    A = np.vander(u, deg + 1, increasing=True)
    y = target / u
    ww = w.copy()
    for _ in range(60):
        c = np.linalg.lstsq(A * ww[:, None], y * ww, rcond=None)[0]
        err = (A @ c - y) * w
        ww = ww * (1 + 4 * np.abs(err) / np.abs(err).max())
        ww /= ww.max()
    q = (A.astype(np.float32) @ c.astype(np.float32)) * u.astype(np.float32) - 1
    E = np.exp2(q.astype(np.float64))
    return c, np.abs(u * E - u * erfc(u / np.sqrt(2)) / 2).max()


if __name__ == "__main__":
    for d in (2, 3, 4):
        c, e = fit(d)
        print(f"degree {d}: max |GELU error| {e:.3e}; p(u) coefficients (ascending) {c}; in v = -u: {[(-1) ** (k + 1) * ck for k, ck in enumerate(c)]}")
