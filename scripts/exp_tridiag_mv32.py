"""CPU experiment behind NELE_TRIDIAG_MV32: Householder tridiagonalisation whose matrix-vector pass reads an
FP32-rounded copy of the trailing matrix (FP64 accumulation, FP64 rank-2 updates on the master copy), compared
with the all-FP64 prototype (scripts/exp_tridiag_eig.py) and numpy.linalg.eigh through the SIIB^Gauss score.
Measured (2 pairs): SIIB relative deviation 6e-8 .. 6e-7, eigen-residual 4e-10 of lambda_max."""
import sys

import numpy as np

sys.path.insert(0, ".")
sys.path.insert(0, "scripts")
import exp_tridiag_eig as E  # noqa: E402
from exp_jacobi_sweeps import sxx_of  # noqa: E402


def tridiag32(A):
    A = A.copy()
    n = A.shape[0]
    V = np.zeros((n, n))
    tau = np.zeros(n)
    for k in range(n - 2):
        x = A[k + 1:, k].copy()
        alpha = x[0]
        sig = np.dot(x[1:], x[1:])
        if sig == 0.0:
            continue
        nrm = np.sqrt(alpha * alpha + sig)
        beta = -np.copysign(nrm, alpha)
        v = x.copy()
        v[0] = alpha - beta
        t = (beta - alpha) / beta
        v /= v[0]
        V[k + 1:, k] = v
        tau[k] = t
        S = A[k + 1:, k + 1:]
        p = t * (S.astype(np.float32).astype(np.float64) @ v)   # the matvec sees an FP32 copy
        w = p - (0.5 * t * np.dot(p, v)) * v
        S -= np.outer(v, w) + np.outer(w, v)
        A[k + 1, k] = A[k, k + 1] = beta
        A[k + 2:, k] = 0
        A[k, k + 2:] = 0
    return np.diag(A).copy(), np.diag(A, -1).copy(), V, tau


if __name__ == "__main__":
    for i, L in ((0, 52345), (3, 40111)):
        A, Xs, Ys = sxx_of(i, L)
        Xc = Xs - Xs.mean(1, keepdims=True)
        Yc = Ys - Ys.mean(1, keepdims=True)
        Sxy, Syy = Xc @ Yc.T, Yc @ Yc.T
        lam0, U0 = np.linalg.eigh(A)
        s0 = E.siib(lam0, U0, Sxy, Syy)
        for name, fn in (("fp64", E.tridiag), ("fp32-matvec", tridiag32)):
            d, e, V, tau = fn(A)
            lam = E.bisect_all(d, e)
            Z = np.stack([E.getvec(d, e, l) for l in lam], axis=1)
            U = E.back(V, tau, Z)
            s1 = E.siib(lam, U, Sxy, Syy)
            res = np.abs(A @ U - U * lam).max() / lam0.max()
            print("pair %d L=%d %-12s SIIB eigh %.6f  got %.6f  rel %.2e  resid %.1e" % (i, L, name, s0, s1, abs(s1 - s0) / s0, res))
