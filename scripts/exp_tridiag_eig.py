"""CPU prototype of the planned full-rank KLT: Householder tridiagonalisation (FP64), bisection,
one twisted-factorisation solve per eigenvalue (no re-orthogonalisation), back-transformation,
then the SIIB^Gauss quadratic forms -- compared with numpy.linalg.eigh on the same matrices."""
import sys
import numpy as np
sys.path.insert(0, "."); sys.path.insert(0, "scripts")
from exp_jacobi_sweeps import sxx_of


def tridiag(A):
    A = A.copy(); n = A.shape[0]
    V = np.zeros((n, n)); tau = np.zeros(n)
    for k in range(n - 2):
        x = A[k + 1:, k].copy()
        alpha = x[0]; sig = np.dot(x[1:], x[1:])
        if sig == 0.0:
            continue
        nrm = np.sqrt(alpha * alpha + sig)
        beta = -np.copysign(nrm, alpha)
        v = x.copy(); v[0] = alpha - beta; t = (beta - alpha) / beta; v /= v[0]
        V[k + 1:, k] = v; tau[k] = t
        S = A[k + 1:, k + 1:]
        p = t * (S @ v)
        w = p - (0.5 * t * np.dot(p, v)) * v
        S -= np.outer(v, w) + np.outer(w, v)
        A[k + 1, k] = A[k, k + 1] = beta
        A[k + 2:, k] = 0; A[k, k + 2:] = 0
    return np.diag(A).copy(), np.diag(A, -1).copy(), V, tau


def bisect_all(d, e, iters=64):
    n = len(d)
    r = np.abs(np.concatenate(([0], e))) + np.abs(np.concatenate((e, [0])))
    lo0, hi0 = (d - r).min(), (d + r).max()
    e2 = e * e
    def count(x):   # eigenvalues < x (vectorised over x)
        q = d[0] - x; c = (q < 0).astype(int)
        for i in range(1, n):
            q = d[i] - x - e2[i - 1] / np.where(q == 0, 1e-300, q)
            c += q < 0
        return c
    lo = np.full(n, lo0); hi = np.full(n, hi0); idx = np.arange(n)
    for _ in range(iters):
        mid = 0.5 * (lo + hi)
        c = count(mid)
        gt = c > idx          # eigenvalue idx is below mid
        hi = np.where(gt, mid, hi); lo = np.where(gt, lo, mid)
    return 0.5 * (lo + hi)


def getvec(d, e, lam):
    """One twisted factorisation solve (dlar1v style) for T - lam I."""
    n = len(d)
    s = np.zeros(n); p = np.zeros(n); Lp = np.zeros(n - 1); Um = np.zeros(n - 1)
    # forward: D+ from the top
    dp = np.zeros(n); dp[0] = d[0] - lam
    for i in range(n - 1):
        if dp[i] == 0: dp[i] = 1e-300
        Lp[i] = e[i] / dp[i]
        dp[i + 1] = (d[i + 1] - lam) - Lp[i] * e[i]
    # backward: D- from the bottom
    dm = np.zeros(n); dm[n - 1] = d[n - 1] - lam
    for i in range(n - 2, -1, -1):
        if dm[i + 1] == 0: dm[i + 1] = 1e-300
        Um[i] = e[i] / dm[i + 1]
        dm[i] = (d[i] - lam) - Um[i] * e[i]
    gamma = dp + dm - (d - lam)
    k = int(np.argmin(np.abs(gamma)))
    z = np.zeros(n); z[k] = 1.0
    for i in range(k - 1, -1, -1):
        z[i] = -Lp[i] * z[i + 1]
    for i in range(k, n - 1):
        z[i + 1] = -Um[i] * z[i]
    return z / np.linalg.norm(z)


def back(V, tau, Z):
    U = Z.copy(); n = V.shape[0]
    for k in range(n - 3, -1, -1):
        v = V[:, k]
        if tau[k] == 0: continue
        U -= np.outer(tau[k] * v, v @ U)
    return U


def siib(lam, U, Sxy, Syy):
    a = np.einsum('ij,ik,kj->j', U, Sxy, U); c = np.einsum('ij,ik,kj->j', U, Syy, U)
    ok = (lam > 1e-10 * lam.max()) & (c > 0)
    rho = np.where(ok, a / np.sqrt(np.where(ok, lam * c, 1)), 0.0)
    rho = np.clip(rho, -1, 1)
    return 80 / 15 * np.sum(-0.5 * np.log2(1 - (0.75 * rho) ** 2))


if __name__ == "__main__":
    for i, L in ((0, 52345), (0, 47999), (3, 40111), (5, 70003)):
        A, Xs, Ys = sxx_of(i, L)
        Xc = Xs - Xs.mean(1, keepdims=True); Yc = Ys - Ys.mean(1, keepdims=True)
        Sxy = Xc @ Yc.T; Syy = Yc @ Yc.T
        lam0, U0 = np.linalg.eigh(A)
        d, e, V, tau = tridiag(A)
        lam = bisect_all(d, e)
        Z = np.stack([getvec(d, e, l) for l in lam], axis=1)
        U = back(V, tau, Z)
        orth = np.abs(U.T @ U - np.eye(len(lam))).max()
        res = np.abs(A @ U - U * lam).max() / lam0.max()
        s0, s1 = siib(lam0, U0, Sxy, Syy), siib(lam, U, Sxy, Syy)
        print("pair %d L=%d: cond %.1e  eig rel err %.1e  orth %.1e  resid %.1e  SIIB eigh %.6f  tridiag %.6f  rel %.1e"
              % (i, L, lam0[-1] / lam0[0], np.max(np.abs(lam - lam0) / lam0), orth, res, s0, s1, abs(s1 - s0) / s0))
