import sys, numpy as np
sys.path.insert(0, "."); sys.path.insert(0, "scripts")
from exp_jacobi_sweeps import sxx_of
for L in (47999, 52345):
    A, Xs, Ys = sxx_of(0, L)
    Xc = Xs - Xs.mean(1, keepdims=True); Yc = Ys - Ys.mean(1, keepdims=True)
    Sxy = Xc @ Yc.T; Syy = Yc @ Yc.T
    lam, U = np.linalg.eigh(A)
    def score(Sxy, Syy):
        a = np.einsum('ij,ik,kj->j', U, Sxy, U); c = np.einsum('ij,ik,kj->j', U, Syy, U)
        rho = a / np.sqrt(np.maximum(lam, 1e-300) * c)
        rho = np.clip(rho, -1, 1)
        return 80 / 15 * np.sum(-0.5 * np.log2(1 - (0.75 * rho) ** 2)), rho
    s0, rho0 = score(Sxy, Syy)
    print(L, "Nf", Xs.shape[1], "cond %.2g" % (lam[-1] / lam[0]), "score", s0)
    dxy = np.sqrt(np.outer(np.diag(A), np.diag(Syy))); dyy = np.sqrt(np.outer(np.diag(Syy), np.diag(Syy)))
    rng = np.random.default_rng(0)
    for sig in (6e-8, 1e-6, 3e-6, 1e-5):
        r = []
        for trial in range(3):
            E1 = rng.standard_normal(Sxy.shape) * sig * dxy
            E2 = rng.standard_normal(Syy.shape) * sig * dyy; E2 = (E2 + E2.T) / 2
            s, rho = score(Sxy + E1, Syy + E2)
            r.append((s - s0) / s0)
        print("  sigma %.0e: rel dscore" % sig, ["%.2e" % v for v in r])
    # contribution of small-lambda components
    I = -0.5 * np.log2(1 - (0.75 * rho0) ** 2)
    order = np.argsort(lam)
    cs = np.cumsum(I[order]) / I.sum()
    for q in (1e-9, 1e-8, 1e-7, 1e-6, 1e-4):
        k = np.searchsorted(lam[order] / lam[-1], q)
        print("  lam/lmax < %.0e: %d comps, info share %.4f" % (q, k, cs[k - 1] if k else 0))
