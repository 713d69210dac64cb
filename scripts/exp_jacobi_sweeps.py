import sys, numpy as np, time
sys.path.insert(0, ".")
from oracle import pysiib_np, intel_np
from nele_gan_b200.synth import make_pair

def sxx_of(i, L):
    x, y, _ = make_pair(i, L)
    x = x.astype(np.float64); y = y.astype(np.float64)
    M, _ = intel_np.siib_tiling_factor(x, 16000)
    xt, yt = np.tile(x, M), np.tile(y, M)
    Xs, Ys, R = pysiib_np.siib_features(xt, yt)
    Xc = Xs - Xs.mean(axis=1, keepdims=True)
    return Xc @ Xc.T, Xs, Ys

def pivchol(A, tol=1e-10):
    A = A.copy(); n = A.shape[0]
    L = np.zeros((n, n)); perm = []
    d = np.diag(A).copy(); done = np.zeros(n, bool)
    t = None
    for k in range(n):
        dd = np.where(done, -np.inf, d); p = int(np.argmax(dd))
        if t is None: t = dd[p] * tol
        if not dd[p] > t: return L[:, :k]
        lkk = np.sqrt(dd[p])
        col = (A[:, p] - L[:, :k] @ L[p, :k]) / lkk
        col[done] = 0.0; col[p] = lkk
        L[:, k] = col; done[p] = True
        d = d - col * col
    return L

def jacobi_onesided(G, tol=1.5e-6, maxsweeps=30, dtype=np.float32):
    """cyclic round-robin one-sided Jacobi on columns of G; returns sweeps, rotations per sweep"""
    G = G.astype(dtype).copy(); n = G.shape[1]
    if n % 2: G = np.concatenate([G, np.zeros((G.shape[0], 1), dtype)], 1); n += 1
    idx = np.arange(n)
    rots = []
    for sw in range(maxsweeps):
        nrot = 0
        order = idx.copy()
        for rnd in range(n - 1):
            p = order[: n // 2]; q = order[n // 2:][::-1]
            P = G[:, p]; Q = G[:, q]
            a = (P * P).sum(0); b = (Q * Q).sum(0); g = (P * Q).sum(0)
            rot = (g * g > tol * tol * a * b) & (a > 0) & (b > 0)
            with np.errstate(all='ignore'):
                zeta = (b - a) / (2 * g)
                t = np.sign(zeta) / (np.abs(zeta) + np.sqrt(1 + zeta * zeta))
            t = np.where(rot, t, 0).astype(dtype)
            c = (1 / np.sqrt(1 + t * t)).astype(dtype); s = c * t
            G[:, p] = c * P - s * Q
            G[:, q] = s * P + c * Q
            nrot += int(rot.sum())
            order = np.concatenate([[order[0]], [order[-1]], order[1:-1]])
        rots.append(nrot)
        if nrot == 0: break
    return G, rots

if __name__ == "__main__":
    L = int(sys.argv[1]) if len(sys.argv) > 1 else 47999
    A, Xs, Ys = sxx_of(0, L)
    lam = np.linalg.eigvalsh(A)
    print("Nf", Xs.shape[1], "lam max/min", lam[-1], lam[0], "cond %.3g" % (lam[-1] / lam[0]))
    print("lam quantiles", np.quantile(lam / lam[-1], [0, .1, .25, .5, .75, .9, 1]))
    Lc = pivchol(A)
    print("rank", Lc.shape[1])
    t = time.time(); G, rots = jacobi_onesided(Lc); print("L cols: sweeps", len(rots), rots, "%.1fs" % (time.time() - t))
    # direct on A columns? (no preconditioning)
    # one extra LR step
    A1 = Lc.T @ Lc
    L1 = pivchol(A1)
    G1, rots1 = jacobi_onesided(L1); print("LR1: sweeps", len(rots1), rots1)
    A2 = L1.T @ L1
    L2 = pivchol(A2)
    G2, rots2 = jacobi_onesided(L2); print("LR2: sweeps", len(rots2), rots2)
    # eigenvalue check
    ev = np.sort((G.astype(np.float64) ** 2).sum(0))[-420:]
    print("eig rel err (L)", np.max(np.abs(ev - lam) / lam))
    ev2 = np.sort((G2.astype(np.float64) ** 2).sum(0))[-420:]
    print("eig rel err (LR2)", np.max(np.abs(ev2 - lam) / lam))
