"""CPU experiments behind the FP32 KLT of round 2 (csrc/siib_klt.cu), all through the SIIB^Gauss score against
numpy.linalg.eigh in FP64 on the same matrices:

  precision   Householder tridiagonalisation with the trailing matrix stored / the arithmetic done in FP64 or FP32
              (s64a64, s32a64, s32a32) and LAPACK's own FP32 eigh, on synthetic pairs (condition numbers 2e5 .. 1e9)
              and on the real speech of tests/golden/haspi_ref.npz.  Measured: s32a32 3e-6 .. 2e-5 on the synthetic
              pairs, up to 3.1e-4 on the toy corpus (SIIB 7.5 .. 95); eigh in FP32 4e-7 .. 1.5e-5.  Tolerance 5e-3.
  iterations  bisection steps per eigenvalue of the (FP32-made) tridiagonal.  Measured: SIIB deviation unchanged
              (2.0e-5, 4.0e-6, 3.8e-6) from 58 down to 38 steps although the orthogonality of the computed basis
              degrades from 2e-7 to 0.6: the vectors that lose orthogonality belong to clusters of tiny eigenvalues,
              whose components carry no information.  The kernel uses 46 (orthogonality <= 1e-3).

  blockwy     (groundwork for the next back-transformation kernel) the reflectors applied as
              compact-WY blocks I - V T V^T of 4 / 16 / 32 reflectors, every product and the T factors in FP32, against
              the reflector-by-reflector application in FP32 and in FP64: does the GEMM form (V^T Z, T W, Z -= V W) cost
              accuracy?  Measured: no -- on the ten cases the SIIB deviation from eigh in FP64 is the one of the FP32
              tridiagonalisation itself (6e-7 .. 3.1e-4) for every block size and equal to the sequential
              application to two digits; orthogonality of the back-transformed basis 6e-7 .. 2e-6.

usage: exp_klt_fp32.py [precision|iterations|blockwy]"""
import sys

import numpy as np
from scipy.linalg import eigh_tridiagonal

sys.path.insert(0, ".")
sys.path.insert(0, "scripts")
import exp_tridiag_eig as E  # noqa: E402
from exp_jacobi_sweeps import sxx_of  # noqa: E402


def tridiag_var(A, store=np.float32, arith=np.float64):
    """Householder tridiagonalisation; trailing matrix stored in `store`, arithmetic in `arith`."""
    A = A.astype(store).copy()
    n = A.shape[0]
    V = np.zeros((n, n))
    tau = np.zeros(n)
    d = np.zeros(n)
    e = np.zeros(n - 1)
    for k in range(n - 2):
        d[k] = A[k, k]
        x = A[k + 1:, k].astype(arith)
        alpha = x[0]
        sig = np.dot(x[1:], x[1:])
        if sig == 0.0:
            e[k] = alpha
            continue
        nrm = np.sqrt(alpha * alpha + sig)
        beta = -np.copysign(nrm, alpha)
        v = x.copy()
        v[0] = alpha - beta
        t = (beta - alpha) / beta
        v = v / v[0]
        V[k + 1:, k] = v
        tau[k] = t
        S = A[k + 1:, k + 1:].astype(arith)
        p = t * (S @ v)
        w = p - (arith(0.5) * t * np.dot(p, v)) * v
        A[k + 1:, k + 1:] = (S - (np.outer(v, w) + np.outer(w, v))).astype(store)
        e[k] = beta
    d[n - 2], d[n - 1], e[n - 2] = A[n - 2, n - 2], A[n - 1, n - 1], A[n - 1, n - 2]
    return d, e, V, tau


def matrices(x, y):
    from oracle import intel_np, pysiib_np
    M, _ = intel_np.siib_tiling_factor(x, 16000)
    Xs, Ys, _ = pysiib_np.siib_features(np.tile(x.astype(np.float64), M), np.tile(y.astype(np.float64), M))
    Xc = Xs - Xs.mean(1, keepdims=True)
    Yc = Ys - Ys.mean(1, keepdims=True)
    return Xc @ Xc.T, Xc @ Yc.T, Yc @ Yc.T


def cases():
    from nele_gan_b200.synth import make_pair
    for i, L in ((0, 52345), (1, 47999), (2, 33536), (3, 40111), (5, 70003), (7, 112345)):
        x, y, _ = make_pair(i, L)
        yield "synth %d L=%d" % (i, L), x, y
    z = np.load("tests/golden/haspi_ref.npz")
    for name in ("bundled_16000", "toy_train_multienh", "toy_train_clean", "toy_test_clean"):
        yield name, z[name + "/x"], z[name + "/y"]


def precision():
    for name, x, y in cases():
        A, Sxy, Syy = matrices(x, y)
        lam0, U0 = np.linalg.eigh(A)
        s0 = E.siib(lam0, U0, Sxy, Syy)
        out = []
        for nm, st, ar in (("s64a64", np.float64, np.float64), ("s32a64", np.float32, np.float64), ("s32a32", np.float32, np.float32)):
            d, e, V, tau = tridiag_var(A, st, ar)
            lam, Z = eigh_tridiagonal(d.astype(np.float64), e.astype(np.float64))
            out.append("%s %.1e" % (nm, abs(E.siib(lam, E.back(V, tau, Z), Sxy, Syy) - s0) / s0))
        l32, U32 = np.linalg.eigh(A.astype(np.float32))
        out.append("eigh32 %.1e" % (abs(E.siib(l32.astype(np.float64), U32.astype(np.float64), Sxy, Syy) - s0) / s0))
        print("%-22s cond %.1e SIIB %9.5f | %s" % (name, lam0[-1] / lam0[0], s0, "  ".join(out)), flush=True)


def iterations():
    from nele_gan_b200.synth import make_pair
    for i, L in ((1, 47999), (0, 52345), (5, 70003)):
        x, y, _ = make_pair(i, L)
        A, Sxy, Syy = matrices(x, y)
        lam0, U0 = np.linalg.eigh(A)
        s0 = E.siib(lam0, U0, Sxy, Syy)
        d, e, V, tau = tridiag_var(A, np.float32, np.float32)
        sc = max(np.abs(d).max(), np.abs(e).max())
        for iters in (58, 50, 46, 42, 38):
            lam = E.bisect_all(d / sc, e / sc, iters=iters) * sc
            Z = np.stack([E.getvec(d, e, l) for l in lam], axis=1)
            orth = np.abs(Z.T @ Z - np.eye(len(lam))).max()
            s1 = E.siib(lam, E.back(V, tau, Z), Sxy, Syy)
            print("pair %d L=%d iterations %d: orthogonality %.1e  SIIB rel %.1e" % (i, L, iters, orth, abs(s1 - s0) / s0), flush=True)


def back_seq32(V, tau, Z):
    """u = H_0 ... H_{n-3} z, one reflector at a time, FP32 throughout (what bt6::backtf6_kernel computes, up to its
    blocks of four)."""
    U = Z.astype(np.float32).copy()
    V32, t32 = V.astype(np.float32), tau.astype(np.float32)
    for k in range(V.shape[0] - 3, -1, -1):
        if t32[k] == 0:
            continue
        v = V32[:, k]
        U -= np.outer(t32[k] * v, v @ U)
    return U.astype(np.float64)


def back_wy32(V, tau, Z, nb):
    """The same product with the reflectors grouped into compact-WY blocks H_k0 ... H_k0+nb-1 = I - Vb T Vb^T
    (T upper triangular, LAPACK larft forward / columnwise), every product in FP32."""
    n = V.shape[0]
    U = Z.astype(np.float32).copy()
    V32, t32 = V.astype(np.float32), tau.astype(np.float32)
    nref = n - 2
    starts = list(range(0, nref, nb))
    for k0 in reversed(starts):
        k1 = min(k0 + nb, nref)
        Vb = V32[:, k0:k1]
        m = k1 - k0
        T = np.zeros((m, m), dtype=np.float32)
        G = (Vb.T @ Vb).astype(np.float32)          # Gram values, as the current kernel computes per panel
        for j in range(m):
            T[j, j] = t32[k0 + j]
            if j:
                T[:j, j] = -t32[k0 + j] * (T[:j, :j] @ G[:j, j])
        W = (Vb.T @ U).astype(np.float32)
        U -= Vb @ (T @ W).astype(np.float32)
    return U.astype(np.float64)


def blockwy():
    for name, x, y in cases():
        A, Sxy, Syy = matrices(x, y)
        lam0, U0 = np.linalg.eigh(A)
        s0 = E.siib(lam0, U0, Sxy, Syy)
        d, e, V, tau = tridiag_var(A, np.float32, np.float32)
        lam, Z = eigh_tridiagonal(d.astype(np.float64), e.astype(np.float64))
        out = ["seq64 %.1e" % (abs(E.siib(lam, E.back(V, tau, Z), Sxy, Syy) - s0) / s0),
               "seq32 %.1e" % (abs(E.siib(lam, back_seq32(V, tau, Z), Sxy, Syy) - s0) / s0)]
        for nb in (4, 16, 32):
            U = back_wy32(V, tau, Z, nb)
            out.append("wy%d %.1e (orth %.0e)" % (nb, abs(E.siib(lam, U, Sxy, Syy) - s0) / s0, np.abs(U.T @ U - np.eye(U.shape[1])).max()))
        print("%-22s SIIB %9.5f | %s" % (name, s0, "  ".join(out)), flush=True)


if __name__ == "__main__":
    mode = sys.argv[1] if len(sys.argv) > 1 else "precision"
    {"iterations": iterations, "blockwy": blockwy}.get(mode, precision)()
