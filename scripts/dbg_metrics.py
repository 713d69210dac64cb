"""Development probe: ESTOI / SIIB stages of the engine against the oracle."""
import sys, time
sys.path.insert(0, ".")
import numpy as np
from nele_gan_b200.engine import Engine
from nele_gan_b200.synth import make_pair
from oracle import pystoi_np as P, pysiib_np as S, intel_np

e = Engine(0)
cases = [(0, 33536), (1, 40000), (2, 48000), (3, 52345), (4, 16000 * 7 + 123)]
if len(sys.argv) > 1:
    cases = cases[: int(sys.argv[1])]
for i, L in cases:
    x, y, snr = make_pair(i, L)
    print("=== case", i, L, "snr", snr, flush=True)
    t = time.time()
    r = e.score_batch([x], [y], metrics=("estoi", "siib"), mapped=False, keep_stages=True)
    print("engine %.3fs scores %s status %x" % (time.time() - t, r.scores[0], r.status[0]), e.last_timing())
    # ---- ESTOI
    st = {}
    d = P.stoi(x.astype(np.float64), y.astype(np.float64), 16000, extended=True, stages=st)
    info = e.stage("estoi.info")
    x10 = e.stage("estoi.x10").reshape(2, -1)
    print(" estoi oracle %.6f engine %.6f diff %.2e" % (d, r.estoi[0], r.estoi[0] - d))
    print("  n10", info, len(st["x10"]), "x10 err", np.abs(x10[0] - st["x10"]).max(), np.abs(x10[1] - st["y10"]).max(), "scale", np.abs(st["x10"]).max())
    kept = e.stage("estoi.kept")
    print("  kept equal:", np.array_equal(kept, np.nonzero(st["mask"])[0]), len(kept), int(st["mask"].sum()))
    tob = e.stage("estoi.tob").reshape(2, -1, 15)
    if "x_tob" in st and tob.shape[1] == st["x_tob"].shape[1]:
        print("  tob rel err", np.abs(tob[0] - st["x_tob"].T).max() / st["x_tob"].max(), np.abs(tob[1] - st["y_tob"].T).max() / st["y_tob"].max())
    # ---- SIIB
    x64, y64 = x.astype(np.float64), y.astype(np.float64)
    M, act = intel_np.siib_tiling_factor(x, 16000)
    tile = e.stage("siib.tile")
    print(" siib tile engine", tile, "oracle M", M, "act", act)
    xt, yt = (np.hstack([x64] * M), np.hstack([y64] * M)) if M != 1 else (x64, y64)
    ss = {}
    t = time.time()
    ref = S.SIIB(xt, yt, 16000, gauss=True, stages=ss)
    tor = time.time() - t
    lam_o, I_o = ss["lam"], ss["I_ch"]
    nonnull = lam_o > 1e-9 * lam_o.max()
    ref_nn = max(0.0, 80 / 15 * float(np.sum(I_o[nonnull])))
    print("  oracle %.5f (%.2fs)  non-null part %.5f (%d comps)  engine %.5f  rel %.2e / %.2e" % (
        ref, tor, ref_nn, nonnull.sum(), r.siib[0], (r.siib[0] - ref) / ref, (r.siib[0] - ref_nn) / ref_nn))
    Fa = int(ss["vad"].sum())
    print("  Fa oracle", Fa, "F oracle", len(ss["vad"]))
    if tile[3] == Fa:
        ls = e.stage("siib.logspec").reshape(2, Fa, 32)
        print("  logspec err", np.abs(ls[0, :, :28] - ss["X"].T).max(), np.abs(ls[1, :, :28] - ss["Y"].T).max(), "pad", np.abs(ls[:, :, 28:]).max())
        Xs = S.stack_frames(ss["X"], 15); Ys = S.stack_frames(ss["Y"], 15)
        xm = Xs - Xs.mean(1, keepdims=True); ym = Ys - Ys.mean(1, keepdims=True)
        sxx = e.stage("siib.sxx").reshape(420, 420)
        sxy = e.stage("siib.sxy").reshape(420, 420); syy = e.stage("siib.syy").reshape(420, 420)
        print("  Sxx rel err %.2e  Sxy %.2e  Syy %.2e" % (np.abs(sxx - xm @ xm.T).max() / np.abs(sxx).max(),
              np.abs(sxy - xm @ ym.T).max() / np.abs(sxy).max(), np.abs(syy - ym @ ym.T).max() / np.abs(syy).max()))
    rk = e.stage("siib.rank")
    lam = np.sort(e.stage("siib.lambda"))[::-1]
    lo = np.sort(lam_o)[::-1] * (ss["nf"] - 1)
    print("  rank/sweeps", rk, " lambda rel err (top 50) %.2e  (all non-null) %.2e" % (
        np.abs(lam[:50] - lo[:50]).max() / lo[0], np.max(np.abs(lam[:rk[0]] - lo[:rk[0]]) / lo[:rk[0]])))
    e.set_profiling(True)
    e.score_batch([x], [y], metrics=("estoi", "siib"), mapped=False)
    print("  kernel ms:", {k: round(v[0], 3) for k, v in e.kernel_times().items()})
    e.set_profiling(False)

# ---- batch timing, non-degenerate lengths
from nele_gan_b200.synth import make_batch
for nb, L in ((592, 52345), (592, 48000)):
    refs, degs = make_batch(nb, L, unique=8)
    e.set_profiling(True)
    for it in range(2):
        t = time.time()
        r = e.score_batch(refs, degs, metrics=("siib", "estoi", "haspi"), mapped=False)
        dt = time.time() - t
    kt = e.kernel_times()
    print("batch", nb, "x", L, "wall %.3f s" % dt, "kernel ms", e.last_timing(), "->", nb * L / 16000 / (e.last_timing()[0] / 1e3), "audio-s/s")
    print("   ", {k: round(v[0], 2) for k, v in sorted(kt.items(), key=lambda kv: -kv[1][0])})
    e.set_profiling(False)
    print("   scores[0:3]", r.scores[:3].tolist())
