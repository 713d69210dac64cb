"""GPU parity of the ESTOI and SIIB^Gauss paths (through the C ABI) against the
CPU oracle.  Tolerances are BASELINE.json's: |dESTOI| <= 1e-3, SIIB within 0.5 %
relative.  The stage checks are much tighter so a drift shows up early."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

ESTOI_TOL = 1e-3
SIIB_RTOL = 5e-3


@pytest.fixture(scope="module")
def eng():
    from nele_gan_b200.engine import Engine
    return Engine(0)


def _pairs():
    from nele_gan_b200.synth import make_pair
    # lengths that are NOT multiples of the 200-sample hop: the tiled signal has full rank
    return [make_pair(i, L)[:2] for i, L in ((0, 33536), (3, 52345), (5, 40123), (6, 16000 * 6 + 77))]


def _oracle_siib(x, y, stages=None):
    from oracle import intel_np, pysiib_np
    x64, y64 = x.astype(np.float64), y.astype(np.float64)
    M, act = intel_np.siib_tiling_factor(x, 16000)
    if M != 1:
        x64, y64 = np.hstack([x64] * M), np.hstack([y64] * M)
    return pysiib_np.SIIB(x64, y64, 16000, gauss=True, stages=stages), M, act


# ------------------------------------------------------------------ ESTOI
def test_estoi_scores_and_stages(eng, golden):
    from oracle import pystoi_np
    xs = [golden[n]["x"] for n in ("toy_train_multienh", "toy_train_clean", "toy_test_clean", "synth_1_31999")]
    ys = [golden[n]["y"] for n in ("toy_train_multienh", "toy_train_clean", "toy_test_clean", "synth_1_31999")]
    r = eng.score_batch(xs, ys, metrics=("estoi",), mapped=False, keep_stages=True)
    for i, (x, y) in enumerate(zip(xs, ys)):
        st = {}
        d = pystoi_np.stoi(x.astype(np.float64), y.astype(np.float64), 16000, extended=True, stages=st)
        assert r.metric_status("estoi")[i] == 0
        assert abs(r.estoi[i] - d) < ESTOI_TOL
        assert abs(r.estoi[i] - d) < 1e-5          # what the FP32 path actually achieves
        x10 = eng.stage("estoi.x10", i).reshape(2, -1)
        assert x10.shape[1] == len(st["x10"])
        assert np.abs(x10[0] - st["x10"]).max() < 1e-6 * np.abs(st["x10"]).max()
        assert np.array_equal(eng.stage("estoi.kept", i), np.nonzero(st["mask"])[0])
        tob = eng.stage("estoi.tob", i).reshape(2, -1, 15)
        assert np.abs(tob[0] - st["x_tob"].T).max() < 1e-5 * st["x_tob"].max()
        assert np.abs(tob[1] - st["y_tob"].T).max() < 1e-5 * st["y_tob"].max()


def test_estoi_mapped_and_identity(eng, golden):
    x = golden["toy_test_clean"]["x"]
    y = golden["toy_test_clean"]["y"]
    raw = eng.score_batch([x, x], [y, x], metrics=("estoi",), mapped=False).estoi
    assert abs(raw[1] - 1.0) < 1e-5                 # identical signals correlate perfectly
    m = eng.score_batch([x, x], [y, x], metrics=("estoi",), mapped=True).estoi
    assert np.allclose(m, 1 / (1 + np.exp(-8.0 * (raw - 0.25))), atol=1e-12)   # intel.py:136-140
    # both signals are RMS-free: ESTOI is invariant to the gain of either input
    g = eng.score_batch([(3.0 * x).astype(np.float32)], [(0.2 * y).astype(np.float32)], metrics=("estoi",), mapped=False).estoi
    assert abs(g[0] - raw[0]) < 1e-5


def test_estoi_too_short_returns_sentinel(eng):
    rng = np.random.default_rng(1)
    x = rng.standard_normal(4000).astype(np.float32)   # 2500 samples at 10 kHz -> < 30 frames
    r = eng.score_batch([x], [x], metrics=("estoi",), mapped=False)
    assert r.estoi[0] == 1e-5 and r.metric_status("estoi")[0] == 2    # pystoi's sentinel + warning


def test_estoi_other_rates(eng, golden):
    from oracle import pystoi_np
    from scipy.signal import resample_poly
    x = golden["toy_test_clean"]["x"].astype(np.float64)
    y = golden["toy_test_clean"]["y"].astype(np.float64)
    for fs, (u, d) in ((10000, (5, 8)), (8000, (1, 2))):
        xr, yr = resample_poly(x, u, d).astype(np.float32), resample_poly(y, u, d).astype(np.float32)
        want = pystoi_np.stoi(xr.astype(np.float64), yr.astype(np.float64), fs, extended=True)
        got = eng.score_batch([xr], [yr], fs=fs, metrics=("estoi",), mapped=False).estoi[0]
        assert abs(got - want) < 1e-5, fs


# ------------------------------------------------------------------- SIIB
def test_siib_scores_and_stages_full_rank(eng):
    from oracle import pysiib_np
    pairs = _pairs()
    r = eng.score_batch([p[0] for p in pairs], [p[1] for p in pairs], metrics=("siib",), mapped=False, keep_stages=True)
    for i, (x, y) in enumerate(pairs):
        st = {}
        want, M, act = _oracle_siib(x, y, st)
        tile = eng.stage("siib.tile", i)
        assert tile[0] == M and tile[1] == act
        assert tile[2] == len(st["vad"]) and tile[3] == int(st["vad"].sum())
        ls = eng.stage("siib.logspec", i).reshape(2, -1, 32)
        assert np.abs(ls[0, :, :28] - st["X"].T).max() < 1e-4
        assert np.abs(ls[1, :, :28] - st["Y"].T).max() < 1e-4
        assert np.abs(ls[:, :, 28:]).max() == 0.0
        Xs = pysiib_np.stack_frames(st["X"], 15)
        xm = Xs - Xs.mean(1, keepdims=True)
        sxx = eng.stage("siib.sxx", i).reshape(420, 420)
        assert np.abs(sxx - xm @ xm.T).max() < 1e-5 * np.abs(sxx).max()
        # Sxy / Syy are kept folded onto the lower triangle (u^T S u only sees the symmetric part of S)
        Ys = pysiib_np.stack_frames(st["Y"], 15)
        ym = Ys - Ys.mean(1, keepdims=True)
        low = np.tril_indices(420)
        for name, S in (("siib.sxy", xm @ ym.T), ("siib.syy", ym @ ym.T)):
            F = S + S.T
            F[np.diag_indices(420)] = np.diag(S)
            got = eng.stage(name, i).reshape(420, 420)
            assert np.abs(got[low] - F[low]).max() < 2e-5 * np.abs(F).max(), name
        rk = eng.stage("siib.rank", i)
        assert rk[0] == 420 and (rk[1] == -1 or 0 < rk[1] < 14)      # -1: tridiagonalisation path (no sweeps)
        lam = np.sort(eng.stage("siib.lambda", i))[::-1]
        lo = np.sort(st["lam"])[::-1] * (st["nf"] - 1)
        assert np.abs(lam[:50] - lo[:50]).max() < 1e-3 * lo[0]
        assert r.metric_status("siib")[i] == 0
        assert abs(r.siib[i] - want) < SIIB_RTOL * want
        assert abs(r.siib[i] - want) < 1e-3 * want      # what the mixed FP64/FP32 path achieves


def test_siib_rank_deficient_tiling(eng):
    """A tiled signal whose length is a multiple of the 200-sample hop repeats
    its frames exactly: cov(X) is singular, the reference's eigenvectors in the
    null space are rounding noise.  The engine drops the null space (the exact-
    arithmetic value); it must agree with the oracle's sum over the components
    whose eigenvalue is numerically non-zero."""
    from nele_gan_b200.synth import make_pair
    for i, L in ((1, 40000), (2, 48000)):
        x, y, _ = make_pair(i, L)
        st = {}
        want_all, M, _ = _oracle_siib(x, y, st)
        nonnull = st["lam"] > 1e-9 * st["lam"].max()
        want = 80.0 / 15 * float(np.sum(st["I_ch"][nonnull]))
        r = eng.score_batch([x], [y], metrics=("siib",), mapped=False, keep_stages=True)
        assert eng.stage("siib.rank")[0] == int(nonnull.sum())
        assert abs(r.siib[0] - want) < 1e-3 * want
        assert r.siib[0] <= want_all * 1.001           # the reference adds a small positive junk term


def test_siib_mapped_no_tile_and_too_short(eng):
    from nele_gan_b200.synth import make_pair
    x, y, _ = make_pair(0, 33536)
    raw = eng.score_batch([x], [y], metrics=("siib",), mapped=False).siib[0]
    m = eng.score_batch([x], [y], metrics=("siib",), mapped=True).siib[0]
    assert abs(m - 1 / (1 + np.exp(-0.06 * (raw - 32)))) < 1e-12            # intel.py:102-106
    # plain pysiib semantics: 2 s of speech is less than the 20 s SIIB needs
    r = eng.score_batch([x], [y], metrics=("siib",), mapped=False, siib_no_tile=True)
    assert np.isnan(r.siib[0]) and r.metric_status("siib")[0] == 2
    # >= 20 s of active speech: the wrapper does not tile (M = 1)
    xl = np.concatenate([make_pair(10 + k, 48000)[0] for k in range(24)])
    yl = np.concatenate([make_pair(10 + k, 48000)[1] for k in range(24)])
    want, M, _ = _oracle_siib(xl, yl)
    rr = eng.score_batch([xl], [yl], metrics=("siib",), mapped=False, keep_stages=True)
    assert eng.stage("siib.tile")[0] == M
    assert abs(rr.siib[0] - want) < SIIB_RTOL * want


def test_siib_bad_rate(eng):
    x = np.random.default_rng(0).standard_normal(30000).astype(np.float32)
    r = eng.score_batch([x], [x], fs=8000, metrics=("siib",), mapped=False)
    assert r.metric_status("siib")[0] == 3 and np.isnan(r.siib[0])


# ------------------------------------------------------------ all together
def test_all_metrics_ragged_batch_matches_oracle(eng):
    from nele_gan_b200.synth import make_pair
    from oracle import intel_np
    pairs = [make_pair(20 + i, L)[:2] for i, L in enumerate((32001, 47999, 56789, 40411, 35555, 61003))]
    r = eng.score_batch([p[0] for p in pairs], [p[1] for p in pairs], mapped=False, no_dither=True)
    rm = eng.score_batch([p[0] for p in pairs], [p[1] for p in pairs], mapped=True, no_dither=True)
    for i, (x, y) in enumerate(pairs):
        want = intel_np.score_pair(x, y, 16000, norm=False, noise=None)
        assert abs(r.siib[i] - want[0]) < SIIB_RTOL * want[0]
        assert abs(r.haspi[i] - want[1]) < 1e-3
        assert abs(r.estoi[i] - want[2]) < ESTOI_TOL
        wm = intel_np.score_pair(x, y, 16000, norm=True, noise=None)
        assert np.abs(rm.scores[i] - wm).max() < 2e-3
    # one call per pair gives the same numbers as the batch (no cross-talk between pairs)
    one = eng.score_batch([pairs[2][0]], [pairs[2][1]], mapped=False, no_dither=True)
    assert np.allclose(one.scores[0], r.scores[2], rtol=1e-6, atol=1e-9)


def test_profiling_lists_every_kernel(eng):
    from nele_gan_b200.synth import make_pair
    x, y, _ = make_pair(1, 33001)
    eng.set_profiling(True)
    eng.score_batch([x], [y], mapped=False)
    kt = eng.kernel_times()
    eng.set_profiling(False)
    for k in ("haspi_ear", "estoi_tob", "siib_cov", "siib_chol", "siib_gram", "siib_tridiag", "siib_backtf", "siib_quad", "siib_projquad"):
        assert k in kt and kt[k][0] > 0 and kt[k][1] >= 1
    ms, launches = eng.last_timing()
    assert launches == sum(v[1] for v in kt.values())


def test_siib_klt_dispatch_paths(eng):
    """Every KLT route against the oracle (components with a numerically non-zero eigenvalue):
    non-repeating tiling -> tridiagonalisation of Sxx; exactly repeating tiling of low rank ->
    Cholesky + Gram matrix; repeating tiling with a long period (rank > 112) -> Cholesky, then
    tridiagonalisation.  `siib.rank` stage: [rank, marker] with marker -1 / -2 for the two paths."""
    from nele_gan_b200.synth import make_pair
    cases = ((7, 36123, -1), (8, 40000, -2), (9, 50000, -2), (10, 44100, None), (11, 30000, -2))
    xs, ys = zip(*[make_pair(i, L)[:2] for i, L, _ in cases])
    r = eng.score_batch(list(xs), list(ys), metrics=("siib",), mapped=False, keep_stages=True)
    for n, (i, L, marker) in enumerate(cases):
        st = {}
        _oracle_siib(xs[n], ys[n], st)
        nonnull = st["lam"] > 1e-9 * st["lam"].max()
        want = 80.0 / 15 * float(np.sum(st["I_ch"][nonnull]))
        rk = eng.stage("siib.rank", n)
        assert abs(r.siib[n] - want) < 2e-3 * want, (L, r.siib[n], want, rk[:2])
        if marker is not None:
            assert rk[1] == marker, (L, rk[:2])
        if marker == -2:
            assert rk[0] == int(nonnull.sum())


def test_all_metrics_on_the_toy_corpus_pairs(eng, golden):
    """BASELINE.json configs[1]: the three labels of every toy_dataset utterance -- Train (Clean, MultiEnh + Noise),
    Train (Clean, Clean + Noise), Test (Clean, Clean + Noise); 33 536 / 34 048 samples, so the SIIB tiling has full
    rank -- raw and mapped, against the oracle on the same arrays (HASPI also against the reference's own value)."""
    from oracle import intel_np
    names = ("toy_train_multienh", "toy_train_clean", "toy_test_clean")
    xs, ys = [golden[n]["x"] for n in names], [golden[n]["y"] for n in names]
    mapped = eng.score_batch(xs, ys, mapped=True, no_dither=True)
    raw = eng.score_batch(xs, ys, mapped=False, no_dither=True, keep_stages=True)
    assert raw.ok.all() and not raw.siib_nullspace_dropped.any()
    for i, n in enumerate(names):
        want = intel_np.score_pair(xs[i], ys[i], 16000, norm=False, noise=None)
        assert abs(raw.siib[i] - want[0]) < SIIB_RTOL * want[0], (n, raw.siib[i], want[0])
        assert abs(raw.siib[i] - want[0]) < 1e-3 * want[0], (n, raw.siib[i], want[0])     # FP32 KLT: measured <= 3e-4
        assert abs(raw.haspi[i] - want[1]) < 1e-3 and abs(raw.haspi[i] - float(golden[n]["v2_zero"])) < 1e-3
        assert abs(raw.estoi[i] - want[2]) < ESTOI_TOL
        wm = intel_np.score_pair(xs[i], ys[i], 16000, norm=True, noise=None)
        assert np.abs(mapped.scores[i] - wm).max() < 2e-3
        rk = eng.stage("siib.rank", i)
        assert rk[0] == 420 and rk[1] == -1                       # tridiagonalisation path at full rank
