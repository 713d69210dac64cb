"""GPU parity of HASPI version 1 -- haspi() of pyhaspi2.py:109-157 -- through the
C ABI (NELE_FLAG_HASPI_V1) against the oracle and the golden outputs of the
unmodified reference.  Tolerance on Intel: 1e-3 (BASELINE.json's HASPI bound);
the stage checks (smoothed envelopes, segment covariances) are per-element."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

CASES = ["bundled_16000", "toy_train_multienh", "toy_train_clean", "toy_test_clean",
         "synth_0_24000", "synth_1_31999", "synth_2_48000"]
TOL = 1e-3


@pytest.fixture(scope="module")
def eng():
    from nele_gan_b200.engine import Engine
    return Engine(0)


def _smooth_from_segsum(seg, nseg):
    """eb_EnvSmooth segments (pyhaspi2.py:692-700) from the engine's block sums."""
    w = np.hanning(384)
    wsum, hsum = w.sum(), w[192:].sum()
    rise, fall = seg[0], seg[1]
    out = np.zeros((nseg, 32))
    out[0] = fall[0] / hsum
    for s in range(1, nseg - 1):
        out[s] = (rise[s] + fall[s + 1]) / wsum
    out[nseg - 1] = rise[nseg - 1] / hsum
    return out


@pytest.mark.parametrize("name", ["bundled_22050"] + CASES)
def test_zero_noise_score_and_stages(eng, golden, name):
    from oracle import haspi_np
    g = golden[name]
    fs = int(g["fs"])
    x, y = haspi_np._prep(g["x"], g["y"])
    xdb, ydb, xbm, ybm, _, _ = haspi_np.ear_model(x, fs, y, fs, None, noise=None, want_bm=True)
    sm_x, sm_y = haspi_np.env_smooth(xdb), haspi_np.env_smooth(ydb)
    cov, msx, _ = haspi_np.bm_covary(xbm, ybm)
    r = eng.score_batch([g["x"]], [g["y"]], fs=fs, metrics=("haspi",), mapped=False, no_dither=True,
                        keep_stages=True, haspi_v1=True)
    assert r.metric_status("haspi")[0] == 0
    n24 = xdb.shape[1]
    nblk, nseg = (n24 + 191) // 192, sm_x.shape[1]
    seg = eng.stage("haspi1.segsum").reshape(4, nblk, 32)
    gx, gy = _smooth_from_segsum(seg[0:2], nseg), _smooth_from_segsum(seg[2:4], nseg)
    assert np.sqrt(np.mean((gx - sm_x.T) ** 2)) < 5e-3          # dB SL, values 0..~60
    assert np.sqrt(np.mean((gy - sm_y.T) ** 2)) < 5e-3
    gcov = eng.stage("haspi1.cov").reshape(nseg, 32)
    gms = eng.stage("haspi1.msx").reshape(nseg, 32)
    assert np.abs(gms - msx.T).max() < 2e-3 * max(1.0, msx.max())
    loud = msx.T > 1.0                                           # covariance of inaudible cells is noise over noise
    assert np.abs(gcov - cov.T)[loud].max() < 5e-3
    assert np.sqrt(np.mean((gcov - cov.T)[loud] ** 2)) < 5e-4
    assert abs(r.haspi[0] - float(g["v1_zero"])) < TOL
    assert np.abs(r.haspi_raw[0, :4] - g["v1_zero_raw"]).max() < TOL
    assert np.all(np.isnan(r.haspi_raw[0, 4:]))


def test_batch_matches_single_and_api(eng, golden):
    from nele_gan_b200 import api
    xs = [golden[n]["x"] for n in CASES]
    ys = [golden[n]["y"] for n in CASES]
    r = eng.score_batch(xs, ys, fs=16000, metrics=("haspi",), mapped=True, no_dither=True, haspi_v1=True)
    for i, n in enumerate(CASES):
        assert abs(r.haspi[i] - float(golden[n]["v1_zero"])) < TOL      # mapped is ignored for version 1
    g = golden["toy_test_clean"]
    np.random.seed(0)
    s, raw = api.haspi(g["x"], 16000, g["y"], 16000)
    # the reference's own seed-to-seed spread (BM threshold noise) is ~2.5e-3
    assert abs(s - float(g["v1_seed0"])) < 1e-2 and raw.shape == (4,)
    s2, _ = api.haspi(g["x"], 16000, g["y"], 16000, alpha=-2.0, seed=3)
    arg = -9.047 + 14.816 * _[0] + 4.616 * _[3]
    assert abs(s2 - 1 / (1 + np.exp(-2.0 * arg))) < 1e-12


def test_noise_is_deterministic_per_seed(eng, golden):
    g = golden["synth_1_31999"]
    f = lambda seed: eng.score_batch([g["x"]], [g["y"]], metrics=("haspi",), mapped=False, seed=seed, haspi_v1=True).haspi[0]
    a, b, c = f(5), f(5), f(6)
    assert a == b and a != c
    assert abs(a - float(g["v1_zero"])) < 1e-2


def test_identical_signals_and_gain_invariance(eng, golden):
    g = golden["synth_0_24000"]
    x, y = g["x"], g["y"]
    r = eng.score_batch([x, x, x], [x, y, (3.1 * y).astype(np.float32)], metrics=("haspi",), mapped=False,
                        no_dither=True, haspi_v1=True)
    assert abs(r.haspi_raw[0, 0] - 1.0) < 1e-5 and abs(r.haspi_raw[0, 3] - 1.0) < 1e-4
    assert abs(r.haspi[1] - r.haspi[2]) < 1e-5


def test_too_short_is_below_threshold(eng):
    x = (0.01 * np.random.default_rng(0).standard_normal(100)).astype(np.float32)
    r = eng.score_batch([x], [x], metrics=("haspi",), mapped=False, no_dither=True, haspi_v1=True)
    assert r.metric_status("haspi")[0] == 1 and np.isnan(r.haspi[0])


@pytest.mark.parametrize("name", ["bundled_22050"] + CASES)
def test_hasqi_v2_against_the_reference(eng, golden, name):
    """hasqi_v2 (pyhaspi2.py:32-74) on the version-1 pipeline, against the outputs of the unmodified
    reference with the BM noise forced to zero (tests/golden/hasqi_ref.npz)."""
    import os
    z = np.load(os.path.join(os.path.dirname(__file__), "golden", "hasqi_ref.npz"))
    g = golden[name]
    fs = int(g["fs"])
    r = eng.score_batch([g["x"]], [g["y"]], fs=fs, metrics=("haspi",), mapped=False, no_dither=True, hasqi=True)
    assert r.metric_status("haspi")[0] == 0
    comb, nonlin, lin = z[name + "/hq_zero"]
    assert abs(r.haspi[0] - comb) < TOL
    assert np.abs(r.haspi_raw[0, :4] - z[name + "/hq_zero_raw"]).max() < TOL
    assert abs(r.haspi_raw[0, 4] - nonlin) < TOL and abs(r.haspi_raw[0, 5] - lin) < TOL


def test_hasqi_api_and_identity(eng, golden):
    from nele_gan_b200 import api
    g = golden["synth_0_24000"]
    comb, nonlin, lin, raw = api.hasqi_v2(g["x"], 16000, g["x"], 16000, seed=1)
    assert abs(lin - 1.0) < 1e-6 and abs(raw[0] - 1.0) < 1e-4 and comb > 0.95 and len(raw) == 4
