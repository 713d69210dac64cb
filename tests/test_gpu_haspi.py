"""GPU parity of the HASPI v2 path (through the C ABI) against the oracle and
the reference-generated golden fixtures.  Tolerance: |dHASPI| <= 1e-3 on the
raw score as BASELINE.json states; the stage checks are tighter."""
import numpy as np
import pytest

from tests.conftest import golden_dither

pytestmark = pytest.mark.gpu

CASES16 = ["bundled_16000", "toy_train_multienh", "toy_train_clean", "toy_test_clean",
           "synth_0_24000", "synth_1_31999", "synth_2_48000"]
TOL = 1e-3


@pytest.fixture(scope="module")
def eng():
    from nele_gan_b200.engine import Engine
    return Engine(0)


@pytest.mark.parametrize("name", ["bundled_22050"] + CASES16)
def test_zero_dither_score_and_stages(eng, golden, name):
    from oracle import haspi_np
    g = golden[name]
    fs = int(g["fs"])
    st = {}
    s_or, raw_or = haspi_np.haspi_v2(g["x"], fs, g["y"], fs, noise=None, stages=st)
    r = eng.score_batch([g["x"]], [g["y"]], fs=fs, metrics=("haspi",), mapped=False, no_dither=True,
                        keep_stages=True)
    assert r.metric_status("haspi")[0] == 0
    n24 = len(st["xmid"])
    mid = eng.stage("haspi.mid").reshape(2, n24)
    scale = np.abs(st["xmid"]).max()
    assert np.abs(mid[0] - st["xmid"]).max() < 2e-6 * scale
    assert np.abs(mid[1] - st["ymid"]).max() < 2e-6 * np.abs(st["ymid"]).max()
    bw = eng.stage("haspi.bw").reshape(2, 32)
    assert np.abs(bw[0] - g["bwx"]).max() < 2e-5
    assert np.abs(bw[1] - g["bwy"]).max() < 2e-5
    assert np.array_equal(eng.stage("haspi.shift"), st["shifts"])
    nsub = st["xlp"].shape[0]
    env = eng.stage("haspi.envlp").reshape(2, nsub, 32)
    assert np.sqrt(np.mean((env[0] - st["xlp"]) ** 2)) < 5e-3
    assert np.sqrt(np.mean((env[1] - st["ylp"]) ** 2)) < 5e-3
    assert abs(int(eng.stage("haspi.nsel")[0]) - int(g["nsel"])) <= 1
    assert abs(r.haspi[0] - float(g["v2_zero"])) < TOL
    assert np.abs(r.haspi_raw[0] - g["v2_zero_raw"]).max() < TOL
    assert abs(r.haspi[0] - s_or) < TOL


@pytest.mark.parametrize("name", CASES16[:4])
def test_shared_dither(eng, golden, name):
    g = golden[name]
    d = np.stack([golden_dither(0), golden_dither(1)]).astype(np.float32)
    r = eng.score_batch([g["x"]], [g["y"]], fs=16000, metrics=("haspi",), mapped=False, dither=d)
    assert abs(r.haspi[0] - float(g["v2_dith"])) < TOL
    assert np.abs(r.haspi_raw[0] - g["v2_dith_raw"]).max() < TOL


def test_ragged_batch_matches_single_calls(eng, golden):
    xs = [golden[n]["x"] for n in CASES16]
    ys = [golden[n]["y"] for n in CASES16]
    r = eng.score_batch(xs, ys, fs=16000, metrics=("haspi",), mapped=False, no_dither=True)
    for i, n in enumerate(CASES16):
        assert abs(r.haspi[i] - float(golden[n]["v2_zero"])) < TOL
    rm = eng.score_batch(xs, ys, fs=16000, metrics=("haspi",), mapped=True, no_dither=True)
    assert np.allclose(rm.haspi, 1 / (1 + np.exp(-0.95 * (r.haspi - 2.8))), atol=1e-12)


def test_philox_dither_is_deterministic_and_close(eng, golden):
    g = golden["toy_test_clean"]
    a = eng.score_batch([g["x"]], [g["y"]], metrics=("haspi",), mapped=False, seed=7).haspi[0]
    b = eng.score_batch([g["x"]], [g["y"]], metrics=("haspi",), mapped=False, seed=7).haspi[0]
    c = eng.score_batch([g["x"]], [g["y"]], metrics=("haspi",), mapped=False, seed=8).haspi[0]
    assert a == b
    assert a != c
    # seed-to-seed spread of the reference itself is 2.4e-3
    assert abs(a - float(g["v2_zero"])) < 1e-2 and abs(c - float(g["v2_zero"])) < 1e-2


def test_properties(eng, golden):
    g = golden["synth_0_24000"]
    x, y = g["x"], g["y"]
    same = eng.score_batch([x], [x], metrics=("haspi",), mapped=False, no_dither=True).haspi[0]
    assert abs(same - 11.467) < 1e-3
    a = eng.score_batch([x, x], [y, (7.3 * y).astype(np.float32)], metrics=("haspi",), mapped=False, no_dither=True).haspi
    assert abs(a[0] - a[1]) < 1e-4


def test_below_threshold_status(eng):
    x = np.zeros(16000, dtype=np.float32)
    x[100] = 1.0  # a click: almost every frame is below the loudness threshold
    y = x.copy()
    r = eng.score_batch([x], [y], metrics=("haspi",), mapped=False, no_dither=True)
    assert r.metric_status("haspi")[0] in (0, 1)


def test_bad_rate_status(eng):
    x = np.random.default_rng(0).standard_normal(48000).astype(np.float32)
    r = eng.score_batch([x], [x], fs=48000, metrics=("haspi",), mapped=False)
    assert r.metric_status("haspi")[0] == 3 and np.isnan(r.haspi[0])


def _extra():
    import os
    return np.load(os.path.join(os.path.dirname(__file__), "golden", "haspi_extra.npz"))


def test_nonzero_audiogram_matches_the_reference(eng, golden):
    """HL != 0 (pyhaspi2.py:1160-1171: hearing-loss dependent OHC / IHC attenuation, compression ratio and
    bandwidth of the processed signal's path).  Golden: the unmodified reference with HL = [20 .. 60] dB on a toy
    pair, zero dither (tests/golden/make_golden_extra.py); the score drops from 2.69 to 1.11."""
    z = _extra()
    g = golden[str(z["hl/case"])]
    r = eng.score_batch([g["x"]], [g["y"]], fs=16000, metrics=("haspi",), mapped=False, no_dither=True,
                        hl=z["hl/HL"], keep_stages=True)
    bw = eng.stage("haspi.bw").reshape(2, 32)
    assert np.abs(bw[0] - z["hl/bwx"]).max() < 2e-5
    assert np.abs(bw[1] - z["hl/bwy"]).max() < 2e-5
    assert abs(int(eng.stage("haspi.nsel")[0]) - int(z["hl/nsel"])) <= 1
    assert abs(r.haspi[0] - float(z["hl/v2_zero"])) < TOL
    assert np.abs(r.haspi_raw[0] - z["hl/v2_zero_raw"]).max() < TOL
    assert abs(float(z["hl/v2_zero"]) - float(g["v2_zero"])) > 1.0      # the audiogram matters: not a no-op
    # a later call without HL must not inherit the loss tables
    r0 = eng.score_batch([g["x"]], [g["y"]], fs=16000, metrics=("haspi",), mapped=False, no_dither=True)
    assert abs(r0.haspi[0] - float(g["v2_zero"])) < TOL


def test_philox_dither_reproduces_the_reference_distribution(eng, golden):
    """The reference's score is a random variable (0.1 dB dither from numpy's global stream).  The engine's
    own dither (Philox, keyed by seed and pair) must have the same mean: 128 independent engine draws against the
    reference run as a user runs it under np.random.seed(0 .. 31) -- means within 3e-4 (standard error of the
    difference 1.3e-4), spreads within a factor 1.5."""
    z = _extra()
    g = golden[str(z["seeds/case"])]
    ref = z["seeds/v2"]
    r = eng.score_batch([g["x"]] * 128, [g["y"]] * 128, fs=16000, metrics=("haspi",), mapped=False, seed=2024)
    assert np.unique(r.haspi).size > 100                                  # every copy draws its own dither
    assert abs(r.haspi.mean() - ref.mean()) < 3e-4, (r.haspi.mean(), ref.mean())
    assert ref.std() / 1.5 < r.haspi.std() < 1.5 * ref.std(), (r.haspi.std(), ref.std())
