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
