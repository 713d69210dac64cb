"""The numpy restatement (oracle/haspi_np.py) against outputs of the UNMODIFIED
reference pyHASPI/pyhaspi2.py (tests/golden/haspi_ref.npz, generated in the
build container by tests/golden/make_golden.py).  This is what pins the
oracle; tolerance 2e-8 absorbs ``a**3`` vs ``a*a*a`` style reassociation."""
import numpy as np
import pytest

from oracle import haspi_np
from tests.conftest import golden_dither

CASES = ["bundled_22050", "bundled_16000", "toy_train_multienh", "toy_train_clean",
         "toy_test_clean", "synth_0_24000", "synth_1_31999", "synth_2_48000"]
TOL = 2e-8


@pytest.mark.parametrize("name", CASES)
def test_v2_zero_noise_and_stages(golden, name):
    g = golden[name]
    fs = int(g["fs"])
    st = {}
    s, raw = haspi_np.haspi_v2(g["x"], fs, g["y"], fs, noise=None, stages=st)
    assert abs(s - float(g["v2_zero"])) < TOL
    assert np.max(np.abs(raw - g["v2_zero_raw"])) < TOL
    assert np.max(np.abs(st["bwx"] - g["bwx"])) < 1e-10
    assert np.max(np.abs(st["bwy"] - g["bwy"])) < 1e-10
    assert st["xcep"].shape[0] == int(g["nsel"])
    assert np.max(np.abs(st["xlp"][::7] - g["xlp"])) < 1e-4   # fixture is float32
    assert np.max(np.abs(st["ylp"][::7] - g["ylp"])) < 1e-4


@pytest.mark.parametrize("name", CASES[:5])
def test_v2_shared_dither(golden, name):
    g = golden[name]
    fs = int(g["fs"])
    noise = {"dither_x": golden_dither(0), "dither_y": golden_dither(1)}
    s, raw = haspi_np.haspi_v2(g["x"], fs, g["y"], fs, noise=noise)
    assert abs(s - float(g["v2_dith"])) < TOL
    assert np.max(np.abs(raw - g["v2_dith_raw"])) < TOL


@pytest.mark.parametrize("name", ["bundled_22050", "toy_test_clean"])
def test_v2_numpy_stream_matches_reference_seed0(golden, name):
    g = golden[name]
    fs = int(g["fs"])
    np.random.seed(0)
    s, _ = haspi_np.haspi_v2(g["x"], fs, g["y"], fs, noise="numpy")
    assert abs(s - float(g["v2_seed0"])) < TOL


@pytest.mark.parametrize("name", ["bundled_22050", "toy_test_clean", "synth_0_24000"])
def test_v1_zero_noise(golden, name):
    g = golden[name]
    fs = int(g["fs"])
    s, raw = haspi_np.haspi(g["x"], fs, g["y"], fs, noise=None)
    assert abs(s - float(g["v1_zero"])) < TOL
    assert np.max(np.abs(raw - g["v1_zero_raw"])) < 1e-7


def test_v1_numpy_stream_matches_reference_seed0(golden):
    g = golden["toy_test_clean"]
    np.random.seed(0)
    s, _ = haspi_np.haspi(g["x"], 16000, g["y"], 16000, noise="numpy")
    assert abs(s - float(g["v1_seed0"])) < TOL


def test_properties_identical_and_gain_invariance(golden):
    g = golden["synth_0_24000"]
    x, y = g["x"], g["y"]
    s_same, _ = haspi_np.haspi_v2(x, 16000, x, 16000, noise=None)
    assert abs(s_same - haspi_np.HASPI2_WEIGHTS.sum()) < 1e-6
    s1, _ = haspi_np.haspi_v2(x, 16000, y, 16000, noise=None)
    s2, _ = haspi_np.haspi_v2(x, 16000, (7.3 * y).astype(np.float32), 16000, noise=None)
    assert abs(s1 - s2) < 1e-5


def test_below_threshold_raises():
    x = np.zeros(16000, dtype=np.float32)
    x[::2] = 1e-3
    with pytest.raises(Exception):
        # a constant-ish tiny signal normalises fine; silence in x must raise
        haspi_np.cep_coef(np.zeros((100, 32)), np.zeros((100, 32)))


@pytest.mark.parametrize("name", ["bundled_22050", "toy_train_multienh", "toy_test_clean", "synth_2_48000"])
def test_hasqi_v2_zero_noise(golden, name):
    """oracle hasqi_v2 against the outputs of the unmodified reference (tests/golden/hasqi_ref.npz)."""
    import os
    from oracle import haspi_np
    z = np.load(os.path.join(os.path.dirname(__file__), "golden", "hasqi_ref.npz"))
    g = golden[name]
    fs = int(g["fs"])
    st = {}
    comb, nonlin, lin, raw = haspi_np.hasqi_v2(g["x"], fs, g["y"], fs, noise=None, stages=st)
    assert np.abs(st["xsl"] - z[name + "/xsl"]).max() < 1e-6
    assert np.abs(st["ysl"] - z[name + "/ysl"]).max() < 1e-6
    assert np.allclose([comb, nonlin, lin], z[name + "/hq_zero"], rtol=0, atol=1e-7)
    assert np.allclose(raw, z[name + "/hq_zero_raw"], rtol=0, atol=1e-7)
