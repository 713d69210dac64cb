"""CPU checks of the ESTOI / SIIB restatements (oracle/pystoi_np.py,
oracle/pysiib_np.py).  pystoi and pysiib are un-vendored dependencies of the
reference (README.md:13-14) with no source, test or golden value in the
reference tree, so these oracles are PARITY UNPINNED; what can be pinned is
(a) the helper functions the reference keeps in-tree (intel.py:16-54) and its
logistic mappings: tests/golden/intel_helpers.npz holds their outputs, produced by
tests/golden/make_golden_intel.py from the unmodified intel.py source text,
(b) the published properties of the metrics, and (c) scipy's own resampler."""
import os
import numpy as np
import pytest

from oracle import intel_np, pysiib_np, pystoi_np
from nele_gan_b200.synth import make_pair


def test_estoi_properties():
    x, y, _ = make_pair(0, 33536)
    x, y = x.astype(np.float64), y.astype(np.float64)
    assert abs(pystoi_np.stoi(x, x, 16000, extended=True) - 1.0) < 1e-9
    d = pystoi_np.stoi(x, y, 16000, extended=True)
    assert 0.0 < d < 1.0
    assert abs(pystoi_np.stoi(3 * x, 0.1 * y, 16000, extended=True) - d) < 1e-9      # gain invariant
    rng = np.random.default_rng(0)
    worse = pystoi_np.stoi(x, y + 0.2 * rng.standard_normal(len(y)), 16000, extended=True)
    assert worse < d                                                                  # more noise, lower score
    with pytest.warns(RuntimeWarning):
        assert pystoi_np.stoi(x[:4000], y[:4000], 16000, extended=True) == 1e-5       # pystoi's sentinel


def test_estoi_resampler_is_resample_poly_with_octave_window():
    from scipy.signal import resample_poly
    h = pystoi_np.resample_window_oct(10000, 16000)
    assert len(h) == 581 and abs(h[290] - h.max()) == 0
    x = np.random.default_rng(1).standard_normal(3000)
    assert np.allclose(pystoi_np.resample_oct(x, 10000, 16000), resample_poly(x, 5, 8, window=h / h.sum()))
    obm, cf, bins = pystoi_np.thirdoct()
    assert obm.shape == (15, 257) and bins[0][0] == 7 and bins[-1][1] == 219
    assert np.allclose(cf[:3], [150.0, 188.98815748, 238.11015779])


def test_siib_helpers_follow_intel_py():
    """framing / get_vad / stft are the helper copies of intel.py:16-54."""
    x = np.random.default_rng(2).standard_normal(3000)
    fr = pysiib_np.framing(x, 400, 200, 'hanning')
    assert fr.shape == (13, 400)                           # arange(0, 3000 - 400, 200)
    w = 0.5 - 0.5 * np.cos(2 * np.pi * np.arange(400) / 400)
    assert np.allclose(fr[3], x[600:1000] * w)
    vad = pysiib_np.get_vad(x, 400, 200, 'hanning', 40)
    assert vad.all()
    spec = pysiib_np.stft(x, 400, 200, 'hanning')
    assert spec.shape == (13, 201) and np.allclose(spec[2], np.fft.fft(fr[2])[:201])
    assert pysiib_np.n_filters() == 28


def test_siib_properties_and_wrapper_tiling():
    x, y, snr = make_pair(3, 52345)
    M, act = intel_np.siib_tiling_factor(x, 16000)
    assert M == int(np.floor(25 / (act / 80.0))) and M > 1          # intel.py:71-75
    s = intel_np.SIIB_Wrapper_raw_harvard(x, y, 16000)
    assert 5 < s < 500
    rng = np.random.default_rng(3)
    y2 = (y + 0.05 * rng.standard_normal(len(y))).astype(np.float32)
    assert intel_np.SIIB_Wrapper_raw_harvard(x, y2, 16000) < s       # more noise, less information
    st = {}
    x64 = np.hstack([x.astype(np.float64)] * M)
    clean = pysiib_np.SIIB(x64, x64, 16000, gauss=True, stages=st)
    # identical signals: every component has rho = 1 -> the production-noise ceiling
    ceil = 80 / 15 * 420 * (-0.5 * np.log2(1 - 0.75 ** 2))
    assert abs(clean - ceil) < 1e-6 * ceil
    with pytest.raises(ValueError):
        pysiib_np.SIIB(x.astype(np.float64), y.astype(np.float64), 16000, gauss=True)   # < 20 s of speech


def test_helpers_and_mappings_against_reference_generated_fixture():
    z = np.load(os.path.join(os.path.dirname(__file__), "golden", "intel_helpers.npz"))
    for i, L in ((0, 33536), (3, 52345), (7, 8000)):
        x, _, _ = make_pair(i, L)
        k = "p%d_%d" % (i, L)
        assert np.array_equal(pysiib_np.get_vad(x, 400, 200, 'hanning', 40), z[k + "/vad"])
        sp = pysiib_np.stft(x, 400, 200, 'hanning')
        assert sp.shape[0] == int(z[k + "/nframes"])
        assert np.allclose(sp[[0, 5, sp.shape[0] - 1]], z[k + "/stft_rows"], rtol=1e-12, atol=1e-14)
    g = z["map/grid"]
    assert np.allclose(intel_np.mapping_SIIB_harvard(g), z["map/siib"], rtol=0, atol=1e-15)
    assert np.allclose(intel_np.mapping_HASPI_harvard(g / 10), z["map/haspi"], rtol=0, atol=1e-15)
    assert np.allclose(intel_np.mapping_ESTOI_harvard(g / 150), z["map/estoi"], rtol=0, atol=1e-15)


def test_mappings_are_the_reference_logistics():
    assert abs(intel_np.mapping_SIIB_harvard(32.0) - 0.5) < 1e-15        # intel.py:102-106
    assert abs(intel_np.mapping_HASPI_harvard(2.8) - 0.5) < 1e-15        # intel.py:116-120
    assert abs(intel_np.mapping_ESTOI_harvard(0.25) - 0.5) < 1e-15       # intel.py:136-140
    assert abs(intel_np.mapping_ESTOI_harvard(1.0) - 1 / (1 + np.exp(-6.0))) < 1e-15


def test_siib_reference_value_on_periodic_tilings_is_rounding_noise():
    """Evidence for the engine's one deliberate deviation.

    intel.py:71-75 tiles a short utterance M times with np.hstack.  When its length is a multiple of the
    200-sample hop the tiled signal repeats its frames exactly, cov(X) (420 x 420) has rank ~ 76-99, and
    numpy's eigh returns arbitrary rounding-noise eigenvectors for the null space whose sample
    correlations add a few percent to the float64 score.  That surplus is not a property of the
    signals: a relative perturbation of 1e-13 of the inputs -- three orders of magnitude below float64's
    own rounding of the int16-derived samples' products -- removes it.  The engine returns the sum over
    the components with a numerically non-zero eigenvalue, i.e. the limit the reference formula itself
    converges to under any perturbation."""
    for i in range(4):
        x, y, _ = make_pair(i, 48000)                      # bench.py's workload: 3.0 s = 240 hops
        M, _ = intel_np.siib_tiling_factor(x, 16000)
        xt, yt = np.tile(x.astype(np.float64), M), np.tile(y.astype(np.float64), M)
        st = {}
        full = pysiib_np.SIIB(xt, yt, 16000, gauss=True, stages=st)
        nonnull = st["lam"] > 1e-9 * st["lam"].max()
        assert 60 < nonnull.sum() < 112                    # rank of the periodic tiling
        engine_value = 80.0 / 15 * float(np.sum(st["I_ch"][nonnull]))
        rng = np.random.default_rng(1000 + i)
        pert = pysiib_np.SIIB(xt * (1 + 1e-13 * rng.standard_normal(len(xt))),
                              yt * (1 + 1e-13 * rng.standard_normal(len(yt))), 16000, gauss=True)
        assert abs(full - engine_value) > 5e-3 * engine_value, (i, full, engine_value)     # 1 % .. 13 % of junk
        assert abs(pert - engine_value) <= 5e-3 * engine_value, (i, pert, engine_value)    # measured 0.05 % .. 0.18 %
        assert abs(pert - full) > 5e-3 * full                                               # float64 value moves under 1e-13
