"""The CPU restatement of the feature front-end (oracle/features_np.py) against fixtures made by the
reference's own code (tests/golden/make_golden_features.py: ``compute_band_E``, ``Sp_and_phase_*`` cut
from the unmodified audio_util.py, the unmodified noise_est/imcra.py).  Only ``librosa.stft`` is a
restatement there (un-vendored): that step is checked against torch.stft instead."""
import os

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CASES = ("p0_33536", "p3_52345", "p5_48000", "p7_8000")
POWER = 1 / 6


@pytest.fixture(scope="module")
def feat_golden():
    return np.load(os.path.join(ROOT, "tests", "golden", "features_ref.npz"))


def rel(a, b):
    return np.max(np.abs(a.astype(np.float64) - b) / np.maximum(np.abs(b), 1e-30))


@pytest.mark.parametrize("case", CASES)
def test_speech_features_match_the_reference(feat_golden, case):
    from oracle import features_np
    g = feat_golden
    band, mag, phase = features_np.sp_and_phase_speech(g[case + "/speech"], POWER)
    assert band.shape == g[case + "/speech_band"].shape and band.dtype == np.float32
    assert np.array_equal(mag, g[case + "/speech_mag"])
    assert np.array_equal(phase, g[case + "/speech_phase"])
    # the reference accumulates bin by bin (float32 products, float64 sum); the oracle uses one matrix product
    assert rel(band, g[case + "/speech_band"]) < 1e-6


@pytest.mark.parametrize("case", CASES)
def test_imcra_matches_the_unmodified_reference(feat_golden, case):
    from oracle import features_np
    g = feat_golden
    F = features_np.stft(g[case + "/noise"])
    psd = features_np.imcra_noise_psd(F)
    assert psd.shape == g[case + "/noise_psd"].shape
    assert rel(psd, g[case + "/noise_psd"]) < 2.5e-7          # one float32 ulp
    band, _, _ = features_np.sp_and_phase_noise(g[case + "/noise"], POWER)
    assert rel(band, g[case + "/noise_band"]) < 1e-6
    raw, _, _ = features_np.sp_and_phase_noise(g[case + "/noise"], POWER, Normalization=False)
    assert rel(raw, g[case + "/noise_band_raw"]) < 1e-6


def test_imcra_on_a_speech_plus_noise_mixture(feat_golden):
    """Speech presence drives the a-priori-absence / posterior-probability branches (noise_est/imcra.py:429-447)."""
    from oracle import features_np
    g = feat_golden
    psd = features_np.imcra_noise_psd(features_np.stft(g["mix/signal"]))
    assert rel(psd, g["mix/psd"]) < 2.5e-7
    assert rel(features_np.sp_and_phase_noise(g["mix/signal"], POWER)[0], g["mix/band"]) < 1e-6


def test_restated_librosa_stft_agrees_with_torch():
    torch = pytest.importorskip("torch")
    from nele_gan_b200.synth import make_pair
    from oracle import features_np
    x = make_pair(21, 20000)[0]
    F = features_np.stft(x)
    win = torch.hann_window(512, periodic=True, dtype=torch.float64)
    Ft = torch.stft(torch.from_numpy(x).double(), 512, hop_length=256, win_length=512, window=win, center=True,
                    pad_mode="reflect", return_complex=True).numpy()
    assert F.shape == Ft.shape == (257, 1 + 20000 // 256) and F.dtype == np.complex64
    assert np.max(np.abs(F - Ft)) < 1e-6 * np.max(np.abs(Ft))


def test_band_matrix_is_compute_band_E():
    """The 64 x 257 map reproduces the triple loop of audio_util.py:30-50 (restated here literally, small case)."""
    from oracle import features_np
    from oracle.resyn_np import GMTBAND
    X = np.abs(np.random.default_rng(3).standard_normal((3, 257))).astype(np.float32)
    out = np.zeros((3, 64), dtype=np.float32)
    for t in range(3):
        s = np.zeros(64)
        for i in range(63):
            size = GMTBAND[i + 1] - GMTBAND[i]
            for j in range(size):
                frac = float(j) / size
                tmp = X[t, GMTBAND[i] + j] ** 2
                s[i] += (1 - frac) * tmp
                s[i + 1] += frac * tmp
        out[t] = s
    assert rel(features_np.compute_band_E(X), out) < 1e-6


def test_host_mirror_rejects_cpu_tensors_and_has_the_reference_names():
    """nele_gan_b200/features.py: same names as audio_util.py:422-457, no CPU path."""
    torch = pytest.importorskip("torch")
    from nele_gan_b200 import features as F
    for name in ("Sp_and_phase_Speech", "Sp_and_phase_Noise", "speech_features", "noise_features", "features_tensors"):
        assert callable(getattr(F, name))
    assert F.power_law == 1 / 6                                   # dataloader.py:14
    with pytest.raises(ValueError):
        F.features_tensors(torch.zeros(2, 4000))                  # CPU tensor: the engine has no CPU path
    import inspect
    src = inspect.getsource(F)
    assert "oracle" not in src.replace("``oracle/``", "")         # the product never imports the oracle


def test_frame_bookkeeping_matches_librosa_centred_framing():
    """T = 1 + L // 256 (centred frames, hop 256): what nele_feature_frames and the Python binding assume."""
    from oracle import features_np
    for L in (257, 511, 512, 513, 4000, 48000):
        assert features_np.stft(np.zeros(L, np.float32)).shape == (257, 1 + L // 256)
