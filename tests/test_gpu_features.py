"""Parity of the CUDA feature front-end (nele_features through the C ABI) with the reference-made
fixtures (tests/golden/features_ref.npz) and the CPU oracle (oracle/features_np.py).

Tolerances (floating point; stated here because BASELINE.json's north_star only names the metric scores):
magnitudes 1e-6 relative to the frame maximum (FP64 FFT rounded to float32, like the reference's
complex64), phases 2e-6 rad on bins that carry signal, band energies and IMCRA noise PSD 2e-5 relative
(FP64 recursion; float32 pow / sqrt at the end)."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CASES = ("p0_33536", "p3_52345", "p5_48000", "p7_8000")
POWER = 1 / 6


@pytest.fixture(scope="module")
def eng():
    from nele_gan_b200.engine import Engine
    e = Engine(0)
    yield e
    e.close()


@pytest.fixture(scope="module")
def g():
    return np.load(os.path.join(ROOT, "tests", "golden", "features_ref.npz"))


def rel(a, b):
    return np.max(np.abs(a.astype(np.float64) - b) / np.maximum(np.abs(b), 1e-30))


def check_mag_phase(mag, phase, gm, gp):
    assert mag.shape == gm.shape and phase.shape == gp.shape
    scale = np.maximum(gm.max(axis=0, keepdims=True), 1e-30)      # silent frames (pauses) are exactly zero
    assert np.max(np.abs(mag - gm) / scale) < 1e-6
    strong = gm > 1e-3 * scale
    d = np.angle(np.exp(1j * (phase.astype(np.float64) - gp)))
    assert np.max(np.abs(d[strong])) < 2e-6
    # DC and Nyquist bins are real: phase 0 or pi exactly as the reference returns them
    assert np.array_equal(np.abs(phase[[0, 256]]), np.abs(gp[[0, 256]]))
    # frames of digital silence are exactly zero, as in the reference
    silent = ~gm.any(axis=0)
    assert not mag[:, silent].any() and not phase[:, silent].any()


def test_speech_features_match_the_reference_fixture(eng, g):
    sigs = [g[c + "/speech"] for c in CASES]
    out = eng.features(sigs, power=POWER)
    for c, (band, mag, phase) in zip(CASES, out):
        check_mag_phase(mag, phase, g[c + "/speech_mag"], g[c + "/speech_phase"])
        assert band.shape == g[c + "/speech_band"].shape
        assert rel(band, g[c + "/speech_band"]) < 2e-5
    ms, launches = eng.last_timing()
    assert launches == 1


def test_noise_features_match_the_reference_fixture(eng, g):
    sigs = [g[c + "/noise"] for c in CASES] + [g["mix/signal"]]
    out = eng.features(sigs, power=POWER, noise=True, want_psd=True)
    for c, (band, mag, phase, psd) in zip(CASES, out):
        assert np.max(np.abs(mag - g[c + "/noise_mag"])) < 1e-6 * g[c + "/noise_mag"].max()
        assert rel(psd, g[c + "/noise_psd"]) < 2e-5
        assert rel(band, g[c + "/noise_band"]) < 2e-5
    assert rel(out[-1][3], g["mix/psd"]) < 2e-5
    assert rel(out[-1][0], g["mix/band"]) < 2e-5
    assert eng.last_timing()[1] == 2
    raw = eng.features(sigs[:2], power=POWER, noise=True, normalization=False)
    assert rel(raw[0][0], g[CASES[0] + "/noise_band_raw"]) < 2e-5


def test_ragged_batch_against_the_oracle(eng):
    """Lengths around the frame and tile boundaries (T = 1 + L // 256; CTAs take 8 frames, IMCRA stages 32)."""
    from oracle import features_np
    lens = [257, 511, 512, 2047, 2048, 2049, 8191, 8192, 8193 + 256, 16000, 30001]
    # white noise with a slow level ramp (make_pair can fall into one of its pauses at these lengths)
    sigs = [(np.random.default_rng(100 + i).standard_normal(L) * np.linspace(0.01, 0.1, L)).astype(np.float32)
            for i, L in enumerate(lens)]
    sp = eng.features(sigs, power=POWER)
    no = eng.features(sigs, power=POWER, noise=True, want_psd=True)
    for x, (band, mag, phase), (nband, nmag, nphase, psd) in zip(sigs, sp, no):
        ob, om, op = features_np.sp_and_phase_speech(x, POWER)
        assert mag.shape == om.shape == (257, 1 + len(x) // 256)
        check_mag_phase(mag, phase, om, op)
        assert rel(band, ob) < 2e-5
        assert np.array_equal(nmag, mag) and np.array_equal(nphase, phase)
        F = features_np.stft(x)
        assert rel(psd, features_np.imcra_noise_psd(F)) < 2e-5
        assert rel(nband, features_np.sp_and_phase_noise(x, POWER)[0]) < 2e-5


def test_dropin_functions_and_errors(eng, g):
    from nele_gan_b200 import features as F
    from nele_gan_b200.engine import NeleError
    c = CASES[0]
    band, mag, phase = F.Sp_and_phase_Speech(g[c + "/speech"], POWER)
    assert rel(band, g[c + "/speech_band"]) < 2e-5 and mag.shape == phase.shape == g[c + "/speech_mag"].shape
    band, mag, phase = F.Sp_and_phase_Noise(g[c + "/noise"], POWER)
    assert rel(band, g[c + "/noise_band"]) < 2e-5
    assert rel(F.noise_psd(g[c + "/noise"]), g[c + "/noise_psd"]) < 2e-5
    with pytest.raises(NeleError):     # librosa raises for signals not longer than n_fft / 2 (reflect padding)
        eng.features([np.zeros(256, np.float32)])
    assert eng.features([]) == []
    # silence: zero spectrum, zero phase, zero band energies (0 ** (1 / 6) = 0)
    band, mag, phase = eng.features([np.zeros(4000, np.float32)], power=POWER)[0]
    assert not band.any() and not mag.any() and not phase.any()


def test_device_tensors_full_size_properties():
    """BASELINE configs[2] size on the device: 4096 x 3 s.  Size-independent properties: identical rows give
    identical features (batch independence), scaling the waveform by 2 scales magnitudes by 2 and band energies
    by 4 ** power exactly in binary floating point, phases unchanged."""
    torch = pytest.importorskip("torch")
    from nele_gan_b200 import features as F
    from nele_gan_b200.synth import make_pair
    base = np.stack([make_pair(200 + i, 48000)[1] for i in range(8)])
    wav = torch.from_numpy(np.tile(base, (512, 1))).cuda()
    band, mag, phase, foff, frames = F.features_tensors(wav, power=1.0)
    T = 188
    assert frames.tolist() == [T] * 4096 and band.shape == (4096 * T, 64)
    b = band.view(512, 8, T, 64)
    m = mag.view(512, 8, 257, T)
    assert torch.equal(b[0], b[511]) and torch.equal(m[0], m[300])
    band2, mag2, phase2, _, _ = F.features_tensors(2 * wav[:64], power=1.0)
    assert torch.equal(mag2, 2 * mag[:257 * 64 * T]) and torch.equal(band2, 4 * band[:64 * T])
    assert torch.equal(phase2, phase[:257 * 64 * T])
    nband, nmag, _, _, _ = F.features_tensors(wav, power=1.0, noise=True, want_phase=False)
    nb = nband.view(512, 8, T, 64)
    assert torch.equal(nmag, mag) and torch.equal(nb[0], nb[17]) and bool(torch.isfinite(nband).all())
    # IMCRA is scale-equivariant: PSD (and so the band energies) scale by 4 -- up to the 1e-6 floor of the
    # initial estimate (noise_est/imcra.py:518), which only enters the first frame's a-posteriori SNR
    nband2, _, _, _, _ = F.features_tensors(2 * wav[:8], power=1.0, noise=True, want_phase=False)
    assert torch.allclose(nband2, 4 * nband[:8 * T], rtol=1e-3, atol=0)
