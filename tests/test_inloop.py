"""Host logic of the in-loop tensor boundary (nele_gan_b200/inloop.py) against the numpy
restatement of audio_util.py:60-115 (oracle/resyn_np.py); runs on the CPU with torch."""
import numpy as np
import pytest

torch = pytest.importorskip("torch")


def _round(n=3, L=256 * 40, seed=0):
    from nele_gan_b200.synth import make_pair
    rng = np.random.default_rng(seed)
    clean = np.stack([make_pair(60 + i, L)[0] for i in range(n)])
    T = 1 + L // 256
    alpha2 = rng.uniform(0.2, 3.0, size=(n, T, 64))
    return clean, alpha2


def test_band_gain_matrix_is_interp_band_gain():
    from nele_gan_b200 import inloop
    from oracle import resyn_np
    W, fixed = inloop.band_gain_matrix()
    e = np.random.default_rng(1).uniform(0.1, 2.0, 64)
    g = np.where(np.isnan(fixed), W @ e, fixed)
    assert np.allclose(g, resyn_np.interp_band_gain(e), rtol=0, atol=1e-15)


def test_stft_and_resyn_match_the_numpy_restatement():
    from nele_gan_b200 import inloop
    from oracle import resyn_np
    clean, alpha2 = _round()
    X = inloop.stft(torch.from_numpy(clean).double())
    for i in range(len(clean)):
        Xo = resyn_np.stft(clean[i])
        assert X[i].shape == Xo.shape
        assert np.abs(X[i].numpy() - Xo).max() < 1e-10
    y = inloop.resyn(X, torch.from_numpy(alpha2))
    for i in range(len(clean)):
        yo = resyn_np.resyn(resyn_np.stft(clean[i]), alpha2[i])
        assert y[i].shape[0] == len(yo) == 256 * (alpha2.shape[1] - 1)
        assert np.abs(y[i].numpy() - yo).max() < 1e-10


def test_unit_gains_reproduce_the_clean_signal_away_from_the_forced_bins():
    from nele_gan_b200 import inloop
    clean, alpha2 = _round(n=1)
    X = inloop.stft(torch.from_numpy(clean).double())
    y = inloop.resyn(X, torch.ones_like(torch.from_numpy(alpha2))).numpy()[0]
    # bins 0, 1 and 256 are attenuated (audio_util.py:111-113): the rest of the band passes unchanged
    d = y - clean[0][: len(y)]
    assert np.sqrt(np.mean(d ** 2)) < 0.2 * np.sqrt(np.mean(clean[0] ** 2))
