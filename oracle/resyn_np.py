"""CPU restatement of the reference's resynthesis step (TEST INFRASTRUCTURE ONLY, see
``oracle/__init__.py``): ``interp_band_gain`` / ``Resyn`` / ``ISTFT`` of audio_util.py:60-115 and
the ``librosa.stft`` / ``librosa.istft`` (librosa 0.7.1, un-vendored) they call.  PARITY UNPINNED
for the librosa part: restated from its published algorithm (periodic Hann window, centred
frames with reflect padding; inverse by windowed overlap-add divided by the summed squared
window, trimmed by n_fft / 2 at both ends)."""
import numpy as np
from scipy.signal import get_window

GMTBAND = [0, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15, 16, 17, 18, 19, 20, 21, 22, 23, 24, 25, 26, 28, 30, 32, 34,
           36, 38, 41, 43, 46, 49, 52, 55, 58, 62, 66, 70, 74, 79, 83, 88, 93, 99, 105, 111, 117, 124, 131, 139, 147,
           156, 165, 174, 184, 195, 206, 218, 230, 243, 257]   # audio_util.py:23
NB_BANDS = 64


def interp_band_gain(bandE):
    """audio_util.py:98-115."""
    g = np.ones(257)
    for i in range(NB_BANDS - 1):
        band_size = GMTBAND[i + 1] - GMTBAND[i]
        for j in range(band_size):
            frac = float(j) / band_size
            g[GMTBAND[i] + j] = (1 - frac) * bandE[i] + frac * bandE[i + 1]
    g[0] = 1e-4
    g[1] = 1e-4
    g[256] = 1e-2
    return g


def stft(x, n_fft=512, hop=256):
    """librosa.stft(x, n_fft=512, hop_length=256, win_length=512) -> [257, T] (audio_util.py:52-57)."""
    w = get_window('hann', n_fft, fftbins=True)
    xp = np.pad(np.asarray(x, dtype=np.float64), n_fft // 2, mode='reflect')
    T = 1 + (len(xp) - n_fft) // hop
    fr = np.stack([xp[t * hop:t * hop + n_fft] * w for t in range(T)], axis=1)
    return np.fft.rfft(fr, axis=0)


def istft(X, n_fft=512, hop=256):
    """librosa.istft(X, hop_length=256, win_length=512) (audio_util.py:60-65)."""
    w = get_window('hann', n_fft, fftbins=True)
    T = X.shape[1]
    n = n_fft + hop * (T - 1)
    y = np.zeros(n)
    ss = np.zeros(n)
    for t in range(T):
        y[t * hop:t * hop + n_fft] += w * np.fft.irfft(X[:, t], n=n_fft)
        ss[t * hop:t * hop + n_fft] += w * w
    nz = ss > np.finfo(np.float32).tiny
    y[nz] /= ss[nz]
    return y[n_fft // 2: n - n_fft // 2]


def resyn(X, alpha2):
    """audio_util.py:84-96: X complex [257, T], alpha2 [T, 64]."""
    gain = np.stack([np.sqrt(interp_band_gain(alpha2[t])) for t in range(alpha2.shape[0])], axis=1)
    return istft(gain * X)
