"""CPU restatement of the reference metric wrappers (intel.py:57-140) and of
the per-file labelling functions (audio_util.py:120-203, :267-321).

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).  ``intel.py`` itself
cannot be imported here (it needs pysiib / pystoi / pypesq, and asks scipy for
a 'hanning' window that scipy >= 1.13 no longer knows), so its ~60 lines on the
hot path are restated over the oracle metrics.
"""
import os

import numpy as np

from oracle import haspi_np, pysiib_np, pystoi_np


def logistic(x, a, b):
    return 1 / (1 + np.exp(a * (x - b)))


def mapping_SIIB_harvard(x):     # intel.py:102-106
    return logistic(x, -0.06, 32)


def mapping_HASPI_harvard(x):    # intel.py:116-120
    return logistic(x, -0.95, 2.8)


def mapping_ESTOI_harvard(x):    # intel.py:136-140
    return logistic(x, -8.0, 0.25)


def _trim(x, y):
    m = min(len(x), len(y))
    return x[:m], y[:m]


def siib_tiling_factor(x, fs):
    """intel.py:62-75: active seconds after VAD -> number of copies M (1 = no tiling)."""
    R = 1 / 200 * fs
    active = int(np.sum(pysiib_np.get_vad(x, 400, 200, 'hanning', 40)))
    if active / R < 20:
        return int(np.floor(25 / (active / R))), active
    return 1, active


def SIIB_Wrapper_raw_harvard(x, y, fs, gauss=True):   # intel.py:57-77
    x, y = _trim(x, y)
    M, _ = siib_tiling_factor(x, fs)
    if M != 1:
        x = np.hstack([x] * M)
        y = np.hstack([y] * M)
    return pysiib_np.SIIB(x, y, fs, gauss=gauss)


def SIIB_Wrapper_harvard(x, y, fs):                   # intel.py:79-100
    return mapping_SIIB_harvard(SIIB_Wrapper_raw_harvard(x, y, fs))


def HASPI_Wrapper_raw_harvard(x, y, fs, noise="numpy"):   # intel.py:112-114
    return haspi_np.haspi_v2(x, fs, y, fs, noise=noise)[0]


def HASPI_Wrapper_harvard(x, y, fs, noise="numpy"):       # intel.py:108-110
    return mapping_HASPI_harvard(HASPI_Wrapper_raw_harvard(x, y, fs, noise=noise))


def ESTOI_Wrapper_raw_harvard(x, y, fs):                  # intel.py:122-127
    x, y = _trim(x, y)
    return pystoi_np.stoi(x, y, fs, extended=True)


def ESTOI_Wrapper_harvard(x, y, fs):                      # intel.py:129-134
    return mapping_ESTOI_harvard(ESTOI_Wrapper_raw_harvard(x, y, fs))


def score_pair(x, y, fs=16000, norm=True, noise=None):
    """All three labels of one (clean, degraded) pair: [SIIB, HASPI, ESTOI]
    in the order train_nele.py:320-322 computes them."""
    if norm:
        return np.array([SIIB_Wrapper_harvard(x, y, fs), HASPI_Wrapper_harvard(x, y, fs, noise),
                         ESTOI_Wrapper_harvard(x, y, fs)])
    return np.array([SIIB_Wrapper_raw_harvard(x, y, fs), HASPI_Wrapper_raw_harvard(x, y, fs, noise),
                     ESTOI_Wrapper_raw_harvard(x, y, fs)])


# ----------------------------------------------------------------------------
# per-file layer (audio_util.py:120-203, 267-321)
# ----------------------------------------------------------------------------
def load16k(path):
    """``librosa.load(path, sr=16000)`` for the 16 kHz PCM-16 files the
    reference asserts on (audio_util.py:131,159,187)."""
    from scipy.io import wavfile
    fs, x = wavfile.read(path)
    assert fs == 16000
    return x.astype(np.float32) / 32768.0


def wave_name(enhanced_file, drc=False):
    """audio_util.py:121-126 (and :268-269 for the _DRC form)."""
    f = enhanced_file.split('/')[-1]
    if drc:
        return f
    return (f.split('@')[0] if '@' in f else f[:-4]) + '.wav'


def read_pair(clean_root, noise_root, enhanced_file, drc=False):
    """audio_util.py:130-137: (clean, enhanced + noise) trimmed to min length."""
    name = wave_name(enhanced_file, drc)
    clean = load16k(os.path.join(clean_root, name) if not clean_root.endswith('/') else clean_root + name)
    noise = load16k(os.path.join(noise_root, name) if not noise_root.endswith('/') else noise_root + name)
    enh = load16k(enhanced_file)
    m = min(len(clean), len(enh))
    return clean[:m], enh[:m] + noise[:m]
