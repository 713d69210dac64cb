"""Minimal stand-in for ``librosa`` so the UNMODIFIED reference module
``pyHASPI/pyhaspi2.py`` (``import librosa`` at :23, ``librosa.resample`` at
:815, ``librosa.load`` at :1254) can be imported in the build container.

TEST INFRASTRUCTURE ONLY.  Used by ``tests/golden/make_golden.py`` (and by
``oracle/reference_bridge.py`` when ``/root/reference`` exists); never by the
product.  ``resample`` is the resampy ``kaiser_best`` restatement in
``oracle/resampy_kaiser.py``.
"""
import numpy as np

from oracle.resampy_kaiser import librosa_resample


def resample(y, orig_sr=None, target_sr=None, *args, **kwargs):
    if args:  # librosa 0.7.1 positional form resample(y, orig_sr, target_sr)
        raise TypeError("shim supports keyword or 3 positional args only")
    return librosa_resample(np.asarray(y), orig_sr, target_sr)


def load(path, sr=22050, mono=True, dtype=np.float32):
    """librosa.load for PCM WAV: int16 -> float32 / 32768, optional resample."""
    from scipy.io import wavfile

    fs, x = wavfile.read(path)
    if x.dtype == np.int16:
        x = x.astype(np.float32) / 32768.0
    elif x.dtype == np.int32:
        x = x.astype(np.float32) / 2147483648.0
    else:
        x = x.astype(np.float32)
    if x.ndim > 1 and mono:
        x = x.mean(axis=1)
    if sr is not None and sr != fs:
        x = librosa_resample(x, fs, sr)
        fs = sr
    return x.astype(dtype), fs
