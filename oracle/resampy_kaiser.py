"""Restatement of resampy ``kaiser_best`` band-limited sinc interpolation.

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).  PARITY UNPINNED: the
real ``resampy`` is not in the reference tree nor in this image.

What the reference calls: ``librosa.resample(x, orig_sr, target_sr)``
(pyHASPI/pyhaspi2.py:815) -> librosa 0.7.1 ``core.audio.resample`` ->
``resampy.resample(y, orig_sr, target_sr, filter='kaiser_best')`` followed by
``util.fix_length(y_hat, ceil(n * ratio))`` and a cast back to the input dtype.

Published algorithm (resampy 0.2.x):
  * filter table: half of a Kaiser-windowed sinc, ``num_zeros = 64`` zero
    crossings, ``2**9 = 512`` table entries per crossing,
    ``rolloff = 0.9475937167399596``, ``beta = 14.769656459379492``;
  * per output sample t: position ``tau = t / ratio`` (accumulated by repeated
    addition of ``1/ratio`` in float64), left wing over x[n - i], right wing
    over x[n + 1 + k], filter value linearly interpolated between table
    entries, table scaled by ``ratio`` when down-sampling;
  * output length ``int(n * ratio)``, accumulated in the input dtype.
"""
import numpy as np

try:  # numba is in the image; keep a pure-python path so import never fails
    from numba import njit
except Exception:  # pragma: no cover
    def njit(*a, **k):
        def deco(f):
            return f
        return deco if not (a and callable(a[0])) else a[0]

NUM_ZEROS = 64
PRECISION = 9
ROLLOFF = 0.9475937167399596
BETA = 14.769656459379492

_cache = {}


def kaiser_best_table():
    """(half window [32769] float64, entries per zero crossing)."""
    if "tab" not in _cache:
        per_zero = 2 ** PRECISION
        n = per_zero * NUM_ZEROS
        t = np.linspace(0, NUM_ZEROS, num=n + 1, endpoint=True)
        sinc_half = ROLLOFF * np.sinc(ROLLOFF * t)
        taper = np.kaiser(2 * n + 1, BETA)[n:]
        _cache["tab"] = (taper * sinc_half, per_zero)
    return _cache["tab"]


@njit(cache=True)
def _interp_loop(x, y, ratio, win, dwin, per_zero):
    scale = min(1.0, ratio)
    step = 1.0 / ratio
    stride = int(scale * per_zero)
    pos = 0.0
    nwin = win.shape[0]
    n_in = x.shape[0]
    for t in range(y.shape[0]):
        n = int(pos)
        frac = scale * (pos - n)
        f = frac * per_zero
        off = int(f)
        eta = f - off
        acc = y[t]
        cnt = min(n + 1, (nwin - off) // stride)
        for i in range(cnt):
            w = win[off + i * stride] + eta * dwin[off + i * stride]
            acc += w * x[n - i]
        frac = scale - frac
        f = frac * per_zero
        off = int(f)
        eta = f - off
        cnt = min(n_in - n - 1, (nwin - off) // stride)
        for k in range(cnt):
            w = win[off + k * stride] + eta * dwin[off + k * stride]
            acc += w * x[n + k + 1]
        y[t] = acc
        pos += step


def resampy_resample(x, sr_orig, sr_new):
    """resampy.resample(x, sr_orig, sr_new, filter='kaiser_best') for 1-D x."""
    x = np.ascontiguousarray(x)
    ratio = float(sr_new) / float(sr_orig)
    n_out = int(x.shape[0] * ratio)
    y = np.zeros(n_out, dtype=x.dtype)
    win, per_zero = kaiser_best_table()
    win = win.copy()
    if ratio < 1:
        win *= ratio
    dwin = np.zeros_like(win)
    dwin[:-1] = np.diff(win)
    _interp_loop(x, y, ratio, win, dwin, per_zero)
    return y


def librosa_resample(y, orig_sr, target_sr):
    """librosa 0.7.1 ``resample(y, orig_sr, target_sr)`` with its defaults
    (res_type='kaiser_best', fix=True, scale=False)."""
    if orig_sr == target_sr:
        return y
    ratio = float(target_sr) / orig_sr
    n_samples = int(np.ceil(y.shape[-1] * ratio))
    y_hat = resampy_resample(y, orig_sr, target_sr)
    if y_hat.shape[0] < n_samples:  # util.fix_length: zero pad
        y_hat = np.concatenate([y_hat, np.zeros(n_samples - y_hat.shape[0], y_hat.dtype)])
    else:
        y_hat = y_hat[:n_samples]
    return np.ascontiguousarray(y_hat, dtype=y.dtype)
