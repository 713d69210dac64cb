"""CPU restatement of ``pysiib.SIIB(x, y, fs, gauss=...)``.

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).  PARITY UNPINNED:
pysiib (kamo-naoyuki/pySIIB) is an un-vendored, un-pinned dependency of the
reference (README.md:13).  Its only call sites are intel.py:77 and :100, both
``SIIB(x, y, fs, gauss=True)`` (SIIB^Gauss).  The in-tree anchors are the
helper copies the reference keeps in intel.py:16-54 (``framing``, ``get_vad``,
``stft``), which this file follows line for line in meaning; everything else
is restated from the published algorithm (Van Kuyk, Kleijn & Hendriks, "An
instrumental intelligibility metric based on information theory", IEEE SPL
2018, and "An evaluation of intrusive instrumental intelligibility metrics",
IEEE/ACM TASLP 2018 for the Gaussian variant):

  1. 16 kHz, remove means, 400/200 periodic-Hann STFT power of x and y;
  2. VAD on x: frame power within 40 dB of the 99.9th percentile; the same
     frames are kept for y; at least 20 s of active speech required;
  3. 28 ERB-spaced gammatone magnitude responses (100 Hz .. 6.5 kHz, 4th
     order, Holdsworth normalisation) applied to the power spectra, log;
  4. forward temporal masking over floor(0.2 s * 80 frames/s) = 16 frames:
     each frame holds the following frames above a floor that decays
     linearly in log-time from the frame's own level to the band minimum
     (Rhebergen et al. 2006);
  5. remove per-band means, stack K = 15 consecutive frames (420 x (F-14)),
     KLT with the eigenvectors of cov(X), same transform on Y;
  6. per-component information, summed assuming independence:
       gauss=True : -1/2 log2(1 - (rho_p * rho_j)^2), rho_p = 0.75,
       gauss=False: min(KSG k-NN estimate, -1/2 log2(1 - rho_p^2));
     SIIB = max(0, (R / K) * sum_j I_j)  [bits/s].
"""
import math

import numpy as np
from scipy.fftpack import fft
from scipy.signal import get_window

EPS = np.finfo(np.float64).eps
FS = 16000
RHO_P = 0.75
K_STACK = 15
N_BANDS_RANGE = (100.0, 6500.0)


def _window(name, n):
    # intel.py:34 asks scipy for 'hanning'; scipy >= 1.13 only knows 'hann'
    # (same periodic window).
    return get_window('hann' if name == 'hanning' else name, n)


def framing(x, window_length, window_shift, window):
    """intel.py:16-35."""
    slen = x.shape[-1]
    if slen < window_length + 1:
        x = np.pad(x, [(0, window_length + 1 - slen)], mode='constant')
    nrow = x.shape[-1] - window_length
    idx = np.arange(0, nrow, window_shift)[:, None] + np.arange(window_length)[None, :]
    return x[idx] * _window(window, window_length)[None, :]


def get_vad(x, window_length, window_shift, window, delta_db):
    """intel.py:37-50."""
    fr = framing(x, window_length, window_shift, window)
    x_db = 10 * np.log10((fr ** 2).mean(axis=1) + EPS)
    ind = int(round(len(x_db) * 0.999) - 1)
    max_x = np.partition(x_db, ind)[ind]
    return x_db > (max_x - delta_db)


def stft(x, window_length, window_shift, window):
    """intel.py:52-54."""
    fr = framing(x, window_length, window_shift, window)
    return fft(fr, n=window_length, axis=-1)[:, :window_length // 2 + 1]


def n_filters(mn=N_BANDS_RANGE[0], mx=N_BANDS_RANGE[1]):
    return int(round(21.4 * np.log10(1 + 0.00437 * mx) - 21.4 * np.log10(1 + 0.00437 * mn)))


def gammatone(fs, n_fft, num_bands, cf_min, cf_max):
    """Gammatone magnitude responses [num_bands, n_fft/2+1], peak-normalised."""
    erb = 21.4 * np.log10(4.37 * (np.array([cf_min, cf_max]) / 1000) + 1)
    cf = (10 ** (np.linspace(erb[0], erb[1], num_bands) / 21.4) - 1) / 4.37 * 1000
    order = 4
    a = math.factorial(order - 1) ** 2 / (math.pi * math.factorial(2 * order - 2) * 2.0 ** -(2 * order - 2))
    b = a * 24.7 * (4.37 * cf / 1000 + 1)
    f = np.linspace(0, fs, n_fft + 1)[: n_fft // 2 + 1]
    A = 1.0 / (b[:, None] ** 2 + (f[None, :] - cf[:, None]) ** 2) ** (order / 2)
    return A / A.max(axis=1, keepdims=True), cf


def forward_masking(X, Tf):
    """In place, frame by frame: frame i keeps frames i..i+Tf-1 above
    X[:,i] - log(d+1)/log(Tf) * (X[:,i] - band minimum), d = 0..Tf-1."""
    J, F = X.shape
    floor = X.min(axis=1)
    decay = np.log(np.arange(1, Tf + 1)) / np.log(Tf)
    for i in range(F):
        n = min(Tf, F - i)
        lvl = X[:, i]
        m = lvl[:, None] - decay[None, :n] * (lvl - floor)[:, None]
        X[:, i:i + n] = np.maximum(X[:, i:i + n], m)
    return X


def stack_frames(X, K):
    """[J, F] -> [J*K, F-K+1]; row k*J + j holds band j delayed by k frames."""
    J, F = X.shape
    return np.concatenate([X[:, k:F - K + 1 + k] for k in range(K)], axis=0)


def _digamma(n):
    from scipy.special import digamma
    return digamma(n)


def mi_ksg(x, y, k):
    """Kraskov-Stoegbauer-Grassberger estimator (algorithm 1, max-norm), bits."""
    from scipy.spatial import cKDTree
    n = len(x)
    pts = np.stack([x, y], axis=1)
    d, _ = cKDTree(pts).query(pts, k=k + 1, p=np.inf)
    eps = d[:, -1]
    xs = np.sort(x)
    ys = np.sort(y)
    nx = np.searchsorted(xs, x + eps, 'left') - np.searchsorted(xs, x - eps, 'right') - 1
    ny = np.searchsorted(ys, y + eps, 'left') - np.searchsorted(ys, y - eps, 'right') - 1
    nx = np.maximum(nx, 0)
    ny = np.maximum(ny, 0)
    nats = _digamma(k) + _digamma(n) - np.mean(_digamma(nx + 1) + _digamma(ny + 1))
    return nats / np.log(2)


def siib_features(x, y, window_length=400, window_shift=200, window='hanning', delta_dB=40.0, stages=None):
    """Steps 1-5 up to (not including) the KLT: returns stacked X, Y and R."""
    R = 1 / window_shift * FS
    x = x - np.mean(x)
    y = y - np.mean(y)
    xh = stft(x, window_length, window_shift, window).T
    yh = stft(y, window_length, window_shift, window).T
    xh = xh.real ** 2 + xh.imag ** 2
    yh = yh.real ** 2 + yh.imag ** 2
    vad = get_vad(x, window_length, window_shift, window, delta_dB)
    xh, yh = xh[:, vad], yh[:, vad]
    if xh.shape[1] / R < 20:
        raise ValueError('stimuli must have at least 20 seconds of speech')
    J = n_filters()
    G, _ = gammatone(FS, window_length, J, *N_BANDS_RANGE)
    X = np.log(G ** 2 @ xh + EPS)
    Y = np.log(G ** 2 @ yh + EPS)
    Tf = int(np.floor(0.2 * R))
    X = forward_masking(X, Tf)
    Y = forward_masking(Y, Tf)
    X = X - X.mean(axis=1, keepdims=True)
    Y = Y - Y.mean(axis=1, keepdims=True)
    if stages is not None:
        stages.update(vad=vad, X=X, Y=Y)
    return stack_frames(X, K_STACK), stack_frames(Y, K_STACK), R


def SIIB(x, y, fs_signal, gauss=False, use_MI_Kraskov=True, window_length=400,
         window_shift=200, window='hanning', delta_dB=40.0, stages=None):
    x = np.asarray(x, dtype=np.float64)
    y = np.asarray(y, dtype=np.float64)
    if x.ndim != 1 or y.ndim != 1:
        raise ValueError('x and y must be 1-D')
    if len(x) != len(y):
        raise ValueError('x and y should have the same length')
    if fs_signal != FS:
        from scipy.signal import resample_poly
        g = math.gcd(int(FS), int(fs_signal))
        x = resample_poly(x, FS // g, int(fs_signal) // g)
        y = resample_poly(y, FS // g, int(fs_signal) // g)
    Xs, Ys, R = siib_features(x, y, window_length, window_shift, window, delta_dB, stages=stages)
    lam, U = np.linalg.eigh(np.cov(Xs))
    Xk = U.T @ Xs
    Yk = U.T @ Ys
    if gauss:
        xc = Xk - Xk.mean(axis=1, keepdims=True)
        yc = Yk - Yk.mean(axis=1, keepdims=True)
        den = np.sqrt(np.sum(xc ** 2, axis=1) * np.sum(yc ** 2, axis=1))
        rho = np.sum(xc * yc, axis=1) / np.maximum(den, np.finfo(float).tiny)
        I_ch = -0.5 * np.log2(1 - (RHO_P * rho) ** 2)
    else:
        k = max(2, int(math.ceil(0.01 * Xk.shape[1])))
        cap = -0.5 * np.log2(1 - RHO_P ** 2)
        I_ch = np.array([min(mi_ksg(Xk[j], Yk[j], k), cap) for j in range(Xk.shape[0])])
    if stages is not None:
        stages.update(nf=Xs.shape[1], lam=lam, I_ch=I_ch, rho=rho if gauss else None)
    return max(0.0, R / K_STACK * float(np.sum(I_ch)))
