"""CPU restatement of ``pystoi.stoi(x, y, fs_sig, extended=...)``.

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).  PARITY UNPINNED:
pystoi (mpariente/pystoi) is an un-vendored, un-pinned dependency of the
reference (README.md:14); its only call sites are intel.py:126 and :133
(``stoi(x, y, fs, extended=True)``).  Restated from the published algorithm
(Taal et al. 2011; Jensen & Taal 2016) as implemented by pystoi 0.3.x, the
release line contemporary with the reference (python 3.7 / 2021):

  1. resample both signals to 10 kHz with an Octave-compatible polyphase
     resampler (Kaiser-windowed sinc, 60 dB rejection, through
     ``scipy.signal.resample_poly``);
  2. remove silent frames: 256-sample Hann frames (MATLAB ``hanning``), hop
     128, frame starts ``range(0, len - 256, 128)``; keep frames whose energy
     is within 40 dB of the loudest frame OF x; overlap-add the kept frames;
  3. STFT (same window/hop, 512-point rfft), one-third-octave band matrix
     15 x 257 from 150 Hz, band magnitudes ``sqrt(OBM @ |X|^2)``;
  4. fewer than 30 frames -> RuntimeWarning and the sentinel 1e-5;
  5. sliding 30-frame segments; ESTOI: row then column mean/norm
     normalisation, score = mean over segments of sum(x_n * y_n) / 30.

``FRAME_RANGE_PLUS_ONE`` switches the frame enumeration to
``range(0, len - 256 + 1, 128)`` (pystoi >= 0.4); the default follows 0.3.x.
pystoi adds ``eps * N(0,1)`` before each normalisation to avoid 0/0 on
all-zero rows; ``jitter=False`` (default here) omits the 2.2e-16 noise so the
oracle is deterministic, ``jitter=True`` reproduces it from ``np.random``.
"""
import warnings

import numpy as np
from scipy.signal import resample_poly

FS = 10000
N_FRAME = 256
NFFT = 512
NUMBAND = 15
MINFREQ = 150
N = 30
BETA = -15.0
DYN_RANGE = 40
EPS = np.finfo("float").eps
FRAME_RANGE_PLUS_ONE = False


def thirdoct(fs=FS, nfft=NFFT, num_bands=NUMBAND, min_freq=MINFREQ):
    """One-third-octave band matrix [num_bands, nfft/2+1] and centre freqs."""
    f = np.linspace(0, fs, nfft + 1)[: nfft // 2 + 1]
    k = np.arange(num_bands, dtype=float)
    cf = np.power(2.0 ** (1.0 / 3), k) * min_freq
    lo = min_freq * np.power(2.0, (2 * k - 1) / 6)
    hi = min_freq * np.power(2.0, (2 * k + 1) / 6)
    obm = np.zeros((num_bands, len(f)))
    bins = np.zeros((num_bands, 2), dtype=int)
    for i in range(num_bands):
        a = int(np.argmin(np.square(f - lo[i])))
        b = int(np.argmin(np.square(f - hi[i])))
        obm[i, a:b] = 1
        bins[i] = (a, b)
    return obm, cf, bins


OBM, CF, OBM_BINS = thirdoct()


def hann_matlab(n):
    """MATLAB ``hanning(n)`` = ``np.hanning(n + 2)[1:-1]``."""
    return np.hanning(n + 2)[1:-1]


def resample_window_oct(p, q):
    """Octave ``resample`` anti-aliasing filter for the rational ratio p/q."""
    g = np.gcd(int(p), int(q))
    p, q = p / g, q / g
    log10_rejection = -3.0
    fc = 1.0 / (2 * max(p, q))
    roll = fc / 10
    rej_db = -20 * log10_rejection
    L = np.ceil((rej_db - 8) / (28.714 * roll))
    t = np.arange(-L, L + 1)
    ideal = 2 * p * fc * np.sinc(2 * fc * t)
    if 21 <= rej_db <= 50:
        beta = 0.5842 * (rej_db - 21) ** 0.4 + 0.07886 * (rej_db - 21)
    elif rej_db > 50:
        beta = 0.1102 * (rej_db - 8.7)
    else:
        beta = 0.0
    return np.kaiser(2 * L + 1, beta) * ideal


def resample_oct(x, p, q):
    h = resample_window_oct(p, q)
    return resample_poly(x, int(p), int(q), window=h / np.sum(h))


def _starts(n):
    return range(0, n - N_FRAME + (1 if FRAME_RANGE_PLUS_ONE else 0), N_FRAME // 2)


def remove_silent_frames(x, y, dyn_range=DYN_RANGE, framelen=N_FRAME, hop=N_FRAME // 2):
    w = hann_matlab(framelen)
    st = list(_starts(len(x)))
    xf = np.array([w * x[i:i + framelen] for i in st])
    yf = np.array([w * y[i:i + framelen] for i in st])
    en = 20 * np.log10(np.linalg.norm(xf, axis=1) + EPS)
    mask = (np.max(en) - dyn_range - en) < 0
    xf, yf = xf[mask], yf[mask]
    n_sil = (len(xf) - 1) * hop + framelen
    xs = np.zeros(n_sil)
    ys = np.zeros(n_sil)
    for i in range(xf.shape[0]):
        xs[i * hop:i * hop + framelen] += xf[i]
        ys[i * hop:i * hop + framelen] += yf[i]
    return xs, ys, mask


def stft(x, win_size=N_FRAME, fft_size=NFFT):
    w = hann_matlab(win_size)
    return np.array([np.fft.rfft(w * x[i:i + win_size], n=fft_size) for i in _starts(len(x))])


def _normalise(v, axis, jitter):
    if jitter:
        v = v + EPS * np.random.standard_normal(v.shape)
    v = v - np.mean(v, axis=axis, keepdims=True)
    return v / np.sqrt(np.sum(np.square(v), axis=axis, keepdims=True))


def row_col_normalize(seg, jitter=False):
    """[J, 15, 30]: normalise each band over time, then each frame over bands."""
    return _normalise(_normalise(seg, -1, jitter), 1, jitter)


def stoi(x, y, fs_sig, extended=False, jitter=False, stages=None):
    if x.shape != y.shape:
        raise Exception('x and y should have the same length,' +
                        'found {} and {}'.format(x.shape, y.shape))
    if fs_sig != FS:
        x = resample_oct(x, FS, fs_sig)
        y = resample_oct(y, FS, fs_sig)
    xs, ys, mask = remove_silent_frames(x, y)
    xspec = stft(xs).T
    yspec = stft(ys).T
    if stages is not None:
        stages.update(x10=x, y10=y, mask=mask, nframes=xspec.shape[-1] if xspec.ndim == 2 else 0)
    if xspec.ndim < 2 or xspec.shape[-1] < N:
        warnings.warn('Not enough STFT frames to compute intermediate '
                      'intelligibility measure after removing silent '
                      'frames. Returning 1e-5. Please check you wav files',
                      RuntimeWarning)
        return 1e-5
    xt = np.sqrt(OBM @ np.square(np.abs(xspec)))
    yt = np.sqrt(OBM @ np.square(np.abs(yspec)))
    nf = xt.shape[1]
    xseg = np.array([xt[:, m - N:m] for m in range(N, nf + 1)])
    yseg = np.array([yt[:, m - N:m] for m in range(N, nf + 1)])
    if stages is not None:
        stages.update(x_tob=xt, y_tob=yt)
    if extended:
        xn = row_col_normalize(xseg, jitter)
        yn = row_col_normalize(yseg, jitter)
        return np.sum(xn * yn / N) / xn.shape[0]
    # classic STOI (not on the NELE-GAN path; kept for completeness)
    norm = np.sqrt(np.sum(np.square(xseg), axis=2, keepdims=True) /
                   (np.sum(np.square(yseg), axis=2, keepdims=True) + EPS))
    yprim = np.minimum(yseg * norm, xseg * (1 + np.power(10.0, -BETA / 20)))
    yprim = yprim - np.mean(yprim, axis=2, keepdims=True)
    xm = xseg - np.mean(xseg, axis=2, keepdims=True)
    yprim = yprim / (np.linalg.norm(yprim, axis=2, keepdims=True) + EPS)
    xm = xm / (np.linalg.norm(xm, axis=2, keepdims=True) + EPS)
    return np.sum(yprim * xm) / (xseg.shape[0] * xseg.shape[1])
