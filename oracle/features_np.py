"""CPU restatement of the reference's feature front-end (TEST INFRASTRUCTURE ONLY, see
``oracle/__init__.py``): what ``dataloader.py:30-84`` computes for every utterance it feeds to the
generator and the discriminator.

  * ``stft``                 ``librosa.stft(x, n_fft=512, hop_length=256, win_length=512)``
                             (audio_util.py:52-57).  librosa 0.7.1 is un-vendored: PARITY UNPINNED for
                             this step, restated from its published algorithm (periodic Hann window,
                             centred frames with reflect padding, float64 FFT stored as complex64).
  * ``compute_band_E``       audio_util.py:30-50, the 257 -> 64 triangular band energies.
  * ``imcra_noise_psd``      ``NoisePSD`` = ``imcra_est(nfft=512).estimate`` (audio_util.py:117-122,
                             noise_est/imcra.py:487-577 driving the ``imcra`` class, :163-484):
                             decision-directed a-priori SNR + IMCRA minima-controlled noise tracking.
  * ``sp_and_phase_speech``  audio_util.py:422-437
  * ``sp_and_phase_noise``   audio_util.py:439-457

``compute_band_E`` and the IMCRA recursion are pinned against the unmodified reference code
(``tests/golden/features_ref.npz``, made by ``tests/golden/make_golden_features.py``).
"""
import numpy as np

from .resyn_np import GMTBAND, NB_BANDS

N_FFT, HOP, FREQ = 512, 256, 257


def stft(x):
    """[N] float32 -> complex64 [257, 1 + N // 256]."""
    x = np.asarray(x)
    n = np.arange(N_FFT)
    w = 0.5 - 0.5 * np.cos(2.0 * np.pi * n / N_FFT)          # get_window('hann', 512, fftbins=True)
    xp = np.pad(x, N_FFT // 2, mode='reflect')
    T = 1 + (len(xp) - N_FFT) // HOP
    idx = np.arange(T)[None, :] * HOP + n[:, None]
    fr = w[:, None] * xp[idx]                                # float64 product, as librosa forms it
    return np.fft.rfft(fr, axis=0).astype(np.complex64)


def band_matrix():
    """``compute_band_E`` as a fixed [64, 257] linear map on the squared magnitudes."""
    W = np.zeros((NB_BANDS, FREQ))
    for i in range(NB_BANDS - 1):
        size = GMTBAND[i + 1] - GMTBAND[i]
        for j in range(size):
            frac = float(j) / size
            W[i, GMTBAND[i] + j] += 1 - frac
            W[i + 1, GMTBAND[i] + j] += frac
    return W


_W = band_matrix()


def compute_band_E(X):
    """audio_util.py:30-50.  X: magnitudes [T, 257] -> float32 [T, 64]."""
    X = np.asarray(X)
    return ((X * X).astype(np.float64) @ _W.T).astype(np.float32)


# ---- IMCRA (noise_est/imcra.py) -------------------------------------------------------------
IS, ALPHA_S, ALPHA_D, U, V = 15, 0.9, 0.85, 8, 15          # :181-196 (imcra_est passes IS=15, :492,513)
GAMMA0, GAMMA1, ZETA0, BETA, BMIN = 4.6, 3.0, 1.67, 1.47, 3.2   # :218-228, Bmin :492
ALPHA_DD, XI_MIN, P_UP = 0.92, 10 ** (-25. / 20), 0.9      # :492, :303


def _fsmooth(P):
    """:336-337 with w = 1: normalised [0.5, 1, 0.5] window, truncated at the edges (:262-272)."""
    K = len(P)
    out = np.empty(K)
    out[1:-1] = (0.5 * P[:-2] + P[1:-1] + 0.5 * P[2:]) / 2.0
    out[0] = (P[0] + 0.5 * P[1]) / 1.5
    out[-1] = (0.5 * P[-2] + P[-1]) / 1.5
    return out


def imcra_noise_psd(Y):
    """``imcra_est(nfft=512).estimate(Y)`` for a fresh estimator: complex64 [257, L] -> float32 [257, L]."""
    K, L = Y.shape
    out = np.zeros((K, L), dtype=np.float32)
    G = np.ones(K)
    Gamma = np.ones(K)
    Lam = 1e-6 * np.ones(K)                                  # :518
    store = np.zeros((K, U))
    tstore = np.zeros((K, U))
    j = u = 0
    for l in range(L):
        a = np.abs(Y[:, l])
        P = (a * a).astype(np.float64)                       # float32 square, then promoted by the division
        xi_G = G * G * Gamma                                 # :544
        Gamma = P / Lam                                      # :546
        xi = ALPHA_DD * xi_G + (1 - ALPHA_DD) * np.maximum(Gamma - 1, 1e-6)   # :548-551
        xi = np.maximum(xi, XI_MIN)                          # :553
        G = xi / (1 + xi)                                    # :557
        # imcra.update (:362-484)
        if l == 0:                                           # init_params (:339-360)
            S = _fsmooth(P)
            tS, Smin, tSmin, Smin_sw, tSmin_sw = S.copy(), S.copy(), S.copy(), S.copy(), S.copy()
            ovLam = P.copy()
            Lam = (a * a)                                    # float32 array: stays float32 through the initial segment
        Sf = _fsmooth(P)
        S = ALPHA_S * S + (1 - ALPHA_S) * Sf
        Smin = np.minimum(Smin, S)
        Smin_sw = np.minimum(Smin_sw, S)
        if l < IS:                                           # :384-399
            # python scalars times float32 arrays: the reference runs this recursion in float32 (:395)
            Lam = np.float32(ALPHA_D) * Lam + np.float32(1 - ALPHA_D) * (a * a)
        else:                                                # :401-482
            Gmin = P / (BMIN * Smin)
            zeta = S / (BMIN * Smin)
            I = ((Gmin < GAMMA0) & (zeta < ZETA0)).astype(np.float64)
            norm = _fsmooth(I)
            tSf = _fsmooth(I * P)
            nz = norm > 0
            tSf[nz] = tSf[nz] / norm[nz]
            tS = ALPHA_S * tS + (1 - ALPHA_S) * tSf
            tSmin = np.minimum(tSmin, tS)
            tSmin_sw = np.minimum(tSmin_sw, tS)
            tG = P / (BMIN * tSmin)
            tz = S / (BMIN * tSmin)
            q = np.zeros(K)
            q[(tG <= 1) & (tz < ZETA0)] = 1
            m = (1 < tG) & (tG < GAMMA1) & (tz < ZETA0)
            q[m] = (GAMMA1 - tG[m]) / (GAMMA1 - 1)
            nu = Gamma * xi / (1 + xi)                       # post_speech_prob (:23-38)
            p = np.zeros(K)
            s = q < 1
            p[s] = 1. / (1 + (q[s] / (1 - q[s])) * (1 + xi[s]) * np.exp(-nu[s]))
            p = np.minimum(p, P_UP)
            ta = ALPHA_D + (1 - ALPHA_D) * p
            ovLam = ta * ovLam + (1 - ta) * P
            Lam = BETA * ovLam
            j += 1
            if j == V:                                       # :451-482
                if u < U:
                    store[:, u] = Smin_sw
                    tstore[:, u] = tSmin_sw
                else:
                    store = np.roll(store, -1, axis=1)
                    store[:, -1] = Smin_sw
                    tstore = np.roll(tstore, -1, axis=1)
                    tstore[:, -1] = tSmin_sw
                Smin = store[:, :u + 1].min(axis=1)
                tSmin = tstore[:, :u + 1].min(axis=1)
                Smin_sw = S.copy()
                tSmin_sw = tS.copy()
                j = 0
                u += 1
        out[:, l] = Lam
    return out


def sp_and_phase_speech(signal, power, Normalization=True):
    """audio_util.py:422-437 -> (bandE [T, 64] f32, mag [257, T] f32, phase [257, T] f32)."""
    F = stft(signal)
    mag = np.abs(F)
    phase = np.angle(F)
    bandE = compute_band_E(mag.T)
    if Normalization:
        bandE = bandE ** power
    return bandE, mag, phase


def sp_and_phase_noise(signal, power, Normalization=True):
    """audio_util.py:439-457: band energies of the IMCRA noise PSD, magnitude and phase of the STFT."""
    F = stft(signal)
    psd = imcra_noise_psd(F).T
    bandE = compute_band_E(np.sqrt(psd))
    if Normalization:
        bandE = bandE ** power
    return bandE, np.abs(F), np.angle(F)
