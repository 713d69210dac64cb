"""In-loop tensor boundary: label the generator's outputs of one
GAN sampling round without the WAV round trip of the reference.

The reference (train_nele.py:286-322) takes the generator's band gains ``mask * beta_2``
[T, 64] for one utterance at a time, moves them to the CPU, interpolates them to 257 STFT bins
(``interp_band_gain``, audio_util.py:98-115), scales the clean complex spectrogram, inverts it
with ``librosa.istft`` (``Resyn`` / ``ISTFT``, audio_util.py:60-96), writes PCM-16 WAV files,
and fans the file names out to 32 processes that re-load them, add the noise and compute the
three metrics.  Here everything up to the degraded waveform is a handful of batched torch ops
on the device (plumbing), and the waveforms go straight into the scoring engine
(``api.score_tensors``, device pointers through the C ABI): no files, no process pool.
"""
import numpy as np

# audio_util.py:23 -- band edges (STFT bins) of the 64 ERB-like bands
GMTBAND = [0, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15, 16, 17, 18, 19, 20, 21, 22, 23, 24, 25, 26, 28, 30, 32, 34,
           36, 38, 41, 43, 46, 49, 52, 55, 58, 62, 66, 70, 74, 79, 83, 88, 93, 99, 105, 111, 117, 124, 131, 139, 147,
           156, 165, 174, 184, 195, 206, 218, 230, 243, 257]
NB_BANDS = 64
N_FFT, HOP, FREQ_SIZE = 512, 256, 257


def band_gain_matrix():
    """``interp_band_gain`` (audio_util.py:98-115) as a fixed linear map: returns ``(W [257, 64],
    fixed [257])`` with ``g = W @ bandE`` wherever ``fixed`` is NaN and ``g = fixed`` elsewhere
    (bins 0, 1 and 256 are overwritten with 1e-4, 1e-4, 1e-2)."""
    W = np.zeros((FREQ_SIZE, NB_BANDS))
    for i in range(NB_BANDS - 1):
        size = GMTBAND[i + 1] - GMTBAND[i]
        for j in range(size):
            frac = float(j) / size
            W[GMTBAND[i] + j, i] = 1 - frac
            W[GMTBAND[i] + j, i + 1] = frac
    fixed = np.full(FREQ_SIZE, np.nan)
    fixed[0], fixed[1], fixed[256] = 1e-4, 1e-4, 1e-2
    return W, fixed


def stft(wav):
    """``librosa.stft(x, n_fft=512, hop_length=256, win_length=512)`` (audio_util.py:52-57) for a
    batch of equally long waveforms ``[n, L]`` -> complex ``[n, 257, T]``."""
    import torch
    win = torch.hann_window(N_FFT, periodic=True, device=wav.device, dtype=wav.dtype)
    return torch.stft(wav, N_FFT, hop_length=HOP, win_length=N_FFT, window=win, center=True, pad_mode='reflect',
                      return_complex=True)


def resyn(X, alpha2):
    """``Resyn`` (audio_util.py:84-96): ``X`` complex ``[n, 257, T]`` clean spectrograms,
    ``alpha2`` ``[n, T, 64]`` band energy gains -> waveforms ``[n, 256 (T - 1)]``
    (``librosa.istft(hop_length=256, win_length=512)``)."""
    import torch
    W, fixed = band_gain_matrix()
    Wt = torch.as_tensor(W, device=X.device, dtype=alpha2.dtype)
    g = torch.einsum('fb,ntb->nft', Wt, alpha2)
    fx = torch.as_tensor(np.nan_to_num(fixed), device=X.device, dtype=alpha2.dtype)
    keep = torch.as_tensor(np.isnan(fixed), device=X.device)
    g = torch.where(keep[None, :, None], g, fx[None, :, None].expand_as(g))
    Xn = torch.sqrt(g).to(X.dtype) * X
    win = torch.hann_window(N_FFT, periodic=True, device=X.device, dtype=alpha2.dtype)
    return torch.istft(Xn, N_FFT, hop_length=HOP, win_length=N_FFT, window=win, center=True)


def label_sampling_round(alpha2, clean_wav, noise_wav, lengths=None, norm=True, pcm16=True, seed=0, **kw):
    """One sampling round (train_nele.py:286-322) on the device.

    alpha2     [n, T, 64]  generator output after the energy normalisation, ``mask * beta_2``
    clean_wav  [n, L]      clean waveforms (zero padded to the longest)
    noise_wav  [n, L]      noise waveforms
    lengths    [n]         valid samples per utterance (default L)

    Returns float64 ``[n, 3]`` {SIIB, HASPI, ESTOI} like ``api.score_tensors``.  ``pcm16``
    reproduces the quantisation of ``sf.write(..., 'PCM_16')`` (train_nele.py:313) on the
    enhanced waveform before the noise is added, as the reference's reload does
    (audio_util.py:186-196)."""
    import torch
    from . import api
    n, L = clean_wav.shape
    X = stft(clean_wav)
    enh = resyn(X, alpha2)
    if pcm16:
        enh = torch.clamp(torch.round(enh * 32768.0), -32768.0, 32767.0) / 32768.0
    lens = np.full(n, L, dtype=np.int64) if lengths is None else np.asarray(lengths, dtype=np.int64)
    m = min(L, enh.shape[1])
    lens = np.minimum(lens, m)                            # audio_util.py:190-193: trim to the shorter of clean / enhanced
    deg = enh[:, :m] + noise_wav[:, :m]
    return api.score_tensors(clean_wav[:, :m], deg, lengths=lens.astype(np.int32), norm=norm, seed=seed, **kw)
