"""In-loop tensor boundary: label the generator's outputs of one
GAN sampling round without the WAV round trip of the reference.

The reference (train_nele.py:286-322) takes the generator's band gains ``mask * beta_2``
[T, 64] for one utterance at a time, moves them to the CPU, interpolates them to 257 STFT bins
(``interp_band_gain``, audio_util.py:98-115), scales the clean complex spectrogram, inverts it
with ``librosa.istft`` (``Resyn`` / ``ISTFT``, audio_util.py:60-96), writes PCM-16 WAV files,
and fans the file names out to 32 processes that re-load them, add the noise and compute the
three metrics.  Here everything up to the degraded waveform is one kernel of the engine
(``nele_resyn``, csrc/features.cu) and the waveforms go straight into the scoring kernels
(device pointers through the C ABI): no files, no process pool, no torch / cuFFT kernels.
``stft`` / ``resyn`` below are the same two steps as batched torch ops for equally long
utterances; they are kept as the host-side restatement the CPU tests check against
``oracle/resyn_np.py`` and are not on the product path.
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


def label_sampling_round(alpha2, clean_wav, noise_wav, lengths=None, norm=True, pcm16=True, seed=0, strict=True,
                         return_deg=False, return_enh=False, **kw):
    """One sampling round (train_nele.py:286-322) on the device, through the engine's own kernels:
    ``nele_resyn`` (band gains -> ``interp_band_gain`` -> ``Resyn`` / ``librosa.istft`` -> PCM-16 rounding ->
    ``+ noise``, one launch for the whole round) and ``nele_score_batch`` on the device buffers.  No torch / cuFFT
    kernel runs: torch only owns the memory.

    alpha2     [n, Tmax, 64]  generator output after the energy normalisation, ``mask * beta_2`` (CUDA, float32);
                              utterance i uses its first ``1 + lengths[i] // 256`` rows
    clean_wav  [n, L]         clean waveforms, zero padded to the longest (CUDA, float32)
    noise_wav  [n, L]         noise waveforms
    lengths    [n]            valid samples per utterance (default L)

    Every utterance is resynthesised at its own length -- reflect padding at its own ends and
    ``256 * (length // 256)`` output samples, to which the reference trims both signals (audio_util.py:190-193) --
    exactly as the reference does one file at a time, so ragged rounds give the per-file numbers.
    Returns float64 ``[n, 3]`` {SIIB, HASPI, ESTOI} like ``api.score_tensors``; with ``return_deg`` /
    ``return_enh`` a tuple ``(scores[, deg][, enh], out_lens)``: the degraded waveforms, the (PCM-16 rounded)
    enhanced waveforms the discriminator's data loader reads back (dataloader.py:59) -- both ``[n, L]`` on the
    device -- and their valid lengths."""
    import torch
    from . import api, engine as _eng
    if not (alpha2.is_cuda and clean_wav.is_cuda and noise_wav.is_cuda):
        raise ValueError("label_sampling_round needs CUDA tensors (the engine has no CPU path)")
    n, L = clean_wav.shape
    if noise_wav.shape != clean_wav.shape or alpha2.dim() != 3 or alpha2.shape[0] != n or alpha2.shape[2] != NB_BANDS:
        raise ValueError("shapes: alpha2 [n, Tmax, 64], clean_wav / noise_wav [n, L]")
    lens = np.full(n, L, dtype=np.int32) if lengths is None else np.asarray(lengths, dtype=np.int32)
    if lens.shape != (n,) or lens.min() <= 256 or lens.max() > L:
        raise ValueError("lengths must be n values in (256, L]")
    if int((1 + lens // HOP).max()) > alpha2.shape[1]:
        raise ValueError("alpha2 has %d frames, the longest utterance needs %d" % (alpha2.shape[1], int((1 + lens // HOP).max())))
    clean = clean_wav.detach().to(torch.float32).contiguous()
    noise = noise_wav.detach().to(torch.float32).contiguous()
    if L % 4:                                              # rows must start 16-byte aligned
        clean = torch.nn.functional.pad(clean, (0, 4 - L % 4))
        noise = torch.nn.functional.pad(noise, (0, 4 - L % 4))
    a2 = alpha2.detach().to(torch.float32).contiguous()
    stride = clean.shape[1]
    offs = np.arange(n, dtype=np.int64) * stride
    arow = np.arange(n, dtype=np.int64) * a2.shape[1]
    deg = torch.zeros_like(clean)
    enh = torch.zeros_like(clean) if return_enh else None
    eng = _eng.default_engine(clean.device.index)
    torch.cuda.current_stream(clean.device).synchronize()  # the engine runs on its own stream
    out_lens = eng.resyn(clean.data_ptr(), noise.data_ptr(), offs, lens, a2.data_ptr(), arow=arow, deg=deg.data_ptr(),
                         enh=None if enh is None else enh.data_ptr(), pcm16=pcm16, enh_rounded=pcm16)
    r = eng.score_packed(clean.data_ptr(), deg.data_ptr(), offs, out_lens, fs=16000, mapped=bool(norm), seed=seed,
                         device_input=True, **kw)
    scores = torch.from_numpy(api.check_status(r, strict=strict).scores)
    if return_deg or return_enh:
        return (scores,) + ((deg[:, :L],) if return_deg else ()) + ((enh[:, :L],) if return_enh else ()) + (out_lens,)
    return scores
