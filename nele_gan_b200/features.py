"""Host-side mirror of the reference's feature front-end, on the CUDA
engine: the per-utterance features the generator / discriminator data loaders compute
(dataloader.py:30-84).  Same names, argument order, defaults and return shapes as the reference:

  audio_util.py:422-437  Sp_and_phase_Speech(signal, power, Normalization=True) -> (bandE [T, 64], mag [257, T], phase [257, T])
  audio_util.py:439-457  Sp_and_phase_Noise(signal, power, Normalization=True)  -> the same, band energies of the IMCRA noise PSD
  audio_util.py:117-122  NoisePSD(MIXED) with MIXED = STFT(x)  -> here ``noise_psd(signal)`` (the engine owns the STFT)
  audio_util.py:30-50    compute_band_E(|STFT|)                -> here ``band_energies(signal)``

plus the batched forms the engine is built for (``speech_features`` / ``noise_features`` on lists, and
``features_tensors`` on CUDA tensors, which keeps everything on the device for the training loop).
Every call goes through ``nele_features`` of ``libnele_score.so``; nothing under ``oracle/`` is imported
and there is no numpy implementation here.
"""
import numpy as np

from . import engine as _eng

__all__ = ["Sp_and_phase_Speech", "Sp_and_phase_Noise", "noise_psd", "band_energies", "speech_features",
           "noise_features", "features_tensors", "power_law"]

power_law = 1 / 6   # dataloader.py:14


def Sp_and_phase_Speech(signal, power, Normalization=True):
    """audio_util.py:422-437."""
    return _eng.default_engine().features([signal], power=power, noise=False, normalization=Normalization)[0]


def Sp_and_phase_Noise(signal, power, Normalization=True):
    """audio_util.py:439-457."""
    return _eng.default_engine().features([signal], power=power, noise=True, normalization=Normalization)[0]


def noise_psd(signal):
    """``NoisePSD(STFT(signal))`` (audio_util.py:52-57, 117-122): float32 [257, T]."""
    return _eng.default_engine().features([signal], power=1.0, noise=True, normalization=False, want_psd=True)[0][3]


def band_energies(signal):
    """``compute_band_E(np.abs(STFT(signal)).T)`` (audio_util.py:30-57): float32 [T, 64]."""
    return _eng.default_engine().features([signal], power=1.0, noise=False, normalization=False)[0][0]


def speech_features(signals, power=power_law, Normalization=True):
    """``[Sp_and_phase_Speech(s, power) for s in signals]`` in one engine call."""
    return _eng.default_engine().features(signals, power=power, noise=False, normalization=Normalization)


def noise_features(signals, power=power_law, Normalization=True):
    """``[Sp_and_phase_Noise(s, power) for s in signals]`` in one engine call."""
    return _eng.default_engine().features(signals, power=power, noise=True, normalization=Normalization)


def features_tensors(wav, lengths=None, power=power_law, noise=False, Normalization=True, want_phase=True):
    """Device-resident form: ``wav`` CUDA float32 tensor ``[n, Lmax]``, ``lengths`` valid samples per
    row.  Returns ``(band [sum T, 64], mag [257 * sum T], phase or None, foff, frames)``: CUDA tensors
    in the layout of ``nele_features`` (row i's ``[257, T_i]`` matrices at ``257 * foff[i]``; with
    equal lengths ``mag.view(n, 257, T)``), ``foff`` / ``frames`` numpy int64."""
    import torch
    if not wav.is_cuda or wav.dim() != 2:
        raise ValueError("features_tensors needs a CUDA tensor [n, Lmax] (the engine has no CPU path)")
    wav = wav.detach().to(torch.float32).contiguous()
    n, lmax = wav.shape
    lens = np.full(n, lmax, dtype=np.int32) if lengths is None else np.asarray(lengths, dtype=np.int32)
    if lens.shape != (n,) or lens.min() <= 256 or lens.max() > lmax:
        raise ValueError("lengths must be n values in (256, Lmax]")
    offs = np.arange(n, dtype=np.int64) * lmax
    tot = int((1 + lens.astype(np.int64) // 256).sum())
    band = torch.empty((tot, 64), dtype=torch.float32, device=wav.device)
    mag = torch.empty(257 * tot, dtype=torch.float32, device=wav.device)
    phase = torch.empty(257 * tot, dtype=torch.float32, device=wav.device) if want_phase else None
    eng = _eng.default_engine(wav.device.index)
    torch.cuda.current_stream(wav.device).synchronize()   # the engine runs on its own stream
    r = eng.features_packed(wav.data_ptr(), offs, lens, power=power, noise=noise, normalization=Normalization,
                            device_io=True,
                            out=(band.data_ptr(), mag.data_ptr(), None if phase is None else phase.data_ptr(), None))
    return band, mag, phase, r["foff"], r["frames"]
