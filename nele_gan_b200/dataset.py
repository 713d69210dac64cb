"""Device-resident discriminator data set of one sampling round.

The reference hands the labels of a sampling round to the discriminators as strings
``"siib,haspi,estoi,pesq,visqol,path"`` (audio_util.py:367-389); ``Discriminator_train_dataset.__getitem__``
(dataloader.py:54-84) splits such a string again, re-loads three WAV files and recomputes the 64-band features of
the enhanced, noise and clean signals with numpy -- per item, per epoch, in 8 DataLoader workers.  Here the round is
labelled on the device (``inloop.label_sampling_round``), the band features of all its utterances come from one
``nele_features`` call per signal kind, and an item is a few views into those tensors: nothing leaves the GPU
between the generator and the discriminators.  ``records()`` still produces the reference's strings for code that
wants them (the replay list of train_nele.py:341-345).

Item layout = what dataloader.py:83 returns, as CUDA tensors:
  ``x3  [3, 64, T]``  enhanced / noise / clean band features (input of ``Discriminator``)
  ``x2  [2, 64, T]``  enhanced / clean                       (input of ``Discriminator_Quality``)
  ``True_score [3]``  {SIIB, HASPI, ESTOI}, ``True_score_Qua [2]``  {PESQ, ViSQOL} (external binaries: passed in or 0)
"""
import numpy as np

from . import features as _feat
from . import records as _rec


class DiscriminatorRoundDataset:
    """``len()`` / ``[]`` like a ``torch.utils.data.Dataset`` (usable with ``DataLoader(batch_size=1, num_workers=0)``,
    which is how the reference iterates it, dataloader.py:93-98)."""

    def __init__(self, enh_band, noise_band, clean_band, foff, frames, scores, quality=None, names=None):
        import torch
        self.enh_band, self.noise_band, self.clean_band = enh_band, noise_band, clean_band   # [sum T, 64] each
        self.foff, self.frames = np.asarray(foff, dtype=np.int64), np.asarray(frames, dtype=np.int64)
        n = len(self.frames)
        dev = enh_band.device
        self.scores = torch.as_tensor(np.asarray(scores, dtype=np.float32).reshape(n, 3), device=dev)
        q = np.zeros((n, 2), np.float32) if quality is None else np.asarray(quality, dtype=np.float32).reshape(n, 2)
        self.quality = torch.as_tensor(q, device=dev)
        self.names = list(names) if names is not None else ["utt%05d.wav" % i for i in range(n)]
        if len(self.names) != n or self.foff.shape != (n,):
            raise ValueError("per-utterance arguments differ in length")

    @classmethod
    def from_round(cls, enh_wav, clean_wav, noise_wav, lengths, out_lens, scores, quality=None, names=None,
                   power=_feat.power_law):
        """``enh_wav`` / ``clean_wav`` / ``noise_wav``: CUDA float32 ``[n, L]``; ``lengths`` the utterance lengths,
        ``out_lens`` the valid samples of ``enh_wav`` (``256 * (lengths // 256)``, what ``label_sampling_round``
        returns).  Features exactly as dataloader.py:59-72: ``Sp_and_phase_Speech`` of the enhanced and the clean
        signal, ``Sp_and_phase_Noise`` (IMCRA) of the noise, power-law 1/6."""
        lengths, out_lens = np.asarray(lengths, dtype=np.int32), np.asarray(out_lens, dtype=np.int32)
        if np.any(1 + out_lens // 256 != 1 + lengths // 256):
            raise ValueError("enhanced and clean signals must have the same number of STFT frames")
        eb, _, _, foff, frames = _feat.features_tensors(enh_wav, out_lens, power=power, noise=False, want_phase=False)
        cb, _, _, _, _ = _feat.features_tensors(clean_wav, lengths, power=power, noise=False, want_phase=False)
        nb, _, _, _, _ = _feat.features_tensors(noise_wav, lengths, power=power, noise=True, want_phase=False)
        return cls(eb, nb, cb, foff, frames, scores, quality, names)

    def __len__(self):
        return len(self.frames)

    def __getitem__(self, i):
        import torch
        if i < 0:
            i += len(self)
        if not 0 <= i < len(self):
            raise IndexError(i)
        a, b = int(self.foff[i]), int(self.foff[i] + self.frames[i])
        e, nz, c = self.enh_band[a:b].t(), self.noise_band[a:b].t(), self.clean_band[a:b].t()   # [64, T] views
        return torch.stack((e, nz, c)), torch.stack((e, c)), self.scores[i], self.quality[i]

    def records(self):
        """The round as the reference's strings (audio_util.py:367-389; train_nele.py:327-328)."""
        s = self.scores.cpu().numpy().astype(np.float64)
        q = self.quality.cpu().numpy().astype(np.float64)
        return _rec.round_records(s, self.names, pesq=[float(v) for v in q[:, 0]], visqol=[float(v) for v in q[:, 1]])
