"""Generate ``tests/golden/features_ref.npz`` from the reference's own feature front-end.

``audio_util.py`` cannot be imported (librosa / pysiib / pystoi / pypesq are not installed), so the
source text of ``compute_band_E``, ``STFT``, ``NoisePSD``, ``Sp_and_phase_Speech`` and
``Sp_and_phase_Noise`` is cut out of the unmodified file and executed in a namespace that provides
numpy, the UNMODIFIED ``noise_est/imcra.py`` (imported from /root/reference: pure numpy) and a
``librosa`` object whose only member is the oracle's restated ``stft`` (librosa 0.7.1 is un-vendored;
that step stays "parity unpinned").  Run in the build container only:
``python tests/golden/make_golden_features.py``.

Per case: the float32 input, the complex64 STFT the reference functions were given, and their outputs
(band energies of speech and noise with power 1/6 as dataloader.py:14 sets it, the IMCRA noise PSD).
"""
import os
import re
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, "/root/reference")

from noise_est.imcra import imcra, imcra_est  # noqa: E402  (the unmodified reference)

from nele_gan_b200.synth import make_pair  # noqa: E402
from oracle import features_np  # noqa: E402

src = open("/root/reference/audio_util.py").read()


def cut(name):
    m = re.search(r"^def %s\(.*?(?=^def |\Z)" % name, src, flags=re.S | re.M)
    return m.group(0)


gm = re.search(r"^gmtband = \[.*?\]", src, flags=re.S | re.M).group(0)
librosa = types.SimpleNamespace(stft=lambda x, n_fft, hop_length, win_length: features_np.stft(x))
ns = {"np": np, "librosa": librosa, "imcra": imcra, "imcra_est": imcra_est, "NB_BANDS": 64}
exec(gm, ns)
for fn in ("compute_band_E", "STFT", "NoisePSD", "Sp_and_phase_Speech", "Sp_and_phase_Noise"):
    exec(cut(fn), ns)

POWER = 1 / 6          # dataloader.py:14
out = {}
# lengths: toy-data size, a ragged one, 3.0 s (188 frames: exercises the rolling minimum store,
# noise_est/imcra.py:455-470, which starts after 15 + 8 * 15 frames), a short one inside the
# initial segment + one minimum-store update
for i, L in ((0, 33536), (3, 52345), (5, 48000), (7, 8000)):
    x, y, _ = make_pair(i, L)
    noise = (y - x).astype(np.float32)
    k = "p%d_%d" % (i, L)
    out[k + "/speech"] = x
    out[k + "/noise"] = noise
    b, mag, ph = ns["Sp_and_phase_Speech"](x, POWER)
    out[k + "/speech_band"] = b
    out[k + "/speech_mag"] = mag
    out[k + "/speech_phase"] = ph
    b, mag, ph = ns["Sp_and_phase_Noise"](noise, POWER)
    out[k + "/noise_band"] = b
    out[k + "/noise_mag"] = mag
    out[k + "/noise_psd"] = ns["NoisePSD"](ns["STFT"](noise))
    out[k + "/noise_band_raw"] = ns["Sp_and_phase_Noise"](noise, POWER, Normalization=False)[0]
# a speech + noise mixture drives the speech-presence branches of IMCRA harder than stationary noise
x, y, _ = make_pair(11, 40000)
out["mix/signal"] = y
out["mix/psd"] = ns["NoisePSD"](ns["STFT"](y))
out["mix/band"] = ns["Sp_and_phase_Noise"](y, POWER)[0]
np.savez_compressed(os.path.join(HERE, "features_ref.npz"), **{k: v for k, v in out.items()})
print("wrote", len(out), "arrays")
for k in ("p0_33536", "p5_48000", "mix"):
    psd = out[k + ("/noise_psd" if k != "mix" else "/psd")]
    mine = features_np.imcra_noise_psd(features_np.stft(out[k + ("/noise" if k != "mix" else "/signal")]))
    print(k, "oracle vs reference IMCRA: max rel diff", np.max(np.abs(mine - psd) / psd))
