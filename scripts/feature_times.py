"""Feature front-end probe (not the benchmark): per-kernel CUDA-event times of nele_features on n
device-resident waveforms of L samples, speech and noise paths, against the algorithmic HBM bytes
(wav read once; mag + phase + band written once; IMCRA reads mag once and writes the band energies),
plus the CPU oracle (oracle/features_np.py = the reference's numpy code path) on a few waveforms.
usage: feature_times.py [n] [L] [cpu_waveforms]"""
import json
import sys
import time

sys.path.insert(0, ".")
import numpy as np
import torch

from nele_gan_b200 import features as F
from nele_gan_b200.engine import default_engine
from nele_gan_b200.synth import make_pair

n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
L = int(sys.argv[2]) if len(sys.argv) > 2 else 48000
ncpu = int(sys.argv[3]) if len(sys.argv) > 3 else 8
base = np.stack([make_pair(300 + i, L)[1] for i in range(16)])
wav = torch.from_numpy(np.tile(base, (n // 16, 1))).cuda()
T = 1 + L // 256
eng = default_engine(0)
eng.set_profiling(True)
res = {"n": n, "L": L, "frames": T}
for noise in (False, True):
    for it in range(4):
        out = F.features_tensors(wav, noise=noise)
    ms, nl = eng.last_timing()
    kt = eng.kernel_times()
    key = "noise" if noise else "speech"
    stft_bytes = n * (4 * L + T * (2 * 257 * 4 + (0 if noise else 64 * 4)))
    imcra_bytes = n * T * (257 * 4 + 64 * 4)
    res[key] = {"ms": ms, "launches": nl, "audio_s_per_s": n * L / 16000 / (ms / 1e3),
                "kernels": {k: v[0] for k, v in kt.items()},
                "feat_stft_GBps": stft_bytes / (kt["feat_stft"][0] * 1e-3) / 1e9}
    if noise:
        res[key]["feat_imcra_GBps"] = imcra_bytes / (kt["feat_imcra"][0] * 1e-3) / 1e9
    del out
if ncpu:
    from oracle import features_np
    t0 = time.perf_counter()
    for i in range(ncpu):
        features_np.sp_and_phase_speech(base[i % 16], 1 / 6)
    t1 = time.perf_counter()
    for i in range(ncpu):
        features_np.sp_and_phase_noise(base[i % 16], 1 / 6)
    t2 = time.perf_counter()
    res["cpu_oracle_1core_audio_s_per_s"] = {"speech": ncpu * L / 16000 / (t1 - t0), "noise": ncpu * L / 16000 / (t2 - t1)}
print(json.dumps(res))
