"""Synthetic speech-shaped (clean, degraded) pairs for tests and benchmarks
(SURVEY.md section 8(d), configs 3 and 5).

Pair ``i`` is a pure function of ``seed + i``: white Gaussian noise through an
LTASS-like shaping filter (one-pole low-pass at 500 Hz cascaded with a
one-pole high-pass at 80 Hz), multiplied by a 4 Hz syllabic envelope
``max(0, sin(2*pi*4*t + phi))**2`` with a 0.3 s pause every 1.2 s (so the VAD,
silent-frame and loudness-threshold logic of the three metrics is exercised),
scaled to RMS 0.03 as the reference normalises its corpus (README.md:36).  The
degradation is independent stationary speech-shaped noise at an SNR drawn from
{-11, -9, -7, -5, -3, -1} dB; ``deg = speech + noise`` in float32, unclipped,
exactly how audio_util.py:139 forms the degraded signal.
"""
import numpy as np
from scipy.signal import lfilter

FS = 16000
SNRS_DB = (-11.0, -9.0, -7.0, -5.0, -3.0, -1.0)


def _shape(w):
    a_lp = np.exp(-2 * np.pi * 500.0 / FS)
    a_hp = np.exp(-2 * np.pi * 80.0 / FS)
    v = lfilter([1 - a_lp], [1, -a_lp], w, axis=-1)
    return lfilter([(1 + a_hp) / 2, -(1 + a_hp) / 2], [1, -a_hp], v, axis=-1)


def make_pair(i, n_samples, seed=666_000):
    """-> (ref float32[n_samples], deg float32[n_samples], snr_db)."""
    rng = np.random.default_rng(seed + i)
    t = np.arange(n_samples) / FS
    phi = rng.uniform(0, 2 * np.pi)
    env = np.maximum(0.0, np.sin(2 * np.pi * 4.0 * t + phi)) ** 2
    env *= ((t + rng.uniform(0, 1.2)) % 1.2) >= 0.3
    speech = _shape(rng.standard_normal(n_samples)) * env
    speech *= 0.03 / np.sqrt(np.mean(speech ** 2))
    snr = SNRS_DB[int(rng.integers(len(SNRS_DB)))]
    noise = _shape(rng.standard_normal(n_samples))
    noise *= 0.03 * 10 ** (-snr / 20) / np.sqrt(np.mean(noise ** 2))
    ref = speech.astype(np.float32)
    deg = (ref + noise.astype(np.float32)).astype(np.float32)
    return ref, deg, snr


def make_batch(n, n_samples, seed=666_000, unique=None):
    """``n`` pairs of ``n_samples`` (int, or a length-n array for a ragged
    batch).  ``unique`` caps the number of distinct generated pairs; the rest
    repeat them cyclically (benchmark set-up time only -- every pair is still
    scored independently)."""
    lens = np.broadcast_to(np.asarray(n_samples, dtype=np.int64), (n,))
    refs, degs = [], []
    u = n if unique is None else min(n, unique)
    for i in range(n):
        if i < u:
            r, d, _ = make_pair(i, int(lens[i]), seed)
        else:
            r, d = refs[i % u], degs[i % u]
            if len(r) != lens[i]:
                reps = -(-int(lens[i]) // len(r))
                r, d = np.tile(r, reps)[: lens[i]], np.tile(d, reps)[: lens[i]]
        refs.append(r)
        degs.append(d)
    return refs, degs
