"""Generate ``tests/golden/haspi_ref.npz`` by running the UNMODIFIED reference
``/root/reference/pyHASPI/pyhaspi2.py`` (behind ``oracle/librosa_shim``).

Run in the build container only (``/root/reference`` does not exist on the GPU
box):  ``python tests/golden/make_golden.py``.  The committed .npz is what the
tests read.

Per case the fixture holds the inputs (float32, as the reference receives them
from ``librosa.load``) and the reference's outputs:
  * ``v2_zero``  / ``v2_zero_raw``  : haspi_v2 with every ``np.random.randn``
    forced to zero (no BM noise, no cepstral dither);
  * ``v2_dith``  / ``v2_dith_raw``  : haspi_v2 with the cepstral dither taken
    from a fixed stream (see ``dither_matrix``) and BM noise zero -- the
    "shared dither" form the engine is compared with;
  * ``v2_seed0``                    : haspi_v2 after ``np.random.seed(0)``,
    i.e. the reference exactly as a user runs it;
  * ``v1_zero`` / ``v1_zero_raw``, ``v1_seed0``: the same for ``haspi`` (v1);
  * stage tensors recorded by wrapping the reference's own functions:
    ``bwx``/``bwy`` (eb_BWadjust returns), ``nsel`` (frames kept by
    ebm_CepCoef), ``xlp``/``ylp`` (ebm_EnvFilt output, every 7th row, f32).
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle", "librosa_shim"))
sys.path.insert(0, "/root/reference")

import librosa  # noqa: E402  (the shim)
from pyHASPI import pyhaspi2 as REF  # noqa: E402  (the unmodified reference)

from nele_gan_b200.synth import make_pair  # noqa: E402
from oracle.intel_np import read_pair  # noqa: E402

DITHER_SEEDS = (1234, 5678)
DITHER_ROWS = 16384
LP_STRIDE = 7


def dither_matrix(which):
    """The shared dither stream: rows [:n] of a fixed (16384, 32) normal draw,
    one matrix for x (which=0) and one for y (which=1)."""
    return np.random.RandomState(DITHER_SEEDS[which]).standard_normal((DITHER_ROWS, 32))


class PatchedRandn:
    """Replaces ``numpy.random.randn`` while the reference runs."""

    def __init__(self, mode):
        self.mode = mode
        self.calls2d = 0

    def __call__(self, *shape):
        if len(shape) == 1 or self.mode == "zero":
            return np.zeros(shape)
        m = dither_matrix(self.calls2d % 2)[: shape[0], : shape[1]]
        self.calls2d += 1
        return m


def run_ref(fn, x, fx, y, fy, mode):
    rec = {"bw": [], "nsel": None, "lp": None}
    orig = (np.random.randn, REF.eb_BWadjust, REF.ebm_EnvFilt, REF.ebm_CepCoef)

    def bw_wrap(*a, **k):
        v = orig[1](*a, **k)
        rec["bw"].append(v)
        return v

    def filt_wrap(*a, **k):
        v = orig[2](*a, **k)
        rec["lp"] = v
        return v

    def cep_wrap(*a, **k):
        v = orig[3](*a, **k)
        rec["nsel"] = v[0].shape[0]
        return v

    try:
        if mode == "seed0":
            np.random.seed(0)
        else:
            np.random.randn = PatchedRandn(mode)
        REF.eb_BWadjust, REF.ebm_EnvFilt, REF.ebm_CepCoef = bw_wrap, filt_wrap, cep_wrap
        score, raw = fn(x, fx, y, fy)
    finally:
        np.random.randn, REF.eb_BWadjust, REF.ebm_EnvFilt, REF.ebm_CepCoef = orig
    return float(score), np.asarray(raw, dtype=np.float64), rec


def cases():
    x, fx = librosa.load('/root/reference/pyHASPI/sig_clean.wav', sr=None)
    y, fy = librosa.load('/root/reference/pyHASPI/sig_out.wav', sr=None)
    yield "bundled_22050", x, y, fx
    x16, _ = librosa.load('/root/reference/pyHASPI/sig_clean.wav', sr=16000)
    y16, _ = librosa.load('/root/reference/pyHASPI/sig_out.wav', sr=16000)
    yield "bundled_16000", x16, y16, 16000
    T = '/root/reference/toy_dataset/'
    tr = 'f_hvd_100#Babble#-11.wav'
    te = 'f_hvd_669#AirportAnnouncement#-9.wav'
    x, y = read_pair(T + 'Train/Clean/', T + 'Train/Noise/', T + 'Train/MultiEnh/' + tr, drc=True)
    yield "toy_train_multienh", x, y, 16000
    x, y = read_pair(T + 'Train/Clean/', T + 'Train/Noise/', T + 'Train/Clean/' + tr, drc=True)
    yield "toy_train_clean", x, y, 16000
    x, y = read_pair(T + 'Test/Clean/', T + 'Test/Noise/', T + 'Test/Clean/' + te, drc=True)
    yield "toy_test_clean", x, y, 16000
    for i, n in ((0, 24000), (1, 31999), (2, 48000)):
        x, y, _ = make_pair(i, n)
        yield "synth_%d_%d" % (i, n), x, y, 16000


def main():
    out = {}
    names = []
    for name, x, y, fs in cases():
        names.append(name)
        x = np.ascontiguousarray(x, dtype=np.float32)
        y = np.ascontiguousarray(y, dtype=np.float32)
        out[name + "/x"] = x
        out[name + "/y"] = y
        out[name + "/fs"] = np.int64(fs)
        for mode in ("zero", "dith", "seed0"):
            s, raw, rec = run_ref(REF.haspi_v2, x, fs, y, fs, mode)
            out["%s/v2_%s" % (name, mode)] = np.float64(s)
            out["%s/v2_%s_raw" % (name, mode)] = raw
            if mode == "zero":
                bw = np.asarray(rec["bw"], dtype=np.float64).reshape(32, 2)
                out[name + "/bwx"] = bw[:, 0].copy()
                out[name + "/bwy"] = bw[:, 1].copy()
                out[name + "/nsel"] = np.int64(rec["nsel"])
                out[name + "/xlp"] = rec["lp"][0][::LP_STRIDE].astype(np.float32)
                out[name + "/ylp"] = rec["lp"][1][::LP_STRIDE].astype(np.float32)
        for mode in ("zero", "seed0"):
            s, raw, _ = run_ref(REF.haspi, x, fs, y, fs, mode)
            out["%s/v1_%s" % (name, mode)] = np.float64(s)
            out["%s/v1_%s_raw" % (name, mode)] = raw
        print(name, len(x), fs, out[name + "/v2_zero"], out[name + "/v2_dith"], out[name + "/v2_seed0"],
              out[name + "/v1_zero"], out[name + "/v1_seed0"], flush=True)
    out["names"] = np.array(names)
    np.savez_compressed(os.path.join(HERE, "haspi_ref.npz"), **out)
    print("wrote", os.path.join(HERE, "haspi_ref.npz"))


if __name__ == "__main__":
    main()
