"""Generate ``tests/golden/intel_helpers.npz`` from the reference's own helper
functions ``framing`` / ``get_vad`` / ``stft`` (intel.py:16-54) and its logistic
mappings (intel.py:102-140).

``intel.py`` cannot be imported (it needs pysiib / pystoi / pypesq), so the
source text of exactly those functions is cut out of the unmodified file and
executed in a namespace that provides numpy, scipy.fftpack.fft and a
``get_window`` that accepts the 'hanning' alias scipy >= 1.13 dropped.  Run in
the build container only: ``python tests/golden/make_golden_intel.py``.
"""
import os
import re
import sys

import numpy as np
from scipy.fftpack import fft
from scipy.signal import get_window as _gw

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from nele_gan_b200.synth import make_pair  # noqa: E402

src = open("/root/reference/intel.py").read()


def cut(name):
    m = re.search(r"^def %s\(.*?(?=^def |\Z)" % name, src, flags=re.S | re.M)
    return m.group(0)


ns = {"np": np, "fft": fft, "EPS": np.finfo(np.float64).eps,
      "get_window": lambda w, n: _gw("hann" if w == "hanning" else w, n)}
for fn in ("framing", "get_vad", "stft", "mapping_SIIB_harvard", "mapping_HASPI_harvard", "mapping_ESTOI_harvard"):
    exec(cut(fn), ns)

out = {}
for i, L in ((0, 33536), (3, 52345), (7, 8000)):
    x, _, _ = make_pair(i, L)
    k = "p%d_%d" % (i, L)
    out[k + "/vad"] = ns["get_vad"](x, 400, 200, "hanning", 40)
    sp = ns["stft"](x, 400, 200, "hanning")
    out[k + "/stft_rows"] = sp[[0, 5, sp.shape[0] - 1]]
    out[k + "/nframes"] = np.int64(sp.shape[0])
grid = np.linspace(-5, 150, 32)
out["map/grid"] = grid
out["map/siib"] = ns["mapping_SIIB_harvard"](grid)
out["map/haspi"] = ns["mapping_HASPI_harvard"](grid / 10)
out["map/estoi"] = ns["mapping_ESTOI_harvard"](grid / 150)
np.savez_compressed(os.path.join(HERE, "intel_helpers.npz"), **out)
print("wrote", len(out), "arrays")
