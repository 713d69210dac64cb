"""Generate ``tests/golden/hasqi_ref.npz``: outputs of the UNMODIFIED reference ``hasqi_v2``
(pyhaspi2.py:32-74) for the cases of ``make_golden.py`` (inputs are in ``haspi_ref.npz``), with
every ``np.random.randn`` forced to zero (no BM threshold noise), plus ``eb_aveSL``'s returns
(xSL, ySL) recorded by wrapping the reference's own function.  Build container only."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import make_golden as MG  # noqa: E402  (sets sys.path for the reference and the librosa shim)

REF = MG.REF


def main():
    out = {}
    for name, x, y, fs in MG.cases():
        x = np.ascontiguousarray(x, dtype=np.float32)
        y = np.ascontiguousarray(y, dtype=np.float32)
        sl = []
        orig = (np.random.randn, REF.eb_aveSL)

        def sl_wrap(*a, **k):
            v = orig[1](*a, **k)
            sl.append(np.asarray(v, dtype=np.float64))
            return v

        try:
            np.random.randn = MG.PatchedRandn("zero")
            REF.eb_aveSL = sl_wrap
            comb, nonlin, lin, raw = REF.hasqi_v2(x, fs, y, fs)
        finally:
            np.random.randn, REF.eb_aveSL = orig
        out[name + "/hq_zero"] = np.array([comb, nonlin, lin], dtype=np.float64)
        out[name + "/hq_zero_raw"] = np.asarray(raw, dtype=np.float64)
        out[name + "/xsl"] = sl[0]
        out[name + "/ysl"] = sl[1]
        print(name, comb, nonlin, lin, raw, flush=True)
    np.savez_compressed(os.path.join(HERE, "hasqi_ref.npz"), **out)


if __name__ == "__main__":
    main()
