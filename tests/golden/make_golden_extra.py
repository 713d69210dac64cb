"""Generate ``tests/golden/haspi_extra.npz`` with the UNMODIFIED reference ``pyhaspi2.py`` (as make_golden.py):

  * ``hl/...``    haspi_v2 with a non-zero audiogram HL = [20, 25, 35, 45, 55, 60] dB on the toy pair
                  ``toy_train_multienh`` of haspi_ref.npz, every randn forced to zero: score, raw, BWx / BWy
                  (pyHASPI/README.txt:14 vouches for HL = 0 only; the engine still has to follow the code);
  * ``seeds/...`` haspi_v2 on ``toy_train_clean`` exactly as a user runs it, after ``np.random.seed(s)``,
                  s = 0 .. 31: the distribution the engine's own Philox dither has to reproduce in the mean
                  (the seed-to-seed spread is larger than the 1e-3 tolerance).

Run in the build container only:  python tests/golden/make_golden_extra.py"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import make_golden as MG  # noqa: E402  (sets up sys.path, imports the reference)

HL = np.array([20.0, 25.0, 35.0, 45.0, 55.0, 60.0])
N_SEEDS = 32


def main():
    z = np.load(os.path.join(HERE, "haspi_ref.npz"))
    out = {"hl/HL": HL, "hl/case": np.array("toy_train_multienh"), "seeds/case": np.array("toy_train_clean")}
    x, y = z["toy_train_multienh/x"], z["toy_train_multienh/y"]
    s, raw, rec = MG.run_ref(lambda a, fa, b, fb: MG.REF.haspi_v2(a, fa, b, fb, HL), x, 16000, y, 16000, "zero")
    bw = np.asarray(rec["bw"], dtype=np.float64).reshape(32, 2)
    out["hl/v2_zero"], out["hl/v2_zero_raw"] = np.float64(s), raw
    out["hl/bwx"], out["hl/bwy"], out["hl/nsel"] = bw[:, 0].copy(), bw[:, 1].copy(), np.int64(rec["nsel"])
    print("HL", HL.tolist(), "->", s, "(HL = 0:", float(z["toy_train_multienh/v2_zero"]), ")", flush=True)
    x, y = z["toy_train_clean/x"], z["toy_train_clean/y"]
    vals = []
    for sd in range(N_SEEDS):
        np.random.seed(sd)
        v, _ = MG.REF.haspi_v2(x, 16000, y, 16000)
        vals.append(float(v))
        print("seed", sd, v, flush=True)
    out["seeds/v2"] = np.asarray(vals)
    print("mean %.6f  std %.2e  (zero dither: %.6f)" % (np.mean(vals), np.std(vals), float(z["toy_train_clean/v2_zero"])))
    np.savez_compressed(os.path.join(HERE, "haspi_extra.npz"), **out)


if __name__ == "__main__":
    main()
