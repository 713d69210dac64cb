"""A/B of the kernel switches in one go: for each configuration a fresh process
(the switches are read once per process) scores the same batch, prints per-kernel CUDA-event times and the
largest score deviation from the default configuration.
usage: ab_switches.py [n] [L]      (needs a B200; e.g. 4096 48000 for the bench workload, 1024 52345 for full rank)"""
import json
import os
import subprocess
import sys

# every configuration is one set of environment switches of DESIGN.md section 9 (A/B references of the round-2 defaults)
CONFIGS = [("default", {}), ("ear_scalar", {"NELE_F32X2": "0"}), ("resample_f64", {"NELE_RESAMPLE_F32": "0"}),
           ("tridiag_f64", {"NELE_TRIDIAG_F64": "1"}), ("backtf_old", {"NELE_BACKTF_OLD": "1"}),
           ("backtf4", {"NELE_BACKTF4": "1"}), ("backtf5", {"NELE_BACKTF5": "1"}), ("cov_xx64", {"NELE_COV_XX64": "1"})]

CHILD = r'''
import json, sys
sys.path.insert(0, ".")
import numpy as np
from nele_gan_b200.engine import Engine, pack
from nele_gan_b200.synth import make_batch
n, L = int(sys.argv[1]), int(sys.argv[2])
refs, degs = make_batch(n, L, unique=32)
fr, offs, lens = pack(refs)
fd, _, _ = pack(degs)
e = Engine(0)
e.set_profiling(True)
for it in range(3):
    r = e.score_packed(fr, fd, offs, lens, mapped=False, no_dither=True)
ms, nl = e.last_timing()
print(json.dumps({"ms": ms, "launches": nl, "kernels": {k: v[0] for k, v in e.kernel_times().items()},
                  "scores": r.scores[:32].tolist()}))
'''

if __name__ == "__main__":
    n = sys.argv[1] if len(sys.argv) > 1 else "4096"
    L = sys.argv[2] if len(sys.argv) > 2 else "48000"
    only = sys.argv[3:]   # optional: names of the configurations to run ("default" first)
    base = None
    for name, env in [c for c in CONFIGS if not only or c[0] in only]:
        p = subprocess.run([sys.executable, "-c", CHILD, n, L], env={**os.environ, **env}, stdout=subprocess.PIPE,
                           stderr=subprocess.PIPE, text=True)
        if p.returncode != 0:
            print("%-14s FAILED: %s" % (name, p.stderr.strip().splitlines()[-1:] or "?"))
            continue
        res = json.loads(p.stdout.strip().splitlines()[-1])
        if base is None:
            base = res
        import numpy as np
        s, s0 = np.array(res["scores"]), np.array(base["scores"])
        dev = np.nanmax(np.abs(s - s0) / np.maximum(np.abs(s0), 1e-12), axis=0)
        changed = {k: (base["kernels"].get(k, 0.0), v) for k, v in res["kernels"].items()
                   if abs(v - base["kernels"].get(k, 0.0)) > 0.03 * max(v, 0.2)}
        print("%-14s %8.2f ms  (default %8.2f)  max rel score deviation {SIIB, HASPI, ESTOI} = %s" % (
            name, res["ms"], base["ms"], ["%.1e" % d for d in dev]))
        for k, (a, b) in sorted(changed.items(), key=lambda kv: -abs(kv[1][1] - kv[1][0])):
            print("    %-20s %8.3f -> %8.3f ms" % (k, a, b))
