"""Development timing probe (not the benchmark): kernel time of one batch."""
import sys, time
sys.path.insert(0, ".")
import numpy as np
from nele_gan_b200.engine import Engine, pack
from nele_gan_b200.synth import make_batch

n = int(sys.argv[1]) if len(sys.argv) > 1 else 512
sec = float(sys.argv[2]) if len(sys.argv) > 2 else 3.0
metrics = tuple(sys.argv[3].split(",")) if len(sys.argv) > 3 else ("haspi",)
refs, degs = make_batch(n, int(sec * 16000), unique=16)
fr, offs, lens = pack(refs); fd, _, _ = pack(degs)
e = Engine(0)
for it in range(3):
    t = time.time()
    r = e.score_packed(fr, fd, offs, lens, metrics=metrics, mapped=False)
    dt = time.time() - t
    ms, nl = e.last_timing()
    print("iter", it, "wall %.1f ms kernel %.1f ms launches %d -> %.0f audio-s/s (kernel)" % (dt * 1e3, ms, nl, n * sec / (ms / 1e3)),
          "scores", r.scores[:2].tolist(), flush=True)
