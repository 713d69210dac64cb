"""BASELINE.json configs[4] in small: n pairs of 3-10 s (lengths 16000 * U[3, 10] from
default_rng(12345), content as configs[2]), all three metrics, one engine call from host buffers."""
import sys, time
sys.path.insert(0, ".")
import numpy as np
from nele_gan_b200.engine import Engine, pack
from nele_gan_b200.synth import make_batch

n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
lens = (16000 * np.random.default_rng(12345).uniform(3, 10, size=n)).astype(np.int64)
refs, degs = make_batch(n, lens, unique=64)
fr, offs, ln = pack(refs); fd, _, _ = pack(degs)
e = Engine(0); e.set_profiling(True)
for it in range(2):
    t = time.perf_counter()
    r = e.score_packed(fr, fd, offs, ln, mapped=True, seed=1)
    dt = time.perf_counter() - t
ms, nl = e.last_timing()
sec = float(ln.sum()) / 16000
ok = int(np.sum((r.status & 0xFFFFFF) == 0))
print("n=%d, %.0f audio-s: wall %.1f ms, kernels %.1f ms, %d launches -> %.0f audio-s/s end to end, %d pairs ok" % (n, sec, dt * 1e3, ms, nl, sec / dt, ok))
for k, (tm, c) in sorted(e.kernel_times().items(), key=lambda kv: -kv[1][0])[:8]:
    print("  %-22s %9.3f ms  (%d)" % (k, tm, c))
print("scores[:3] =", r.scores[:3].tolist())
