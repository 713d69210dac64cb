"""Development probe: per-kernel CUDA-event times of one batch (not the benchmark).
usage: kernel_times.py n L [metrics] [unique]"""
import sys
sys.path.insert(0, ".")
from nele_gan_b200.engine import Engine, pack
from nele_gan_b200.synth import make_batch

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
L = int(sys.argv[2]) if len(sys.argv) > 2 else 48000
metrics = tuple(sys.argv[3].split(",")) if len(sys.argv) > 3 else ("siib", "haspi", "estoi")
unique = int(sys.argv[4]) if len(sys.argv) > 4 else 32
refs, degs = make_batch(n, L, unique=unique)
fr, offs, lens = pack(refs)
fd, _, _ = pack(degs)
e = Engine(0)
e.set_profiling(True)
for it in range(3):
    r = e.score_packed(fr, fd, offs, lens, metrics=metrics, mapped=False, seed=1)
ms, nl = e.last_timing()
kt = e.kernel_times()
print("n=%d L=%d: %.1f ms, %d launches -> %.0f audio-s/s; scores[0]=%s" % (n, L, ms, nl, n * L / 16000 / (ms / 1e3), r.scores[0].tolist()))
for k, (t, c) in sorted(kt.items(), key=lambda kv: -kv[1][0]):
    print("  %-22s %9.3f ms  (%d)" % (k, t, c))
