"""Profiling target: two engine calls on a synthetic batch (first = warm-up)."""
import sys
sys.path.insert(0, ".")
from nele_gan_b200.engine import Engine, pack
from nele_gan_b200.synth import make_batch

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
L = int(sys.argv[2]) if len(sys.argv) > 2 else 48000
calls = int(sys.argv[3]) if len(sys.argv) > 3 else 2
refs, degs = make_batch(n, L, unique=32)
fr, offs, lens = pack(refs)
fd, _, _ = pack(degs)
e = Engine(0)
for it in range(calls):
    r = e.score_packed(fr, fd, offs, lens, mapped=True, seed=1)
    print(it, e.last_timing(), r.scores[0].tolist(), flush=True)
