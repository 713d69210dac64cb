import sys; sys.path.insert(0,".")
from nele_gan_b200.engine import Engine, pack
from nele_gan_b200.synth import make_batch
refs, degs = make_batch(256, 47999, unique=32)
fr, offs, lens = pack(refs); fd,_,_ = pack(degs)
e = Engine(0); e.set_profiling(True)
for it in range(2):
    r = e.score_packed(fr, fd, offs, lens, metrics=("siib",), mapped=False, siib_knn=True)
print(e.last_timing(), r.scores[:3,0].tolist())
for k,(t,c) in sorted(e.kernel_times().items(), key=lambda kv:-kv[1][0])[:6]: print("  %-20s %9.3f ms (%d)"%(k,t,c))
