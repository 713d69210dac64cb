"""Development probe: end-to-end time of one 4096 x 3 s call from pinned host buffers."""
import sys, time
sys.path.insert(0, ".")
import numpy as np, torch
from nele_gan_b200.engine import Engine, pack
from nele_gan_b200.synth import make_batch
refs, degs = make_batch(4096, 48000, unique=64)
fr, offs, lens = pack(refs); fd, _, _ = pack(degs)
hr, hd = torch.from_numpy(fr).pin_memory(), torch.from_numpy(fd).pin_memory()
e = Engine(0)
for it in range(5):
    torch.cuda.synchronize(); t = time.perf_counter()
    r = e.score_packed(hr.data_ptr(), hd.data_ptr(), offs, lens, mapped=True, seed=1)
    torch.cuda.synchronize(); dt = time.perf_counter() - t
    print("iter %d: e2e %.1f ms, kernels %.1f ms" % (it, dt * 1e3, e.last_timing()[0]), flush=True)
