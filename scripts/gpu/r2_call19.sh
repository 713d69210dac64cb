#!/bin/bash
# round 2, call 19: do concurrent metric pipelines (NELE_CONCURRENT=1) help the general-case step?
mkdir -p gpurun_out
O=gpurun_out/r2c19
for c in 0 1; do
NELE_CONCURRENT=$c timeout 600 python - > ${O}_concurrent_$c.txt 2>&1 <<'PY'
import sys, time
sys.path.insert(0, ".")
import numpy as np
from nele_gan_b200.engine import Engine, pack
from nele_gan_b200.synth import make_batch
for n, L in ((4096, 47999), (4096, 48000), (135, 44001)):
    refs, degs = make_batch(n, L, unique=32)
    fr, offs, lens = pack(refs); fd, _, _ = pack(degs)
    e = Engine(0)
    for it in range(4):
        t = time.perf_counter()
        r = e.score_packed(fr, fd, offs, lens, mapped=True, seed=1)
        dt = time.perf_counter() - t
    print("n=%d L=%d: kernels %.1f ms, wall %.1f ms" % (n, L, e.last_timing()[0], dt * 1e3), flush=True)
    del e
PY
echo "NELE_CONCURRENT=$c"; cat ${O}_concurrent_$c.txt
done
