import sys; sys.path.insert(0, ".")
import numpy as np
from nele_gan_b200.engine import Engine
from nele_gan_b200.synth import make_pair
from oracle import intel_np
e = Engine(0)
x, y, _ = make_pair(21, 16000 * 30 + 123)
x1, y1, _ = make_pair(22, 16000)
x2, y2, _ = make_pair(23, 40000)
z = np.zeros(32000, np.float32)
refs = [x, x1, x2, z, x2, (x2 * 1e-6).astype(np.float32)]
degs = [y, y1, x2, z, (y2 * 50).astype(np.float32), (y2 * 1e-6).astype(np.float32)]
r = e.score_batch(refs, degs, mapped=False, no_dither=True)
print(r.scores); print([hex(s) for s in r.status])
for i in (0, 1, 2, 4, 5):
    try:
        w = intel_np.score_pair(refs[i], degs[i], 16000, norm=False, noise=None)
        print(i, "oracle", w, "rel siib %.2e dhaspi %.2e destoi %.2e" % (abs(r.scores[i,0]-w[0])/max(abs(w[0]),1e-9), abs(r.scores[i,1]-w[1]), abs(r.scores[i,2]-w[2])))
    except Exception as ex:
        print(i, "oracle raised", repr(ex)[:100])
