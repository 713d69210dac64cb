"""Wire format of the score records (nele_gan_b200/records.py) against the reference's own
string functions, taken from the unmodified source text of audio_util.py:367-389."""
import os
import re

import numpy as np
import pytest

REF = "/root/reference/audio_util.py"


def _reference_functions():
    src = open(REF).read()
    ns = {}
    for name in ("List_concat", "List_concat_score", "List_concat_3scores", "List_concat_5scores"):
        m = re.search(r"^def %s\(.*?(?=^def )" % name, src, re.S | re.M)
        exec(m.group(0), ns)
    return ns


def test_records_round_trip():
    from nele_gan_b200 import records
    s = np.array([[0.53125, 0.4732119, 1e-5], [0.9, 0.25, 0.125]])
    names = ["/tmp/out/spk_hvd_001#Cafe#-9@3.wav", "/tmp/out/spk_hvd_002#Cafe#-9@3.wav"]
    rec = records.round_records(s, names, pesq=[0.7, 0.6], visqol=[0.5, 0.4])
    assert rec[0] == "0.53125,0.4732119,1e-05,0.7,0.5," + names[0]
    ts, tq, paths = records.records_to_tensors(rec)
    assert ts.dtype == np.float32 and ts.shape == (2, 3) and tq.shape == (2, 2) and paths == names
    assert np.allclose(ts, s.astype(np.float32)) and np.allclose(tq, [[0.7, 0.5], [0.6, 0.4]])


@pytest.mark.skipif(not os.path.exists(REF), reason="reference tree not mounted (GPU box)")
def test_same_strings_as_the_reference_functions():
    from nele_gan_b200 import records
    ref = _reference_functions()
    a, b, c, d, e = ([0.1 * k + 0.01 * j for k in range(4)] for j in range(5))
    names = ["p%d@1.wav" % k for k in range(4)]
    assert records.List_concat_5scores(a, b, c, d, e) == ref["List_concat_5scores"](a, b, c, d, e)
    assert records.List_concat_3scores(a, b, c) == ref["List_concat_3scores"](a, b, c)
    assert records.List_concat_score(a, b) == ref["List_concat_score"](a, b)
    assert records.List_concat(a, names) == ref["List_concat"](a, names)
