"""Host logic of nele_gan_b200.api that needs no GPU: how per-pair engine status becomes the
reference's error behaviour in the batched entry points (NaN must not reach the
discriminator's records silently), the fx != fy contract, and WAV loading."""
import os
import warnings

import numpy as np
import pytest

from nele_gan_b200 import api, engine


def _result(hst, sst, est, info=0):
    n = len(hst)
    status = (np.asarray(hst) | (np.asarray(sst) << 8) | (np.asarray(est) << 16) | info).astype(np.int32)
    return engine.BatchResult(np.full((n, 3), np.nan), np.zeros((n, 10)), status)


def test_ok_batch_passes_and_info_bit_is_not_an_error():
    r = _result([0, 0], [0, 0], [0, 0], info=engine.INFO_SIIB_NULLSPACE)
    assert api.check_status(r) is r
    assert r.ok.all() and r.siib_nullspace_dropped.all()
    assert (r.metric_status("siib") == 0).all()


def test_haspi_below_threshold_raises_like_the_reference():
    r = _result([0, engine.ST_BELOW_THR, 0], [0, 0, 0], [0, 0, 0])
    with pytest.raises(Exception, match="Signal below threshold") as ei:      # pyhaspi2.py:357-358
        api.check_status(r)
    assert ei.value.indices == [1] and ei.value.metric == "haspi"
    with pytest.warns(RuntimeWarning, match="NaN"):
        api.check_status(r, strict=False)


def test_siib_too_short_is_a_valueerror_and_bad_rate_notimplemented():
    r = _result([0, 0], [engine.ST_TOO_SHORT, 0], [0, 0])
    with pytest.raises(ValueError, match="at least 20 seconds"):              # pysiib
        api.check_status(r)
    r = _result([0, 0], [engine.ST_BAD_RATE, engine.ST_BAD_RATE], [0, 0])
    with pytest.raises(NotImplementedError) as ei:
        api.check_status(r)
    assert ei.value.indices == [0, 1]
    # a metric that was not requested is skipped, not an error
    api.check_status(_result([engine.ST_SKIPPED] * 2, [0, 0], [engine.ST_SKIPPED] * 2))


def test_estoi_too_short_only_warns():
    r = _result([0], [0], [engine.ST_TOO_SHORT])
    with pytest.warns(RuntimeWarning, match="Returning 1e-5"):                # pystoi's sentinel + warning
        api.check_status(r)


def test_fx_ne_fy_is_a_valueerror_as_in_the_reference():
    # the unmodified pyhaspi2.haspi_v2(x, 16000, y, 22050) ends in numpy's "operands could not be broadcast together"
    # ValueError (x and y are trimmed to equal sample counts, then resampled from different rates)
    x = np.zeros(100, np.float32)
    for fn in (api.haspi_v2, api.haspi, api.hasqi_v2):
        with pytest.raises(ValueError):
            fn(x, 16000, x, 22050)


def test_load16k_downmixes_and_rejects_other_rates(tmp_path):
    from scipy.io import wavfile
    st = (np.arange(2000, dtype=np.int16).reshape(1000, 2) - 500)
    p = str(tmp_path / "st.wav")
    wavfile.write(p, 16000, st)
    x = api._load16k(p)
    assert x.dtype == np.float32 and x.shape == (1000,)
    assert np.allclose(x, st.astype(np.float32).mean(axis=1) / 32768.0)
    p2 = str(tmp_path / "r.wav")
    wavfile.write(p2, 22050, st[:, 0].copy())
    with pytest.raises(ValueError, match="16000"):
        api._load16k(p2)
