"""Host-side mirror of the reference's metric interface, on the CUDA engine.

Same names, argument order, defaults and return types as the reference
callables this path replaces:

  pyHASPI/pyhaspi2.py:76     haspi_v2(x, fx, y, fy, HL=np.zeros(6)) -> (Intel, raw[10])
  pysiib (intel.py:77,100)   SIIB(x, y, fs, gauss=True)            -> float
  pystoi (intel.py:126,133)  stoi(x, y, fs, extended=True)         -> float
  intel.py:57-140            {SIIB,HASPI,ESTOI}_Wrapper[_raw]_harvard(x, y, fs), mapping_*_harvard
  audio_util.py:120-203      read_batch_{STOI,SIIB,HASPI}(clean_root, noise_root, enhanced_list, norm=True)
  audio_util.py:267-321      read_batch_{STOI,SIIB,HASPI}_DRC(clean_root, noise_root, enhanced_list)

Every call goes through ``libnele_score.so`` (include/nele_score.h); there is
no numpy implementation of any metric here, and nothing under ``oracle/`` is
imported.  The per-pair functions are thin (one pair per engine call); the
``read_batch_*`` functions and :func:`score_batch` score a whole list in one
call, which is what the engine is built for.
"""
import os
import warnings

import numpy as np

from . import engine as _eng

__all__ = ["haspi_v2", "haspi", "hasqi_v2", "SIIB", "stoi", "score_batch", "score_tensors", "read_batch_all",
           "check_status", "MetricError", "SiibTooShort", "BadRate",
           "SIIB_Wrapper_harvard", "SIIB_Wrapper_raw_harvard", "mapping_SIIB_harvard",
           "HASPI_Wrapper_harvard", "HASPI_Wrapper_raw_harvard", "mapping_HASPI_harvard",
           "ESTOI_Wrapper_harvard", "ESTOI_Wrapper_raw_harvard", "mapping_ESTOI_harvard",
           "read_batch_STOI", "read_batch_SIIB", "read_batch_HASPI",
           "read_batch_STOI_DRC", "read_batch_SIIB_DRC", "read_batch_HASPI_DRC"]

fs = 16000  # audio_util.py:10


def _engine():
    return _eng.default_engine()


def _f32(x):
    return np.ascontiguousarray(np.asarray(x), dtype=np.float32)


# ------------------------------------------------------------------ HASPI
def haspi_v2(x, fx, y, fy, HL=np.zeros(6), seed=None):
    """pyhaspi2.py:76-107.  Returns ``(Intel, raw)`` with ``raw`` the ten
    modulation-band correlations.  The reference draws its 0.1 dB cepstral
    dither from the global numpy stream (pyhaspi2.py:362-365); here it comes
    from a counter-based generator keyed by ``seed`` (default: drawn from
    ``np.random`` so that ``np.random.seed`` makes calls reproducible, as it
    does for the reference)."""
    if fx != fy:
        # the reference raises here too: it trims x and y to the same sample count, resamples each from its own
        # rate and then fails in numpy ("operands could not be broadcast together", pyhaspi2.py:1185-1230)
        raise ValueError("haspi_v2: fx and fy must be equal (got %r, %r)" % (fx, fy))
    if fx > 24000:
        raise NotImplementedError  # pyhaspi2.py:819-820
    x, y = _f32(x), _f32(y)
    L = min(len(x), len(y))
    if seed is None:
        seed = int(np.random.randint(0, 2 ** 31 - 1))
    hl = np.asarray(HL, dtype=np.float64)
    r = _engine().score_batch([x[:L]], [y[:L]], fs=int(fx), metrics=("haspi",), mapped=False, seed=seed,
                              hl=None if not hl.any() else hl)
    if r.metric_status("haspi")[0] == _eng.ST_BELOW_THR:
        raise Exception('Function ebm_CepCoef: Signal below threshold')  # pyhaspi2.py:357-358
    return np.float64(r.haspi[0]), r.haspi_raw[0].copy()


def haspi(x, fx, y, fy, HL=np.zeros(6), alpha=-1.0, seed=None):
    """pyhaspi2.py:109-157 (HASPI version 1).  Returns ``(Intel, raw)`` with
    ``raw = [CepCorr, cov3_low, cov3_mid, cov3_high]``.  Not on the NELE-GAN
    labelling path (intel.py calls haspi_v2 only) but part of the module's
    interface.  The basilar-membrane threshold noise the reference draws from
    the global numpy stream (pyhaspi2.py:1091-1095) comes from the engine's
    counter-based generator keyed by ``seed`` (default: drawn from
    ``np.random``, so ``np.random.seed`` makes calls reproducible)."""
    if fx != fy:
        # the reference raises here too: it trims x and y to the same sample count, resamples each from its own
        # rate and then fails in numpy ("operands could not be broadcast together", pyhaspi2.py:1185-1230)
        raise ValueError("haspi: fx and fy must be equal (got %r, %r)" % (fx, fy))
    if fx > 24000:
        raise NotImplementedError  # pyhaspi2.py:819-820
    x, y = _f32(x), _f32(y)
    L = min(len(x), len(y))
    if seed is None:
        seed = int(np.random.randint(0, 2 ** 31 - 1))
    hl = np.asarray(HL, dtype=np.float64)
    r = _engine().score_batch([x[:L]], [y[:L]], fs=int(fx), metrics=("haspi",), mapped=False, seed=seed,
                              hl=None if not hl.any() else hl, haspi_v1=True)
    if r.metric_status("haspi")[0] == _eng.ST_BELOW_THR:
        raise Exception('Function eb_melcor: Signal below threshold, outputs set to 0.')  # pyhaspi2.py:722-723
    raw = r.haspi_raw[0, :4].copy()
    arg = -9.047 + 14.816 * raw[0] + 4.616 * raw[3]                       # pyhaspi2.py:146-149
    return np.float64(1.0 / (1.0 + np.exp(alpha * arg))), raw            # :152


def hasqi_v2(x, fx, y, fy, HL=np.zeros(6), seed=None):
    """pyhaspi2.py:32-74 (HASQI version 2).  Returns ``(Combined, Nonlin, Linear, raw)`` with
    ``raw = [CepCorr, BMsync5, Dloud, Dslope]``.  Not called by NELE-GAN; same ear model and
    noise convention as :func:`haspi`."""
    if fx != fy:
        # the reference raises here too: it trims x and y to the same sample count, resamples each from its own
        # rate and then fails in numpy ("operands could not be broadcast together", pyhaspi2.py:1185-1230)
        raise ValueError("hasqi_v2: fx and fy must be equal (got %r, %r)" % (fx, fy))
    if fx > 24000:
        raise NotImplementedError  # pyhaspi2.py:819-820
    x, y = _f32(x), _f32(y)
    L = min(len(x), len(y))
    if seed is None:
        seed = int(np.random.randint(0, 2 ** 31 - 1))
    hl = np.asarray(HL, dtype=np.float64)
    r = _engine().score_batch([x[:L]], [y[:L]], fs=int(fx), metrics=("haspi",), mapped=False, seed=seed,
                              hl=None if not hl.any() else hl, hasqi=True)
    if r.metric_status("haspi")[0] == _eng.ST_BELOW_THR:
        raise Exception('Function eb_melcor: Signal below threshold, outputs set to 0.')  # pyhaspi2.py:722-723
    raw = r.haspi_raw[0]
    return np.float64(r.haspi[0]), np.float64(raw[4]), np.float64(raw[5]), [raw[0], raw[1], raw[2], raw[3]]


# ------------------------------------------------------------------- SIIB
def SIIB(x, y, fs_signal, gauss=False, use_MI_Kraskov=True, window_length=400, window_shift=200,
         window='hanning', delta_dB=40.0):
    """pysiib.SIIB.  NELE-GAN always passes ``gauss=True`` (intel.py:77,100),
    SIIB^Gauss; ``gauss=False`` (pysiib's default) runs the k-nearest-neighbour
    (Kraskov) estimator on the same KLT.  No tiling here -- that is the
    wrapper's job (intel.py:71-75) -- so fewer than 20 s of active speech raises
    like pysiib does."""
    if not gauss and not use_MI_Kraskov:
        raise NotImplementedError("only the Kraskov k-NN estimator is built for gauss=False")
    if (window_length, window_shift, window, delta_dB) != (400, 200, 'hanning', 40.0):
        raise NotImplementedError("only pysiib's default analysis parameters are supported")
    x, y = _f32(x), _f32(y)
    if x.shape != y.shape:
        raise ValueError('x and y should have the same length')
    r = _engine().score_batch([x], [y], fs=int(fs_signal), metrics=("siib",), mapped=False, siib_no_tile=True,
                              siib_knn=not gauss)
    st = r.metric_status("siib")[0]
    if st == _eng.ST_TOO_SHORT:
        raise ValueError('stimuli must have at least 20 seconds of speech')
    if st == _eng.ST_BAD_RATE:
        raise NotImplementedError("SIIB: only 16 kHz input is supported (audio_util.py:131 asserts it)")
    if st == _eng.ST_UNSUPPORTED:
        raise NotImplementedError("SIIB k-NN estimator: more than 16384 frames after stacking")
    return float(r.siib[0])


# ------------------------------------------------------------------ ESTOI
def stoi(x, y, fs_sig, extended=False):
    """pystoi.stoi.  NELE-GAN always passes ``extended=True`` (intel.py:126,133);
    classic STOI (``extended=False``, pystoi's default) runs on the same pipeline
    with the clipped-correlation back-end."""
    x, y = np.asarray(x), np.asarray(y)
    if x.shape != y.shape:
        raise Exception('x and y should have the same length,' + 'found {} and {}'.format(x.shape, y.shape))
    r = _engine().score_batch([_f32(x)], [_f32(y)], fs=int(fs_sig), metrics=("estoi",), mapped=False,
                              stoi_classic=not extended)
    if r.metric_status("estoi")[0] == _eng.ST_TOO_SHORT:
        warnings.warn('Not enough STFT frames to compute intermediate intelligibility measure after removing '
                      'silent frames. Returning 1e-5. Please check you wav files', RuntimeWarning)
    return float(r.estoi[0])


# --------------------------------------------------- intel.py wrappers
def mapping_SIIB_harvard(x):      # intel.py:102-106
    return 1 / (1 + np.exp(-0.06 * (x - 32)))


def mapping_HASPI_harvard(x):     # intel.py:116-120
    return 1 / (1 + np.exp(-0.95 * (x - 2.8)))


def mapping_ESTOI_harvard(x):     # intel.py:136-140
    return 1 / (1 + np.exp(-8.0 * (x - 0.25)))


class MetricError(Exception):
    """A pair of a batch could not be scored.  ``indices`` lists the failing pairs, ``metric`` and ``code`` the
    engine status (include/nele_score.h, NELE_ST_*).  Subclasses the reference's own exception types where it has
    one, so ``except ValueError`` / ``except Exception`` written against the reference still catch it."""

    def __init__(self, msg, metric, code, indices):
        super().__init__("%s (%s, %d of the batch: pairs %s%s)" % (msg, metric, len(indices), list(indices[:8]),
                                                                  " ..." if len(indices) > 8 else ""))
        self.metric, self.code, self.indices = metric, code, list(indices)


class SiibTooShort(MetricError, ValueError):      # pysiib: ValueError
    pass


class BadRate(MetricError, NotImplementedError):   # pyhaspi2.py:819-820 / audio_util.py:131 assert
    pass


def check_status(r, metrics=("siib", "haspi", "estoi"), strict=True):
    """Turn per-pair engine status into the reference's error behaviour for a batch.

    The reference's wrappers raise out of the whole ``Parallel`` call when one utterance fails --
    'Signal below threshold' (pyhaspi2.py:357-358), pysiib's 'at least 20 seconds of speech' -- and pystoi warns
    and returns 1e-5 for fewer than 30 frames.  ``strict=True`` (default of every batched entry point) does the
    same and names the failing pairs; ``strict=False`` leaves NaN in the scores and warns once per condition, so
    NaN never reaches the discriminator's training records unnoticed."""
    for m in metrics:
        st = r.metric_status(m)
        bad = np.nonzero((st != _eng.ST_OK) & (st != _eng.ST_SKIPPED))[0]
        if not len(bad):
            continue
        for code in np.unique(st[bad]):
            idx = bad[st[bad] == code]
            if m == "estoi" and code == _eng.ST_TOO_SHORT:
                warnings.warn('Not enough STFT frames to compute intermediate intelligibility measure after removing '
                              'silent frames. Returning 1e-5. Please check you wav files (pairs %s)' % list(idx[:8]),
                              RuntimeWarning)
                continue
            if code == _eng.ST_BELOW_THR:
                err = MetricError('Function ebm_CepCoef: Signal below threshold', m, int(code), idx)
            elif code == _eng.ST_TOO_SHORT:
                err = SiibTooShort('stimuli must have at least 20 seconds of speech', m, int(code), idx)
            elif code == _eng.ST_BAD_RATE:
                err = BadRate('sampling rate not supported for this metric', m, int(code), idx)
            else:
                err = MetricError('engine status %d' % int(code), m, int(code), idx)
            if strict:
                raise err
            warnings.warn("%s -- scores of these pairs are NaN" % err, RuntimeWarning)
    return r


def _one(metric, x, y, fs_, mapped):
    r = _engine().score_batch([_f32(x)], [_f32(y)], fs=int(fs_), metrics=(metric,), mapped=mapped,
                              seed=int(np.random.randint(0, 2 ** 31 - 1)))
    return check_status(r, (metric,))


def SIIB_Wrapper_raw_harvard(x, y, fs):      # intel.py:57-77 (VAD, tile to >= 25 s, SIIB^Gauss)
    return float(_one("siib", x, y, fs, False).siib[0])


def SIIB_Wrapper_harvard(x, y, fs):          # intel.py:79-100
    return float(_one("siib", x, y, fs, True).siib[0])


def HASPI_Wrapper_raw_harvard(x, y, fs):     # intel.py:112-114
    return float(_one("haspi", x, y, fs, False).haspi[0])


def HASPI_Wrapper_harvard(x, y, fs):         # intel.py:108-110
    return float(_one("haspi", x, y, fs, True).haspi[0])


def ESTOI_Wrapper_raw_harvard(x, y, fs):     # intel.py:122-127
    return float(_one("estoi", x, y, fs, False).estoi[0])


def ESTOI_Wrapper_harvard(x, y, fs):         # intel.py:129-134
    return float(_one("estoi", x, y, fs, True).estoi[0])


# ------------------------------------------------------- batched forms
def score_batch(refs, degs, fs=16000, metrics=("siib", "haspi", "estoi"), norm=True, seed=0, strict=True, **kw):
    """All labels of a list of (clean, degraded) pairs in one engine call.
    Returns ``float64[n, 3]`` in the column order {SIIB, HASPI, ESTOI} that
    train_nele.py:320-322 computes and audio_util.py:367-389 serialises.
    ``strict``: see :func:`check_status`."""
    r = _engine().score_batch(refs, degs, fs=fs, metrics=metrics, mapped=norm, seed=seed, **kw)
    return check_status(r, metrics, strict).scores


def _load16k(path):
    """``librosa.load(path, sr=16000)`` for the 16 kHz PCM WAV files the
    reference writes (train_nele.py:313) and asserts on (audio_util.py:131):
    float32 in [-1, 1), channels averaged to mono as librosa does.  A file at
    another rate is an error here (librosa would resample it; the reference's
    callers assert ``sr == 16000`` right after, audio_util.py:131,159,187)."""
    from scipy.io import wavfile
    sr, x = wavfile.read(path)
    if sr != 16000:
        raise ValueError("%s: sampling rate %d, the labelling path needs 16000 Hz (audio_util.py:131)" % (path, sr))
    if x.dtype == np.int16:
        x = x.astype(np.float32) / 32768.0
    elif x.dtype == np.int32:
        x = (x.astype(np.float64) / 2147483648.0).astype(np.float32)
    elif x.dtype == np.uint8:
        x = (x.astype(np.float32) - 128.0) / 128.0
    else:
        x = x.astype(np.float32)
    if x.ndim == 2:                                   # librosa.load(mono=True): mean over channels
        x = x.mean(axis=1, dtype=np.float32)
    return x


def _wave_name(enhanced_file, drc):
    f = enhanced_file.split('/')[-1]
    if drc:
        return f                                     # audio_util.py:268-269
    return (f.split('@')[0] if '@' in f else f[:-4]) + '.wav'   # audio_util.py:121-126


def _read_pairs(clean_root, noise_root, enhanced_list, drc):
    refs, degs = [], []
    for en in enhanced_list:
        name = _wave_name(en, drc)
        clean, noise, enh = _load16k(clean_root + name), _load16k(noise_root + name), _load16k(en)
        m = min(len(clean), len(enh))                # audio_util.py:134-137
        refs.append(clean[:m])
        degs.append(enh[:m] + noise[:m])
    return refs, degs


def _read_pcm16(clean_root, noise_root, enhanced_list, drc):
    """The three files of every utterance as int16, trimmed like audio_util.py:134-137 and packed for
    ``nele_score_batch_pcm16`` -- or None when a file is not 16-bit mono PCM at 16 kHz (then the float path runs)."""
    from scipy.io import wavfile
    trip = []
    for en in enhanced_list:
        name = _wave_name(en, drc)
        cur = []
        for path in (clean_root + name, en, noise_root + name):
            sr, x = wavfile.read(path)
            if sr != 16000 or x.dtype != np.int16 or x.ndim != 1:
                return None
            cur.append(x)
        m = min(len(cur[0]), len(cur[1]))
        if len(cur[2]) < m:
            return None
        trip.append([a[:m] for a in cur])
    lens = np.array([len(t[0]) for t in trip], dtype=np.int32)
    padded = (lens.astype(np.int64) + 7) // 8 * 8
    offs = np.concatenate(([0], np.cumsum(padded)[:-1])).astype(np.int64)
    flat = [np.zeros(int(padded.sum()), dtype=np.int16) for _ in range(3)]
    for t, o, n in zip(trip, offs, lens):
        for k in range(3):
            flat[k][o:o + n] = t[k]
    return flat, offs, lens


def _score_files(clean_root, noise_root, enhanced_list, drc, metrics, norm, seed):
    """One engine call for a list of files: int16 upload with ``enhanced + noise`` formed on the device when the
    files are 16-bit PCM (what train_nele.py:313 writes), else the float path."""
    pcm = _read_pcm16(clean_root, noise_root, enhanced_list, drc)
    if pcm is not None:
        flat, offs, lens = pcm
        return _engine().score_packed_pcm16(flat[0], flat[1], flat[2], offs, lens, fs=fs, metrics=metrics, mapped=bool(norm),
                                            seed=seed)
    refs, degs = _read_pairs(clean_root, noise_root, enhanced_list, drc)
    return _engine().score_batch(refs, degs, fs=fs, metrics=metrics, mapped=bool(norm), seed=seed)


def _read_batch(metric, col, clean_root, noise_root, enhanced_list, norm, drc, strict=True):
    if not len(enhanced_list):
        return []
    r = _score_files(clean_root, noise_root, list(enhanced_list), drc, (metric,), norm, int(np.random.randint(0, 2 ** 31 - 1)))
    check_status(r, (metric,), strict)
    return [float(v) for v in r.scores[:, col]]


def read_batch_STOI(clean_root, noise_root, enhanced_list, norm=True):      # audio_util.py:145-147
    return _read_batch("estoi", _eng.COL_ESTOI, clean_root, noise_root, enhanced_list, norm, False)


def read_batch_SIIB(clean_root, noise_root, enhanced_list, norm=True):      # audio_util.py:173-175
    return _read_batch("siib", _eng.COL_SIIB, clean_root, noise_root, enhanced_list, norm, False)


def read_batch_HASPI(clean_root, noise_root, enhanced_list, norm=True):     # audio_util.py:201-203
    return _read_batch("haspi", _eng.COL_HASPI, clean_root, noise_root, enhanced_list, norm, False)


def read_batch_STOI_DRC(clean_root, noise_root, enhanced_list):             # audio_util.py:281-283
    return _read_batch("estoi", _eng.COL_ESTOI, clean_root, noise_root, enhanced_list, True, True)


def read_batch_SIIB_DRC(clean_root, noise_root, enhanced_list):             # audio_util.py:300-302
    return _read_batch("siib", _eng.COL_SIIB, clean_root, noise_root, enhanced_list, True, True)


def read_batch_HASPI_DRC(clean_root, noise_root, enhanced_list):            # audio_util.py:319-321
    return _read_batch("haspi", _eng.COL_HASPI, clean_root, noise_root, enhanced_list, True, True)


def read_batch_all(clean_root, noise_root, enhanced_list, norm=True, drc=False, seed=None, strict=True):
    """The three read_batch_* calls of one sampling round (train_nele.py:320-322
    or :333-335) fused: the WAV files are read once and scored in one engine
    call.  Returns ``(siib, haspi, estoi)`` lists."""
    if not len(enhanced_list):
        return [], [], []
    if seed is None:
        seed = int(np.random.randint(0, 2 ** 31 - 1))
    s = check_status(_score_files(clean_root, noise_root, list(enhanced_list), drc, ("siib", "haspi", "estoi"), norm, seed),
                     strict=strict).scores
    return [float(v) for v in s[:, 0]], [float(v) for v in s[:, 1]], [float(v) for v in s[:, 2]]


def score_tensors(ref, deg, lengths=None, fs=16000, metrics=("siib", "haspi", "estoi"), norm=True, seed=0,
                  pcm16=False, strict=True, **kw):
    """In-loop tensor boundary (train_nele.py:303-322 without the WAV round trip):
    ``ref`` / ``deg`` are CUDA float32 torch tensors ``[n, Lmax]`` (clean, and
    enhanced + noise as audio_util.py:139 forms it), ``lengths`` the valid
    samples per row (default: all of ``Lmax``).  The waveforms stay on the
    device -- the engine reads them through ``data_ptr()`` -- and only the
    per-pair records come back.  ``pcm16=True`` reproduces what the reference's
    ``sf.write(..., 'PCM_16')`` + ``librosa.load`` round trip (train_nele.py:313,
    audio_util.py:186-189) does to the degraded signal's *enhanced* component;
    here it is applied to ``deg`` as a whole.  Returns a float64 torch tensor
    ``[n, 3]`` = {SIIB, HASPI, ESTOI} on the CPU."""
    import torch
    if not (ref.is_cuda and deg.is_cuda):
        raise ValueError("score_tensors needs CUDA tensors (the engine has no CPU path)")
    if ref.shape != deg.shape or ref.dim() != 2:
        raise ValueError("ref and deg must both be [n, Lmax]")
    ref = ref.detach().to(torch.float32).contiguous()
    deg = deg.detach().to(torch.float32).contiguous()
    if pcm16:
        deg = torch.clamp(torch.round(deg * 32768.0), -32768.0, 32767.0) / 32768.0
    n, lmax = ref.shape
    if lmax % 4:  # rows must start 16-byte aligned
        pad = 4 - lmax % 4
        ref = torch.nn.functional.pad(ref, (0, pad))
        deg = torch.nn.functional.pad(deg, (0, pad))
    stride = ref.shape[1]
    lens = np.full(n, lmax, dtype=np.int32) if lengths is None else np.asarray(lengths, dtype=np.int32)
    if lens.shape != (n,) or lens.min() <= 0 or lens.max() > lmax:
        raise ValueError("lengths must be n values in (0, Lmax]")
    offs = np.arange(n, dtype=np.int64) * stride
    eng = _eng.default_engine(ref.device.index)
    torch.cuda.current_stream(ref.device).synchronize()   # the engine runs on its own stream
    r = eng.score_packed(ref.data_ptr(), deg.data_ptr(), offs, lens, fs=fs, metrics=metrics, mapped=bool(norm),
                         seed=seed, device_input=True, **kw)
    return torch.from_numpy(check_status(r, metrics, strict).scores)


def dropin_path():
    """Directory to put in front of ``sys.path`` so that the reference's own
    ``intel.py`` / ``audio_util.py`` import the engine-backed ``pyHASPI.pyhaspi2``,
    ``pysiib`` and ``pystoi.stoi``."""
    return os.path.join(os.path.dirname(os.path.abspath(__file__)), "dropin")
