"""ctypes binding of ``libnele_score.so`` (include/nele_score.h) and the batched
host API.

The engine has no CPU implementation: if the shared library is missing or no
CUDA device is visible, constructing :class:`Engine` raises
:class:`NeleError` -- it never falls back to the oracle or to numpy.
"""
import ctypes as C
import os
import threading

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libnele_score.so")

METRIC_HASPI, METRIC_SIIB, METRIC_ESTOI = 1, 2, 4
METRIC_ALL = 7
FLAG_MAPPED, FLAG_DEVICE_INPUT, FLAG_NO_DITHER, FLAG_SIIB_NO_TILE, FLAG_KEEP_STAGES, FLAG_HASPI_V1, FLAG_STOI_CLASSIC, FLAG_SIIB_KNN, FLAG_HASQI_V2 = 1, 2, 4, 8, 16, 32, 64, 128, 256
ST_OK, ST_BELOW_THR, ST_TOO_SHORT, ST_BAD_RATE, ST_UNSUPPORTED, ST_SKIPPED = 0, 1, 2, 3, 4, 0xFF
INFO_SIIB_NULLSPACE = 0x01000000   # informational bit above the three status bytes (include/nele_score.h)
_METRIC_BITS = {"haspi": METRIC_HASPI, "siib": METRIC_SIIB, "estoi": METRIC_ESTOI, "stoi": METRIC_ESTOI}
COL_SIIB, COL_HASPI, COL_ESTOI = 0, 1, 2

SYMBOLS = ("nele_abi_version", "nele_create", "nele_destroy", "nele_last_error", "nele_score_batch", "nele_prefetch",
           "nele_prefetch_cancel", "nele_score_batch_pcm16", "nele_prefetch_pcm16",
           "nele_get_stage", "nele_last_timing", "nele_set_profiling", "nele_kernel_time", "nele_feature_frames",
           "nele_features", "nele_resyn")
FEAT_NOISE, FEAT_DEVICE_IO, FEAT_NO_POWER = 0x1, 0x2, 0x4


class NeleError(RuntimeError):
    pass


_lib = None
_lib_lock = threading.Lock()


def load_library(path=None):
    """dlopen the C-ABI library and declare its prototypes."""
    global _lib
    with _lib_lock:
        if _lib is not None and path is None:
            return _lib
        p = path or LIB_PATH
        if not os.path.exists(p):
            raise NeleError("%s not found: build it with `python -m nele_gan_b200.build` "
                            "(there is no CPU fallback)" % p)
        lib = C.CDLL(p)
        lib.nele_abi_version.restype = C.c_int
        lib.nele_create.argtypes = [C.c_int, C.POINTER(C.c_void_p)]
        lib.nele_create.restype = C.c_int
        lib.nele_destroy.argtypes = [C.c_void_p]
        lib.nele_destroy.restype = None
        lib.nele_last_error.argtypes = [C.c_void_p]
        lib.nele_last_error.restype = C.c_char_p
        lib.nele_score_batch.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int,
                                         C.c_int, C.c_uint32, C.c_uint32, C.c_void_p, C.c_int64, C.c_uint64,
                                         C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        lib.nele_score_batch.restype = C.c_int
        lib.nele_prefetch.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_uint32]
        lib.nele_prefetch.restype = C.c_int
        lib.nele_score_batch_pcm16.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int,
                                               C.c_int, C.c_uint32, C.c_uint32, C.c_void_p, C.c_int64, C.c_uint64,
                                               C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        lib.nele_score_batch_pcm16.restype = C.c_int
        lib.nele_prefetch_pcm16.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int,
                                            C.c_uint32]
        lib.nele_prefetch_pcm16.restype = C.c_int
        lib.nele_prefetch_cancel.argtypes = [C.c_void_p]
        lib.nele_prefetch_cancel.restype = C.c_int
        lib.nele_get_stage.argtypes = [C.c_void_p, C.c_char_p, C.c_int, C.c_void_p, C.c_size_t,
                                       C.POINTER(C.c_size_t)]
        lib.nele_get_stage.restype = C.c_int
        lib.nele_last_timing.argtypes = [C.c_void_p, C.POINTER(C.c_double), C.POINTER(C.c_int64)]
        lib.nele_last_timing.restype = C.c_int
        lib.nele_set_profiling.argtypes = [C.c_void_p, C.c_int]
        lib.nele_set_profiling.restype = C.c_int
        lib.nele_kernel_time.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_char_p), C.POINTER(C.c_double),
                                         C.POINTER(C.c_int64)]
        lib.nele_kernel_time.restype = C.c_int
        lib.nele_feature_frames.argtypes = [C.c_int32]
        lib.nele_feature_frames.restype = C.c_int64
        lib.nele_features.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_uint32, C.c_double,
                                      C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        lib.nele_features.restype = C.c_int
        lib.nele_resyn.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p,
                                   C.c_void_p, C.c_uint32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        lib.nele_resyn.restype = C.c_int
        if path is None:
            _lib = lib
        return lib


def metric_mask(metrics):
    if isinstance(metrics, int):
        return metrics
    m = 0
    for name in metrics:
        m |= _METRIC_BITS[name.lower()]
    return m


def pack(signals, align=4):
    """Concatenate ragged 1-D float32 signals -> (flat float32, offs int64, lens int32)."""
    lens = np.array([len(s) for s in signals], dtype=np.int32)
    padded = (lens.astype(np.int64) + (align - 1)) // align * align
    offs = np.concatenate(([0], np.cumsum(padded)[:-1])).astype(np.int64) if len(lens) else np.zeros(0, np.int64)
    flat = np.zeros(int(padded.sum()), dtype=np.float32)
    for s, o, n in zip(signals, offs, lens):
        flat[o:o + n] = s
    return flat, offs, lens


class BatchResult:
    """scores[n,3] in the column order {SIIB, HASPI, ESTOI} (train_nele.py:320-322),
    haspi_raw[n,10], status[n] (one byte per metric, see include/nele_score.h)."""

    def __init__(self, scores, haspi_raw, status):
        self.scores, self.haspi_raw, self.status = scores, haspi_raw, status

    siib = property(lambda self: self.scores[:, COL_SIIB])
    haspi = property(lambda self: self.scores[:, COL_HASPI])
    estoi = property(lambda self: self.scores[:, COL_ESTOI])

    def metric_status(self, name):
        k = {"haspi": 0, "siib": 1, "estoi": 2}[name]
        return (self.status >> (8 * k)) & 0xFF

    @property
    def ok(self):
        """Per pair: every requested metric is ST_OK (informational bits ignored)."""
        st = np.stack([self.metric_status(m) for m in ("haspi", "siib", "estoi")], axis=1)
        return np.all((st == ST_OK) | (st == ST_SKIPPED), axis=1)

    @property
    def siib_nullspace_dropped(self):
        """Per pair: SIIB ignored the null space of a rank-deficient covariance -- the utterance length is a
        multiple of the 200-sample hop, so the wrapper's tiling repeats exactly and pysiib's float64 value
        carries 1-13 % of rounding-noise information."""
        return (self.status & INFO_SIIB_NULLSPACE) != 0


class Engine:
    """One scoring engine on one CUDA device."""

    def __init__(self, device=0, lib_path=None):
        self._lib = load_library(lib_path)
        h = C.c_void_p()
        rc = self._lib.nele_create(int(device), C.byref(h))
        if rc != 0:
            msg = self._lib.nele_last_error(None)
            raise NeleError("nele_create(device=%d) failed (%d): %s" % (device, rc, msg.decode() if msg else "?"))
        self._h = h
        self.device = int(device)

    def close(self):
        if getattr(self, "_h", None):
            self._lib.nele_destroy(self._h)
            self._h = None

    __del__ = close

    def _check(self, rc, what):
        if rc != 0:
            msg = self._lib.nele_last_error(self._h)
            raise NeleError("%s failed (%d): %s" % (what, rc, msg.decode() if msg else "?"))

    # ------------------------------------------------------------------ low level
    def score_packed(self, ref, deg, offs, lens, fs=16000, metrics=METRIC_ALL, mapped=True, dither=None,
                     seed=0, no_dither=False, hl=None, siib_no_tile=False, keep_stages=False,
                     device_input=False, stream=None, out=None, haspi_v1=False, stoi_classic=False, siib_knn=False, hasqi=False):
        """``ref``/``deg``: flat float32 numpy arrays, or raw pointers (ints, e.g.
        ``tensor.data_ptr()``; device pointers when ``device_input``).  ``offs``
        int64[n], ``lens`` int32[n] are host numpy arrays."""
        offs = np.ascontiguousarray(offs, dtype=np.int64)
        lens = np.ascontiguousarray(lens, dtype=np.int32)
        n = int(lens.shape[0])
        if isinstance(ref, np.ndarray):
            ref = np.ascontiguousarray(ref, dtype=np.float32)
            deg = np.ascontiguousarray(deg, dtype=np.float32)
            pref, pdeg = ref.ctypes.data, deg.ctypes.data
        else:
            pref, pdeg = int(ref), int(deg)
        flags = (FLAG_MAPPED if mapped else 0) | (FLAG_NO_DITHER if no_dither else 0) | \
                (FLAG_SIIB_NO_TILE if siib_no_tile else 0) | (FLAG_KEEP_STAGES if keep_stages else 0) | \
                (FLAG_DEVICE_INPUT if device_input else 0) | (FLAG_HASPI_V1 if haspi_v1 else 0) | \
                (FLAG_STOI_CLASSIC if stoi_classic else 0) | (FLAG_SIIB_KNN if siib_knn else 0) | \
                (FLAG_HASQI_V2 if hasqi else 0)
        pd, drows = None, 0
        if dither is not None:
            dither = np.ascontiguousarray(dither, dtype=np.float32)
            assert dither.ndim == 3 and dither.shape[0] == 2 and dither.shape[2] == 32
            pd, drows = dither.ctypes.data, dither.shape[1]
        phl = None
        if hl is not None:
            hl = np.ascontiguousarray(hl, dtype=np.float64)
            assert hl.shape == (6,)
            phl = hl.ctypes.data
        if out is None:
            scores = np.empty((n, 3), dtype=np.float64)
            raw = np.empty((n, 10), dtype=np.float64)
            status = np.empty(n, dtype=np.int32)
        else:
            scores, raw, status = out
        rc = self._lib.nele_score_batch(self._h, pref, pdeg, offs.ctypes.data, lens.ctypes.data, n, int(fs),
                                        metric_mask(metrics), flags, pd, drows, int(seed) & (2 ** 64 - 1), phl,
                                        scores.ctypes.data, raw.ctypes.data, status.ctypes.data,
                                        None if stream is None else int(stream))
        self._check(rc, "nele_score_batch")
        return BatchResult(scores, raw, status)

    def prefetch(self, ref, deg, offs, lens, haspi_v1=False):
        """Start uploading the host waveforms of an upcoming ``score_packed`` call (same ``ref`` /
        ``deg`` pointers and pair count) while the current one computes; see ``nele_prefetch``."""
        offs = np.ascontiguousarray(offs, dtype=np.int64)
        lens = np.ascontiguousarray(lens, dtype=np.int32)
        if isinstance(ref, np.ndarray):
            pref, pdeg = ref.ctypes.data, deg.ctypes.data
        else:
            pref, pdeg = int(ref), int(deg)
        rc = self._lib.nele_prefetch(self._h, pref, pdeg, offs.ctypes.data, lens.ctypes.data, int(lens.shape[0]),
                                     FLAG_HASPI_V1 if haspi_v1 else 0)
        self._check(rc, "nele_prefetch")

    def score_packed_pcm16(self, clean, enhanced, noise, offs, lens, fs=16000, metrics=METRIC_ALL, mapped=True, seed=0,
                           no_dither=False, stream=None, out=None):
        """``nele_score_batch_pcm16``: host int16 arrays (or raw host pointers) of the clean, enhanced and noise
        signals in the packed layout of :meth:`score_packed`; ``deg = enhanced + noise`` is formed on the device."""
        offs = np.ascontiguousarray(offs, dtype=np.int64)
        lens = np.ascontiguousarray(lens, dtype=np.int32)
        n = int(lens.shape[0])
        ptrs = []
        for a in (clean, enhanced, noise):
            if isinstance(a, np.ndarray):
                if a.dtype != np.int16 or not a.flags.c_contiguous:
                    raise ValueError("PCM-16 inputs must be contiguous int16 arrays")
                ptrs.append(a.ctypes.data)
            else:
                ptrs.append(int(a))
        flags = (FLAG_MAPPED if mapped else 0) | (FLAG_NO_DITHER if no_dither else 0)
        if out is None:
            scores, raw, status = np.empty((n, 3)), np.empty((n, 10)), np.empty(n, dtype=np.int32)
        else:
            scores, raw, status = out
        rc = self._lib.nele_score_batch_pcm16(self._h, ptrs[0], ptrs[1], ptrs[2], offs.ctypes.data, lens.ctypes.data, n,
                                              int(fs), metric_mask(metrics), flags, None, 0, int(seed) & (2 ** 64 - 1), None,
                                              scores.ctypes.data, raw.ctypes.data, status.ctypes.data,
                                              None if stream is None else int(stream))
        self._check(rc, "nele_score_batch_pcm16")
        return BatchResult(scores, raw, status)

    def prefetch_pcm16(self, clean, enhanced, noise, offs, lens):
        """Upload-ahead for :meth:`score_packed_pcm16` (``nele_prefetch_pcm16``)."""
        offs = np.ascontiguousarray(offs, dtype=np.int64)
        lens = np.ascontiguousarray(lens, dtype=np.int32)
        p = [a.ctypes.data if isinstance(a, np.ndarray) else int(a) for a in (clean, enhanced, noise)]
        self._check(self._lib.nele_prefetch_pcm16(self._h, p[0], p[1], p[2], offs.ctypes.data, lens.ctypes.data,
                                                  int(lens.shape[0]), 0), "nele_prefetch_pcm16")

    def prefetch_cancel(self):
        """Drop pending prefetches (``nele_prefetch_cancel``)."""
        self._check(self._lib.nele_prefetch_cancel(self._h), "nele_prefetch_cancel")

    # ----------------------------------------------------------------- high level
    def score_batch(self, refs, degs, fs=16000, metrics=("siib", "haspi", "estoi"), mapped=True, **kw):
        """``refs``/``degs``: sequences of 1-D float arrays (ragged allowed).  Each
        pair is trimmed to its common length like intel.py:58-60 does."""
        refs = [np.asarray(r, dtype=np.float32) for r in refs]
        degs = [np.asarray(d, dtype=np.float32) for d in degs]
        if len(refs) != len(degs):
            raise ValueError("refs and degs differ in count")
        if not refs:
            return BatchResult(np.zeros((0, 3)), np.zeros((0, 10)), np.zeros(0, np.int32))
        for i, (r, d) in enumerate(zip(refs, degs)):
            m = min(len(r), len(d))
            if m == 0:
                raise ValueError("pair %d is empty" % i)
            refs[i], degs[i] = r[:m], d[:m]
        fr, offs, lens = pack(refs)
        fd, _, _ = pack(degs)
        return self.score_packed(fr, fd, offs, lens, fs=fs, metrics=metrics, mapped=mapped, **kw)

    # ------------------------------------------------------------- feature front-end
    def features_packed(self, wav, offs, lens, power=1.0 / 6.0, noise=False, normalization=True, want_mag=True,
                        want_phase=True, want_psd=False, device_io=False, out=None, stream=None):
        """``nele_features`` (audio_util.py:422-457 for a batch).  Host mode: ``wav`` flat float32 numpy
        array -> dict of numpy arrays ``band`` [sum T, 64], ``mag`` / ``phase`` / ``psd`` flat
        [257 * sum T] (waveform i's [257, T_i] matrix at 257 * foff[i]) and ``foff`` / ``frames``.
        Device mode (``device_io``): ``wav`` and ``out = (band, mag, phase, psd)`` are raw device
        pointers (ints or None)."""
        offs = np.ascontiguousarray(offs, dtype=np.int64)
        lens = np.ascontiguousarray(lens, dtype=np.int32)
        n = int(lens.shape[0])
        frames = 1 + lens.astype(np.int64) // 256
        foff = np.concatenate(([0], np.cumsum(frames)[:-1])).astype(np.int64) if n else np.zeros(0, np.int64)
        tot = int(frames.sum())
        flags = (FEAT_NOISE if noise else 0) | (FEAT_DEVICE_IO if device_io else 0) | (0 if normalization else FEAT_NO_POWER)
        res = {"foff": foff, "frames": frames}
        if device_io:
            pw = int(wav)
            ptrs = [None if p is None else int(p) for p in out]
        else:
            wav = np.ascontiguousarray(wav, dtype=np.float32)
            pw = wav.ctypes.data
            res["band"] = np.empty((tot, 64), dtype=np.float32)
            if want_mag:
                res["mag"] = np.empty(257 * tot, dtype=np.float32)
            if want_phase:
                res["phase"] = np.empty(257 * tot, dtype=np.float32)
            if want_psd and noise:
                res["psd"] = np.empty(257 * tot, dtype=np.float32)
            ptrs = [res[k].ctypes.data if k in res else None for k in ("band", "mag", "phase", "psd")]
        rc = self._lib.nele_features(self._h, pw, offs.ctypes.data, lens.ctypes.data, n, flags, float(power), ptrs[0],
                                     ptrs[1], ptrs[2], ptrs[3], None if stream is None else int(stream))
        self._check(rc, "nele_features")
        return res

    def features(self, signals, power=1.0 / 6.0, noise=False, normalization=True, want_psd=False):
        """Per-signal ``(bandE [T, 64], mag [257, T], phase [257, T])`` tuples (plus the noise PSD
        [257, T] when ``want_psd``) for a list of 1-D signals -- the return values of
        ``Sp_and_phase_Speech`` / ``Sp_and_phase_Noise`` (audio_util.py:422-457)."""
        signals = [np.asarray(x, dtype=np.float32) for x in signals]
        if not signals:
            return []
        flat, offs, lens = pack(signals)
        r = self.features_packed(flat, offs, lens, power=power, noise=noise, normalization=normalization,
                                 want_psd=want_psd)
        out = []
        for i in range(len(signals)):
            T, fo = int(r["frames"][i]), int(r["foff"][i])
            item = [r["band"][fo:fo + T], r["mag"][257 * fo:257 * (fo + T)].reshape(257, T),
                    r["phase"][257 * fo:257 * (fo + T)].reshape(257, T)]
            if want_psd and noise:
                item.append(r["psd"][257 * fo:257 * (fo + T)].reshape(257, T))
            out.append(tuple(item))
        return out

    def resyn(self, clean, noise, offs, lens, alpha2, arow=None, enh=None, deg=None, pcm16=True, enh_rounded=False,
              stream=None):
        """``nele_resyn``: device pointers (ints) ``clean`` / ``noise`` / ``alpha2`` / ``enh`` / ``deg``; host ``offs``
        int64[n], ``lens`` int32[n], ``arow`` int64[n] (first row of each utterance in ``alpha2``; None = packed).
        Returns the valid output lengths ``256 * (lens // 256)`` (int32[n])."""
        offs = np.ascontiguousarray(offs, dtype=np.int64)
        lens = np.ascontiguousarray(lens, dtype=np.int32)
        out_lens = np.empty_like(lens)
        if arow is not None:
            arow = np.ascontiguousarray(arow, dtype=np.int64)
        rc = self._lib.nele_resyn(self._h, int(clean), None if noise is None else int(noise), offs.ctypes.data,
                                  lens.ctypes.data, int(lens.shape[0]), int(alpha2),
                                  None if arow is None else arow.ctypes.data, (1 if pcm16 else 0) | (2 if enh_rounded else 0),
                                  None if enh is None else int(enh), None if deg is None else int(deg),
                                  out_lens.ctypes.data, None if stream is None else int(stream))
        self._check(rc, "nele_resyn")
        return out_lens

    def last_timing(self):
        """(kernel milliseconds, kernel launches) of the last call."""
        ms, nl = C.c_double(), C.c_int64()
        self._check(self._lib.nele_last_timing(self._h, C.byref(ms), C.byref(nl)), "nele_last_timing")
        return ms.value, nl.value

    def set_profiling(self, on=True):
        """Bracket every kernel launch of the following calls with CUDA events."""
        self._check(self._lib.nele_set_profiling(self._h, 1 if on else 0), "nele_set_profiling")

    def kernel_times(self):
        """{kernel name: (summed ms, launches)} of the last call made with profiling on."""
        out, i = {}, 0
        name, ms, nl = C.c_char_p(), C.c_double(), C.c_int64()
        while self._lib.nele_kernel_time(self._h, i, C.byref(name), C.byref(ms), C.byref(nl)) == 0:
            out[name.value.decode()] = (ms.value, nl.value)
            i += 1
        return out

    _STAGE_DTYPES = {"haspi.mid": np.float32, "haspi.x24": np.float32, "haspi.bw": np.float64,
                     "haspi.shift": np.int32, "haspi.envlp": np.float32, "haspi.nsel": np.int32,
                     "haspi.cep": np.float32, "haspi.cepmean": np.float64, "estoi.x10": np.float32,
                     "haspi1.segsum": np.float32, "haspi1.cov": np.float32, "haspi1.msx": np.float32,
                     "estoi.tob": np.float32, "estoi.info": np.int32, "estoi.kept": np.int32,
                     "siib.tile": np.int32, "siib.logspec": np.float32, "siib.lambda": np.float32,
                     "siib.rho": np.float32, "siib.rank": np.int32, "siib.sxx": np.float64,
                     "siib.sxy": np.float32, "siib.syy": np.float32}

    def stage(self, name, pair=0):
        """Flat array of stage ``name`` for ``pair`` of the last keep_stages call."""
        nb = C.c_size_t()
        self._check(self._lib.nele_get_stage(self._h, name.encode(), int(pair), None, 0, C.byref(nb)), "nele_get_stage")
        buf = np.empty(nb.value, dtype=np.uint8)
        self._check(self._lib.nele_get_stage(self._h, name.encode(), int(pair), buf.ctypes.data, nb.value, C.byref(nb)),
                    "nele_get_stage")
        return buf.view(self._STAGE_DTYPES[name])


_default = {}
_default_lock = threading.Lock()


def default_engine(device=None):
    """Process-wide engine per device (device defaults to $NELE_DEVICE, $LOCAL_RANK or 0)."""
    if device is None:
        device = int(os.environ.get("NELE_DEVICE", os.environ.get("LOCAL_RANK", "0")))
    with _default_lock:
        if device not in _default:
            _default[device] = Engine(device)
        return _default[device]
