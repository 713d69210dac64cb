/*
 * nele_score.h -- C ABI of the B200-native batched intelligibility-scoring
 * engine (libnele_score.so).
 *
 * This is the drop-in boundary for NELE-GAN's metric-labelling hot path.  The
 * reference has no FFI: its boundary is a set of Python callables.  Each entry
 * point below names the reference interface(s) it replaces (file:line under
 * the reference tree); INTEGRATION.md shows the ctypes binding a maintainer
 * adds on the reference side.
 *
 *   reference callable                                   replaced by
 *   ---------------------------------------------------  ------------------
 *   pyHASPI/pyhaspi2.py:76   haspi_v2(x, fx, y, fy, HL)   nele_score_batch, NELE_METRIC_HASPI
 *   pysiib.SIIB(x, y, fs, gauss=True)   (intel.py:77,100) nele_score_batch, NELE_METRIC_SIIB
 *   pystoi.stoi(x, y, fs, extended=True)(intel.py:126,133) nele_score_batch, NELE_METRIC_ESTOI
 *   intel.py:57-140  *_Wrapper[_raw]_harvard(x, y, fs)    nele_score_batch, NELE_FLAG_MAPPED on/off
 *   audio_util.py:145-203,281-321 read_batch_*            one nele_score_batch call per list
 *   audio_util.py:422-457  Sp_and_phase_Speech / _Noise   nele_features (the data loaders' feature front-end,
 *     (compute_band_E :30-50, STFT :52-57, NoisePSD        dataloader.py:30-84)
 *      :117-122 -> noise_est/imcra.py:487-577)
 *
 * Conventions: plain pointers and sizes, no ownership transfer.  The caller
 * owns inputs and outputs; the engine owns a grow-only device workspace.  One
 * engine per (process, device); calls on one engine are serialised by the
 * engine (a mutex per handle), so threads may share a handle but gain nothing
 * by it.  Every function returns 0 on success or a negative NELE_E_* code;
 * nele_last_error() gives the message.  Per-pair conditions that the
 * reference reports by raising (signal below threshold, too few frames) come
 * back in status[] and never abort the batch.
 */
#ifndef NELE_SCORE_H
#define NELE_SCORE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define NELE_ABI_VERSION 1

/* metric mask */
#define NELE_METRIC_HASPI 0x1u /* HASPI v2  (pyhaspi2.py:76-107); version 1 (:109-157) with NELE_FLAG_HASPI_V1 */
#define NELE_METRIC_SIIB  0x2u /* SIIB^Gauss (pysiib.SIIB(..., gauss=True)) with the wrapper's >=25 s tiling (intel.py:57-100) */
#define NELE_METRIC_ESTOI 0x4u /* ESTOI     (pystoi.stoi extended=True) */
#define NELE_METRIC_ALL   0x7u

/* flags */
#define NELE_FLAG_MAPPED        0x01u /* logistic mapping of intel.py:102-140 (norm=True); else raw scores */
#define NELE_FLAG_DEVICE_INPUT  0x02u /* ref/deg are device pointers (offs/lens stay on the host) */
#define NELE_FLAG_NO_DITHER     0x04u /* HASPI: zero cepstral dither (parity / deterministic mode) */
#define NELE_FLAG_SIIB_NO_TILE  0x08u /* SIIB: plain pysiib.SIIB semantics, no wrapper tiling */
#define NELE_FLAG_KEEP_STAGES   0x10u /* keep per-stage tensors of this call for nele_get_stage() */
#define NELE_FLAG_HASPI_V1      0x20u /* NELE_METRIC_HASPI computes HASPI version 1, haspi() of pyhaspi2.py:109-157:
                                         scores[.][1] = Intel (alpha = -1, never mapped), haspi_raw[.][0..3] =
                                         {CepCorr, cov3 low, mid, high}, [4..9] = NaN.  NELE_FLAG_NO_DITHER also
                                         zeroes the basilar-membrane threshold noise (pyhaspi2.py:1091-1095). */

#define NELE_FLAG_STOI_CLASSIC  0x40u /* NELE_METRIC_ESTOI computes classic STOI, pystoi.stoi(..., extended=False) */
#define NELE_FLAG_SIIB_KNN      0x80u /* NELE_METRIC_SIIB uses pysiib's k-nearest-neighbour (Kraskov) estimator,
                                         SIIB(x, y, fs, gauss=False), instead of SIIB^Gauss */
#define NELE_FLAG_HASQI_V2      0x100u /* NELE_METRIC_HASPI computes HASQI version 2, hasqi_v2() of pyhaspi2.py:32-74, on the
                                          version-1 pipeline: scores[.][1] = Combined, haspi_raw[.][0..5] = {CepCorr,
                                          BMsync5, Dloud, Dslope, Nonlin, Linear}.  Never mapped. */

/* error codes (function return values) */
#define NELE_OK              0
#define NELE_E_ARG          -1 /* bad argument (null pointer, n < 0, unsupported fs ...) */
#define NELE_E_CUDA         -2 /* CUDA runtime error; see nele_last_error */
#define NELE_E_NOMEM        -3
#define NELE_E_NODEVICE     -4 /* no CUDA device: the engine has no CPU path */

/* per-pair status: one byte per metric, (status >> 8*k) & 0xff,
 * k = 0 HASPI, 1 SIIB, 2 ESTOI */
#define NELE_ST_OK           0
#define NELE_ST_BELOW_THR    1 /* HASPI "Signal below threshold" (pyhaspi2.py:357-358): score = NaN */
#define NELE_ST_TOO_SHORT    2 /* ESTOI < 30 frames: score = 1e-5 (pystoi sentinel);
                                  SIIB < 20 s of active speech after tiling: score = NaN */
#define NELE_ST_BAD_RATE     3 /* sampling rate not supported for this metric: score = NaN */
#define NELE_ST_UNSUPPORTED  4 /* SIIB k-NN estimator: more than 16384 KLT frames (score = NaN) */
#define NELE_ST_SKIPPED      0xff /* metric not requested */
/* informational bits above the three status bytes (a pair that carries one is still NELE_ST_OK) */
#define NELE_INFO_SIIB_NULLSPACE 0x01000000 /* SIIB: the tiled signal repeats its frames exactly (utterance length a
                                               multiple of the 200-sample hop, intel.py:71-75), cov(X) is rank deficient
                                               and its null space was given zero information.  pysiib's float64 value is
                                               1-13 % higher on such inputs: the sample correlations of eigh's
                                               rounding-noise eigenvectors. */

typedef struct nele_engine nele_engine;

/* Create an engine on CUDA device `device`.  Fails with NELE_E_NODEVICE when
 * there is no usable GPU -- there is deliberately no CPU fallback. */
int nele_create(int device, nele_engine** out);
void nele_destroy(nele_engine* e);

/* Message of the last failure on this engine (or of the last failed
 * nele_create when e == NULL).  Valid until the next call. */
const char* nele_last_error(const nele_engine* e);
int nele_abi_version(void);

/*
 * Score n (clean, degraded) pairs.
 *
 *   ref, deg   concatenated float32 waveforms; pair i occupies
 *              [offs[i], offs[i] + lens[i]) in both.  x = clean reference,
 *              y = degraded = enhanced + noise, as audio_util.py:139 forms it.
 *              Host pointers unless NELE_FLAG_DEVICE_INPUT.
 *   offs,lens  host arrays [n]; lens[i] is the common (already trimmed)
 *              length -- the reference trims to min(len) in intel.py:58-60.
 *   fs         sampling rate of every pair.  16000 for all three metrics
 *              (audio_util.py:131 asserts it); HASPI alone accepts any
 *              fs <= 24000 (pyhaspi2.py:810-821).
 *   dither     NULL, or a host float32 matrix [2][dither_rows][32] of
 *              unit-variance normals shared by all pairs: row r of matrix 0
 *              (1) is added, scaled by 0.1 dB, to the r-th above-threshold
 *              envelope frame of x (y) -- the injected form of
 *              pyhaspi2.py:362-365 used for parity tests.  With NULL and no
 *              NELE_FLAG_NO_DITHER the engine draws its own normals from a
 *              counter-based generator keyed by (seed, pair, signal, row).
 *   hl         NULL (normal hearing, the only configuration the reference
 *              vouches for, pyHASPI/README.txt:14) or 6 audiogram losses in
 *              dB at 250..6000 Hz applied to y as pyhaspi2.py:1160 does.
 *   scores     host double [n][3], column order {SIIB, HASPI, ESTOI} -- the
 *              order train_nele.py:320-322 computes and
 *              audio_util.py:367-389 serialises them.
 *   haspi_raw  NULL or host double [n][10]: the ten modulation-band
 *              correlations haspi_v2 returns as `raw`.
 *   status     NULL or host int32 [n] (see NELE_ST_*).
 *   stream     cudaStream_t to run on, or NULL for the engine's own stream.
 *              The call returns after the results are in the output arrays.
 */
int nele_score_batch(nele_engine* e, const float* ref, const float* deg, const int64_t* offs,
                     const int32_t* lens, int n, int fs, uint32_t metrics, uint32_t flags,
                     const float* dither, int64_t dither_rows, uint64_t seed, const double* hl,
                     double* scores, double* haspi_raw, int32_t* status, void* stream);

/*
 * The same call for host inputs held as 16-bit PCM, which is what the reference's files are (the corpus, and the
 * generator's outputs written by sf.write(..., 'PCM_16'), train_nele.py:313): clean, enhanced and noise int16 arrays
 * with the layout of ref / deg above.  The engine forms ref = clean / 32768 (librosa.load) and
 * deg = enhanced / 32768 + noise / 32768 (audio_util.py:139) on the device -- both exact in float32, so the scores are
 * bit-identical to nele_score_batch on the converted arrays -- and 6 instead of 8 bytes per sample cross PCIe.
 * nele_prefetch_pcm16 is the matching upload-ahead call.
 */
int nele_score_batch_pcm16(nele_engine* e, const int16_t* clean, const int16_t* enhanced, const int16_t* noise,
                           const int64_t* offs, const int32_t* lens, int n, int fs, uint32_t metrics, uint32_t flags,
                           const float* dither, int64_t dither_rows, uint64_t seed, const double* hl,
                           double* scores, double* haspi_raw, int32_t* status, void* stream);
int nele_prefetch_pcm16(nele_engine* e, const int16_t* clean, const int16_t* enhanced, const int16_t* noise,
                        const int64_t* offs, const int32_t* lens, int n, uint32_t flags);

/*
 * Optional pipelining across calls for host inputs: start the host -> device upload of an upcoming
 * nele_score_batch(e, ref, deg, offs, lens, n, ..., flags, ...) call now, on the engine's copy
 * stream, into the staging slot the current call does not use.  Returns at once; the matching
 * nele_score_batch call (same ref, deg, n) finds its waveforms already on the device, so the
 * upload of call k + 1 hides behind the kernels of call k.  The host buffers must stay unchanged
 * until that call returns.  The reference has no counterpart (its pool re-reads WAV files per
 * task, audio_util.py:128-139); a no-op for device inputs and for calls that need more than one
 * chunk (those pipeline their own uploads).
 */
int nele_prefetch(nele_engine* e, const float* ref, const float* deg, const int64_t* offs, const int32_t* lens, int n,
                  uint32_t flags);
/* Drop every pending prefetch (e.g. when the step it was issued for is skipped).  A prefetch is matched by the
 * buffer addresses and by a hash of offs[] / lens[], and expires by itself once two nele_score_batch calls have
 * passed without consuming it. */
int nele_prefetch_cancel(nele_engine* e);

/*
 * Parity-test access to the per-stage tensors of the last nele_score_batch
 * call made with NELE_FLAG_KEEP_STAGES.  Copies stage `name` of pair `pair`
 * into dst (capacity cap bytes) and reports its size; dst may be NULL to query
 * the size.  Stages (element type, shape):
 *   "haspi.mid"    f32 [2][n24]        middle-ear output of x and y (pyhaspi2.py:1185-1186), computed in FP64
 *   "haspi.bw"     f64 [2][32]         BWx, BWy                      (pyhaspi2.py:1204-1205)
 *   "haspi.shift"  i32 [32]            group-delay shifts            (pyhaspi2.py:1118-1122)
 *   "haspi.envlp"  f32 [2][nsub][32]   ebm_EnvFilt output            (pyhaspi2.py:412-413)
 *   "haspi.nsel"   i32 [1]             frames above threshold        (pyhaspi2.py:355-356)
 *   "haspi.cep"    f32 [2][5][nsel]    de-meaned cepstra 2..6        (pyhaspi2.py:366-374)
 *   "haspi1.segsum" f32 [4][nblk][32]  (NELE_FLAG_HASPI_V1) Hann-weighted sums of the envelopes over the
 *                                      192-sample blocks: {x rising half, x falling half, y rising, y falling};
 *                                      eb_EnvSmooth segment s = (rise[s] + fall[s+1]) / sum(window) (pyhaspi2.py:692-700)
 *   "haspi1.cov"   f32 [nseg][32]      eb_BMcovary sigcov, transposed   (pyhaspi2.py:643-657)
 *   "haspi1.msx"   f32 [nseg][32]      eb_BMcovary sigMSx, transposed
 *   "estoi.x10"    f32 [2][n10]        10 kHz signals (pystoi resample_oct)
 *   "estoi.info"   i32 [3]             {n10, analysis frames, frames kept by the 40 dB mask}
 *   "estoi.kept"   i32 [kept]          indices of the kept frames
 *   "estoi.tob"    f32 [2][kept-1][15] one-third-octave magnitudes of the silence-removed signals
 *   "siib.tile"    i32 [4]             {M, active frames (wrapper VAD), frames of the tiled signal, active frames}
 *   "siib.logspec" f32 [2][Fa][32]     masked, de-meaned log band energies (28 bands + 4 zero lanes)
 *   "siib.sxx"     f64 [420][420]      centred scatter matrix of the stacked clean features
 *   "siib.sxy" / "siib.syy"  f32 [420][420]  cross / degraded scatter matrices FOLDED onto the lower triangle, the form the
 *                                     quadratic forms read: F[c][a] = S[c][a] + S[a][c] (c > a), F[a][a] = S[a][a]; the
 *                                     upper triangle is not written (nor any of it for pairs on the projection route)
 *   "siib.rank"    i32 [2]             {numerical rank of Sxx, Jacobi sweeps}
 *   "siib.lambda"  f32 [420]           eigenvalues of Sxx (order of the Jacobi columns; 0 beyond the rank)
 *   "siib.rho"     f32 [420]           per-component correlation
 */
int nele_get_stage(nele_engine* e, const char* name, int pair, void* dst, size_t cap, size_t* nbytes);

/* Device-side time (ms, CUDA events on the launching stream) of the kernels of
 * the last nele_score_batch call, excluding host<->device copies; and the
 * number of kernel launches it made. */
int nele_last_timing(const nele_engine* e, double* kernel_ms, int64_t* launches);

/* Per-kernel device times of the last nele_score_batch call (bench.py's roofline
 * figures).  nele_set_profiling(e, 1) makes every following call bracket each
 * kernel launch with CUDA events on the launching stream; nele_kernel_time
 * enumerates the kernels by idx = 0, 1, ... (NELE_E_ARG past the end): name,
 * summed milliseconds and number of launches.  Off by default. */
int nele_set_profiling(nele_engine* e, int on);
int nele_kernel_time(const nele_engine* e, int idx, const char** name, double* ms, int64_t* launches);

/*
 * Feature front-end of the generator / discriminator data loaders (dataloader.py:30-84): for each of
 * n waveforms the centred 512 / 256 STFT with a periodic Hann window and reflect padding
 * (librosa.stft as audio_util.py:52-57 calls it), its magnitude and phase, and the 64 band energies
 * of compute_band_E (audio_util.py:30-50) raised to `power` (dataloader.py:14: 1/6).
 *
 *   Sp_and_phase_Speech(signal, power, Normalization)  (audio_util.py:422-437): flags = 0
 *   Sp_and_phase_Noise(signal, power, Normalization)   (audio_util.py:439-457): NELE_FEAT_NOISE -- the band
 *       energies are those of the IMCRA noise PSD, NoisePSD = imcra_est(nfft=512).estimate
 *       (audio_util.py:117-122, noise_est/imcra.py:487-577 and :362-484), one fresh estimator per waveform
 *
 *   wav, offs, lens   concatenated float32 waveforms, waveform i at [offs[i], offs[i] + lens[i]); lens[i] >= 257
 *                     (reflect padding needs more than n_fft / 2 samples, as in librosa).  offs / lens on the host.
 *   frames            T_i = 1 + lens[i] / 256; F_i = sum of T_j over j < i  (nele_feature_frames)
 *   band              float32 [sum T][64]: rows F_i .. F_i + T_i of waveform i -- bandE[T, 64] of the reference
 *   mag, phase, psd   NULL or float32 [257 * sum T]: waveform i's [257][T_i] matrix (the reference's layout) at
 *                     257 * F_i; psd (NELE_FEAT_NOISE only) is the noise PSD NoisePSD returns
 *   flags             NELE_FEAT_NOISE; NELE_FEAT_DEVICE_IO: wav and every output are device pointers;
 *                     NELE_FEAT_NO_POWER: Normalization=False, band energies not raised to `power`
 */
#define NELE_FEAT_NOISE     0x1u
#define NELE_FEAT_DEVICE_IO 0x2u
#define NELE_FEAT_NO_POWER  0x4u
int64_t nele_feature_frames(int32_t len);
int nele_features(nele_engine* e, const float* wav, const int64_t* offs, const int32_t* lens, int n, uint32_t flags,
                  double power, float* band, float* mag, float* phase, float* psd, void* stream);

/*
 * In-loop boundary of a GAN sampling round (train_nele.py:286-314): from the generator's band energy gains to the
 * degraded waveforms the metrics score, on the device -- SP_to_wav / Resyn / interp_band_gain / librosa.istft
 * (audio_util.py:60-115, 458-461), the PCM-16 rounding of sf.write(..., 'PCM_16') (train_nele.py:313) and
 * `enhanced + noise` (audio_util.py:196).  Every utterance is processed at its own length (reflect padding at its own
 * ends), as the reference does one file at a time.
 *
 *   clean, noise   device float32, utterance i at [offs[i], offs[i] + lens[i]) (offs / lens on the host, lens[i] > 256)
 *   alpha2, arow   device float32 [rows][64]: `mask * beta_2`, T_i = 1 + lens[i] / 256 rows per utterance starting at row
 *                  arow[i] (host int64 [n]; NULL: packed back to back, arow[i] = sum of T_j over j < i).  A generator
 *                  output padded to [n][Tmax][64] is passed with arow[i] = i * Tmax.
 *   enh, deg       device float32 outputs at the same offsets (either may be NULL): the resynthesised signal before the
 *                  PCM-16 rounding, and round16(enh) + noise
 *   out_lens       host int32 [n]: 256 * (lens[i] / 256) valid samples -- len(librosa.istft(...)), to which
 *                  audio_util.py:190-193 trims both signals; pass it as `lens` to nele_score_batch
 *                  (NELE_FLAG_DEVICE_INPUT, ref = clean, deg = deg, same offs)
 *   flags          NELE_RESYN_PCM16: round the enhanced signal to 16 bits before the noise is added;
 *                  NELE_RESYN_ENH_ROUNDED: `enh` receives the rounded signal (what the discriminator's data loader reads
 *                  back from the WAV file, dataloader.py:59) instead of the unrounded one
 */
#define NELE_RESYN_PCM16       0x1u
#define NELE_RESYN_ENH_ROUNDED 0x2u
int nele_resyn(nele_engine* e, const float* clean, const float* noise, const int64_t* offs, const int32_t* lens, int n,
               const float* alpha2, const int64_t* arow, uint32_t flags, float* enh, float* deg, int32_t* out_lens,
               void* stream);

#ifdef __cplusplus
}
#endif
#endif /* NELE_SCORE_H */
