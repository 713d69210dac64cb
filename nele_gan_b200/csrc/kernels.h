// Launch interface between the engine (engine.cu) and the kernel files.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "ear_core.cuh"

namespace nele {

// Per-pair geometry of one sub-batch ("chunk"), device arrays of length n.
struct PairGeom {
  const int64_t* off16;  // start of the pair in the 16 kHz (input-rate) buffers
  const int32_t* len16;  // samples at the input rate
  const int64_t* off24;  // start in the 24 kHz buffers (per signal plane)
  const int32_t* n24;    // ceil(len * 24000 / fs)
  const int64_t* offsub; // start (in rows of 32) in the decimated-envelope buffers
  const int32_t* nsub;   // ceil(n24 / 9)
  const int64_t* offblk; // start (in rows of 32) in the HASPI v1 buffers indexed by 192-sample blocks
};

struct HaspiBuffers {
  const float* ref;      // device, input rate
  const float* deg;
  float* x24;            // [2][tot24] resampled signal before the level match (scratch of haspi_prep)
  float* mid;            // [2][tot24] middle-ear output
  int64_t tot24;
  double* bw;            // [n][2][32]
  double* cave;          // [n][2][32] RMS of the control envelope (eb_EarModel xcave / ycave), or null
  int32_t* shift;        // [n][32]
  float* envlp;          // [2][totsub][32]
  int64_t totsub;
  int32_t* rowsel;       // [totsub] per pair: ordered list of the envelope rows above the loudness threshold
  int32_t* nsel;         // [n]
  float* cep;            // [2][5][totsub]  (pair p, signal q, coef j at ((q*5+j)*totsub + offsub[p]))
  double* cepmean;       // [n][2][5]
  double* modsum;        // [n][5][10][5]  {Sx, Sy, Sxx, Syy, Sxy}
  const BandConst* bands;  // [32] device
  const double* rs_taps;   // [up][128] device
  int rs_up, rs_down;      // 24 kHz resampler ratio (1/1 = identity)
  const float* dither;     // [2][dither_rows][32] device or null
  int64_t dither_rows;
  uint64_t seed;
  int no_dither;
  int64_t pair_base;       // global index of the chunk's first pair (dither keying)
};

// Optional per-kernel CUDA-event timing on the launching stream (bench.py's
// roofline figures).  Entries are filled in launch order; the engine sums them
// by name after the stream has been synchronised.
struct KernelTimer {
  static constexpr int kMax = 512;
  bool enabled = false;
  int count = 0;
  const char* names[kMax];
  cudaEvent_t ev0[kMax], ev1[kMax];
};
inline void kt_begin(KernelTimer* kt, const char* name, cudaStream_t s) {
  if (!kt || !kt->enabled || kt->count >= KernelTimer::kMax) return;
  kt->names[kt->count] = name;
  cudaEventRecord(kt->ev0[kt->count], s);
}
inline void kt_end(KernelTimer* kt, cudaStream_t s) {
  if (!kt || !kt->enabled || kt->count >= KernelTimer::kMax) return;
  cudaEventRecord(kt->ev1[kt->count], s);
  ++kt->count;
}

void haspi_upload_constants(cudaStream_t s);
// returns number of kernel launches issued
int haspi_run(const PairGeom& g, const HaspiBuffers& b, int n, int max_nsub, bool f64, KernelTimer* kt, cudaStream_t s);
// scores: raw Intel [n] and aveCM [n][10]; status byte per pair
int haspi_finish(const HaspiBuffers& b, int n, double* intel, double* raw10, int32_t* status, KernelTimer* kt,
                 cudaStream_t s);

// --------------------------------------------------------------- HASPI v1
// Back-end of haspi() (pyhaspi2.py:109-157) on the same ear model: the envelopes are smoothed
// to 16 ms segments inside the ear kernel, the basilar-membrane motion goes through HBM once.
struct HaspiV1Buffers {
  float* bm;         // [2][tot24][32] delay-compensated BM motion (+ threshold noise) of x and y
  float* segsum;     // [2 signals][2 (rising, falling half window)][totblk][32] windowed block sums
  int64_t totblk;
  float* cov;        // [totblk][32] segment cross-covariance (eb_BMcovary sigcov)
  float* msx;        // [totblk][32] 2 * MSx (eb_BMcovary sigMSx)
  double* xsum;      // [totblk] segment loudness of eb_3LevelCovary (-1e300 = below threshold)
  double* cepcorr;   // [n]
  double* cov3;      // [n][3]
  double* ave;       // [n][2][32] RMS of the signal-path envelope before compression (xave / yave), HASQI only
  double* sync5;     // [n] eb_AveCovary2 syncov[4], HASQI only
  int hasqi;         // compute the HASQI extras
  int32_t* status;   // [n] 0 ok, 1 below threshold
};
void haspi_v1_upload_tables(const float* cepm, cudaStream_t s);
// prep / control / shift kernels of haspi_run, then the v1 ear kernel and back-end
int haspi_v1_run(const PairGeom& g, const HaspiBuffers& b, const HaspiV1Buffers& v, int n, int max_n24, bool f64,
                 KernelTimer* kt, cudaStream_t s);
int haspi_v1_finish(const PairGeom& g, const HaspiBuffers& b, const HaspiV1Buffers& v, int n, double* intel, double* raw10,
                    int32_t* status, KernelTimer* kt, cudaStream_t s);
// front half of haspi_run (prep, control, shift), shared by both versions
int haspi_run_front(const PairGeom& g, const HaspiBuffers& b, int n, bool f64, KernelTimer* kt, cudaStream_t s);

// ------------------------------------------------------------------ ESTOI
struct EstoiGeom {
  const int64_t* off16;
  const int32_t* len16;
  const int64_t* off10;  // start in the 10 kHz buffers (per signal plane)
  const int32_t* n10;    // ceil(len * 10000 / fs)
  const int64_t* offfr;  // start in the frame-indexed buffers
  const int32_t* nfa;    // analysis frames: len(range(0, n10 - 256, 128))
};
struct EstoiBuffers {
  const float* ref;
  const float* deg;
  float* x10;        // [2][tot10]
  int64_t tot10;
  double* energy;    // [totfr] frame energies of x (dB)
  int32_t* kept;     // [totfr] indices of the frames that survive silent-frame removal
  int32_t* nkept;    // [n]
  float* tob;        // [2][totfr][15] one-third-octave magnitudes
  int64_t totfr;
  const double* taps;  // [up][2K+1]
  int up, down, K;
  double* score;     // [n]
  int32_t* status;   // [n]
};
void estoi_upload_tables(const float* win, const int* lo, const int* hi, const float* tw, cudaStream_t s);
void estoi_upload_polytaps(const double* taps, int up, int K, cudaStream_t s);
int estoi_run(const EstoiGeom& g, const EstoiBuffers& b, int n, int max_n10, int max_nfa, bool classic, KernelTimer* kt,
              cudaStream_t s);

// ------------------------------------------------------------------- SIIB
struct SiibGeom {
  const int64_t* off16;
  const int32_t* len16;
  const int64_t* offW;   // start in wrapdb (frames of the untiled signal)
  const int64_t* offF;   // start in the buffers indexed by frames of the tiled signal
  const int64_t* F;      // frames of the tiled signal: ceil((M L - 400) / 200)
};
struct SiibBuffers {
  const float* ref;
  const float* deg;
  double* wrapdb;        // [totW]
  int32_t* M;            // [n] tiling factor of intel.py:71-75 (0: no active frame)
  int32_t* wrap_active;  // [n]
  double* mean;          // [n][2]
  double* xdb;           // [totF]
  int32_t* act;          // [totF] indices of the active frames
  int32_t* aidx;         // [totF] frame (first period only) -> index in act
  int32_t* src;          // [totF] active frame -> row of lograw holding its spectrum (first occurrence)
  int32_t* Fa;           // [n]
  int32_t* Pact;         // [n] active frames per period of the tiled signal (0 = no repetition)
  int32_t* perflag;      // [n][2] masked features of x / y verified periodic from the second period on
  int no_proj;           // 1: never take the projection route for periodic pairs (NELE_SIIB_QUADFORM=1, A/B runs)
  int xx32;              // 1: pairs whose x features do not repeat take the xx lag products in FP32 too (set by siib_run when the
                         // tridiagonalisation path follows, which rounds Sxx to FP32 anyway; NELE_COV_XX64=1 keeps FP64, A/B runs)
  float* lograw;         // [2][totF][32] log band energies of the distinct active frames
  float* logspec;        // [2][totF][32] after forward masking and mean removal
  int64_t totF;
  // per sub-chunk (indexed by pair - pair_lo)
  int pair_lo;
  double* base;          // [59][32][32] lag products
  double* Sxx;           // [420][420]
  float* Sxy;            // [420][420]
  float* Syy;            // [420][420]
  double* Lc;            // [420][420] Cholesky factor, Lc[m][i] = L[i][m]
  float* G;              // [420][448] FP32 copy of the factor; Jacobi works in place
  int32_t* perm;         // [420]
  // per pair of the chunk
  int32_t* rank;         // [n]
  int32_t* sweeps;       // [n]
  int32_t* sweep_rot;    // [n][16] rotations applied in each sweep (diagnostic)
  float* lambda;         // [n][420]
  float* rho;            // [n][420]
  double* score;         // [n]
  int32_t* status;       // [n]
  double* info_part;     // [n][8] per-tile information sums of siib_launch_quadform
};
// k-NN (Kraskov) estimator of pysiib.SIIB(..., gauss=False): siib_knn.cu
struct SiibKnnBuffers {
  float* xk;             // [sub][2][420][ld] KLT-domain series of x and y (per sub-chunk)
  int64_t ld;            // row stride (frames, padded)
  double* info;          // [n][420] per-component information, bits
  const double* digamma; // [ndigamma] digamma(m), m = 0 unused
  int ndigamma;
};
void siib_knn_setup();
int siib_run_knn(const SiibGeom& g, const SiibBuffers& b, const SiibKnnBuffers& kb, int n, int64_t max_F, KernelTimer* kt,
                 cudaStream_t s);
// tridiagonalisation-based KLT for full-rank pairs: siib_eig.cu
struct SiibEigBuffers {
  double* d;       // [sub][448] diagonal of T
  double* e;       // [sub][448] sub-diagonal of T
  double* tau;     // [sub][448] Householder scalars
  double* lam;     // [sub][448] eigenvalues, ascending
  double* znorm;   // [sub][448] squared norms of the unnormalised eigenvectors of T
  float* scratch;  // [sub][420][448] per-thread D- sequence (aliases G: dead before G is written)
  float* zt;       // [sub][420][448] eigenvectors of T, entry-major (low-rank path: scratch + V)
  double* gram;    // [sub][112][112] zero-padded Gram matrix L^T L of the low-rank pairs
  float* refl;     // [sub][420][448] Householder vectors in FP32 (row k = reflector k)
};
int siib_run_small_eig(const SiibGeom& g, const SiibBuffers& b, const SiibEigBuffers& eb, int n, int rank_hi, KernelTimer* kt,
                       cudaStream_t s);
int siib_run_eig(const SiibGeom& g, const SiibBuffers& b, const SiibEigBuffers& eb, int n, int rank_lo, KernelTimer* kt,
                 cudaStream_t s);
// FP32 lower-triangle tridiagonalisation (siib_klt.cu): same outputs as siib_tridiag_kernel
int siib_launch_tridiag32(const SiibBuffers& b, const SiibEigBuffers& eb, int n, int rank_lo, cudaStream_t s);
// back-transformation with the reflectors applied four at a time (siib_klt.cu): same inputs / output as siib_backtf_kernel
int siib_launch_backtf4(const SiibBuffers& b, const SiibEigBuffers& eb, int n, int rank_lo, cudaStream_t s);
int siib_launch_backtf5(const SiibBuffers& b, const SiibEigBuffers& eb, int n, int rank_lo, cudaStream_t s);  // 14 x 4 register tile per lane
int siib_launch_backtf6(const SiibBuffers& b, const SiibEigBuffers& eb, int n, int rank_lo, cudaStream_t s);  // warp = 4 vectors over all rows
// register-tiled quadratic forms + score (siib_klt.cu); info_part: [n_chunk][8] doubles of scratch
int siib_launch_quadform(const SiibBuffers& b, double* info_part, int n, KernelTimer* kt, cudaStream_t s);
void siib_upload_tables(const float* win, const float* decay, const float* g2t, const float* tw, cudaStream_t s);
int siib_run_wrapvad(const SiibGeom& g, const SiibBuffers& b, int n, bool no_tile, KernelTimer* kt, cudaStream_t s);
// kb != nullptr: k-NN estimator instead of the Gaussian quadratic forms
// max_unique: upper bound over the pairs of the number of distinct frames, min(F, L / gcd(L, 200))
// eb != nullptr: pairs of rank > 112 take the tridiagonalisation path instead of the cluster Jacobi
int siib_run(const SiibGeom& g, const SiibBuffers& b, const SiibKnnBuffers* kb, const SiibEigBuffers* eb, int n, int64_t max_F,
             int64_t max_unique, KernelTimer* kt, cudaStream_t s);

// --------------------------------------------------------------- feature front-end (features.cu)
// STFT magnitude / phase / 64 band energies, and for noise inputs the IMCRA noise PSD
// (audio_util.py:422-457).  foff: frame offsets [n]; tiles: (waveform, first frame) per CTA of
// feat_stft.  All pointers are device pointers; mag must be non-null when noise is set.
int features_run(const float* wav, const int64_t* offs, const int32_t* lens, const int64_t* foff, const int2* tiles,
                 int n, int ntiles, bool noise, float power, bool normalize, float* band, float* mag, float* phase,
                 float* psd, KernelTimer* kt, cudaStream_t s);

// Resynthesis of a sampling round (audio_util.py:60-115, train_nele.py:303-314): tiles = (utterance, first frame) per CTA,
// 7 output hops each; alpha2 [sum T][64]; enh / deg written at the clean signal's offsets (either may be null)
int resyn_run(const float* clean, const float* noise, const int64_t* offs, const int32_t* lens, const int64_t* foff,
              const int2* tiles, int ntiles, const float* alpha2, int pcm16 /* NELE_RESYN_* bits */, float* enh, float* deg,
              KernelTimer* kt, cudaStream_t s);

}  // namespace nele
