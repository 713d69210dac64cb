// Feature front-end of NELE-GAN's generator / discriminator data loaders (SURVEY.md section 8f,
// rank 3): what dataloader.py:30-84 computes per utterance through audio_util.py:422-457 --
//
//   Sp_and_phase_Speech : librosa.stft(512 / 256, periodic Hann, centred, reflect padding) ->
//                         |F|, angle(F), compute_band_E(|F|) ** power        (audio_util.py:30-57, 422-437)
//   Sp_and_phase_Noise  : the same STFT, then the IMCRA noise PSD of noise_est/imcra.py
//                         (imcra_est.estimate, :487-577, driving imcra.update, :362-484) ->
//                         compute_band_E(sqrt(PSD)) ** power                (audio_util.py:117-122, 439-457)
//
// Two kernels.  feat_stft: one CTA per 8 consecutive frames of one utterance; the frames are
// paired into four 512-point complex FFTs (frame a + i frame b), radix-8 Stockham in FP64 through
// shared memory (the reference transforms in float64 and stores complex64, so the spectra are
// rounded to float32 from the same precision), split into the two real spectra, magnitude / phase
// written in the reference's [257][T] layout, band energies with the reference's exact
// accumulation order.  feat_imcra: one CTA per utterance, thread = frequency bin, the frame
// recursion in FP64 with the spectra staged 32 frames at a time through shared memory; the 3-tap
// frequency smoothing goes through shared memory, the minimum store is a register ring; the noise
// band energies are formed in the same pass, so the PSD only goes to HBM when the caller asks.
#include <math.h>

#include "kernels.h"

namespace nele {

namespace {

constexpr int kNfft = 512, kHop = 256, kBins = 257, kFeatBands = 64;
constexpr int kTileFrames = 8;

// audio_util.py:23
__constant__ int c_gmt[kFeatBands] = {0,  3,  4,  5,  6,  7,  8,  9,  10,  11,  12,  13,  14,  15,  16,  17,
                                      18, 19, 20, 21, 22, 23, 24, 25, 26,  28,  30,  32,  34,  36,  38,  41,
                                      43, 46, 49, 52, 55, 58, 62, 66, 70,  74,  79,  83,  88,  93,  99,  105,
                                      111, 117, 124, 131, 139, 147, 156, 165, 174, 184, 195, 206, 218, 230, 243, 257};

__device__ __forceinline__ double2 cmul(double2 a, double2 b) {
  return make_double2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}
__device__ __forceinline__ double2 cadd(double2 a, double2 b) { return make_double2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ double2 csub(double2 a, double2 b) { return make_double2(a.x - b.x, a.y - b.y); }
__device__ __forceinline__ double2 mul_mi(double2 a) { return make_double2(a.y, -a.x); }  // a * (-i)

// 4-point DFT, natural order in and out
__device__ __forceinline__ void fft4(double2& a0, double2& a1, double2& a2, double2& a3) {
  double2 s0 = cadd(a0, a2), s1 = csub(a0, a2), s2 = cadd(a1, a3), s3 = mul_mi(csub(a1, a3));
  a0 = cadd(s0, s2);
  a2 = csub(s0, s2);
  a1 = cadd(s1, s3);
  a3 = csub(s1, s3);
}

// 8-point DFT (forward, exp(-2 pi i r q / 8)), natural order in and out
__device__ __forceinline__ void fft8(double2* v) {
  const double c = 0.70710678118654752440;
  fft4(v[0], v[2], v[4], v[6]);
  fft4(v[1], v[3], v[5], v[7]);
  double2 o1 = make_double2(c * (v[3].x + v[3].y), c * (v[3].y - v[3].x));    // * W8
  double2 o2 = mul_mi(v[5]);                                                  // * W8^2
  double2 o3 = make_double2(c * (v[7].y - v[7].x), -c * (v[7].x + v[7].y));   // * W8^3
  double2 e0 = v[0], e1 = v[2], e2 = v[4], e3 = v[6], o0 = v[1];
  v[0] = cadd(e0, o0);
  v[4] = csub(e0, o0);
  v[1] = cadd(e1, o1);
  v[5] = csub(e1, o1);
  v[2] = cadd(e2, o2);
  v[6] = csub(e2, o2);
  v[3] = cadd(e3, o3);
  v[7] = csub(e3, o3);
}

// compute_band_E (audio_util.py:30-50) for one band from the squared magnitudes `sq` (float32, as
// numpy squares them): band b collects frac * sq over the bins of band b - 1, then (1 - frac) * sq
// over its own bins -- in that order, float32 products accumulated in float64.  d_wown[k] =
// float32(1 - frac), d_wnext[k] = float32(frac) of bin k (numpy multiplies the float32 square by the
// python float as a float32).
__device__ float d_wown[kBins], d_wnext[kBins];

__device__ __forceinline__ float band_energy(const float* sq, int b) {
  double acc = 0.0;
  if (b > 0)
    for (int k = c_gmt[b - 1]; k < c_gmt[b]; ++k) acc += (double)__fmul_rn(__ldg(&d_wnext[k]), sq[k]);
  if (b < kFeatBands - 1)
    for (int k = c_gmt[b]; k < c_gmt[b + 1]; ++k) acc += (double)__fmul_rn(__ldg(&d_wown[k]), sq[k]);
  return (float)acc;
}

// bandE ** power in float32 (audio_util.py:432): numpy calls powf too (CUDA's is within 4 ulp; the
// correctly rounded double pow costs ten times as much -- it was the largest item of feat_stft).
__device__ __forceinline__ float band_power(float v, float power, int normalize) {
  return normalize ? powf(v, power) : v;
}

// exp(-2 pi i m / 512), m = 0..511 (host-computed, uploaded once; 8 KB, L1-resident)
__device__ double2 d_tw512[kNfft];

struct StftMagPh {
  float mag[kTileFrames][kBins + 3];
  float ph[kTileFrames][kBins + 3];
};
union StftSmem {   // the spectra are consumed into registers before magnitude / phase overwrite them
  double2 buf[4][kNfft];
  StftMagPh o;
};

__global__ void __launch_bounds__(256) feat_stft_kernel(const float* __restrict__ wav, const int64_t* __restrict__ offs,
                                                         const int32_t* __restrict__ lens,
                                                         const int64_t* __restrict__ foff, const int2* __restrict__ tiles,
                                                         float power, int normalize, float* __restrict__ band,
                                                         float* __restrict__ mag, float* __restrict__ phase) {
  __shared__ __align__(16) StftSmem sm;
  __shared__ int nonzero[kTileFrames];   // an all-zero frame (digital silence) must come out exactly zero:
                                         // paired with a live frame it would pick up that frame's roundoff
  const int tid = threadIdx.x;
  if (tid < kTileFrames) nonzero[tid] = 0;
  __syncthreads();
  const int2 tile = tiles[blockIdx.x];
  const int u = tile.x, t0 = tile.y;
  const int L = lens[u];
  const int T = 1 + L / kHop;
  const float* x = wav + offs[u];

  // windowed frames: FFT q takes frame 2q as its real and frame 2q + 1 as its imaginary part
  for (int e = tid; e < kTileFrames * kNfft; e += 256) {
    int f = e >> 9, n = e & 511;
    int t = t0 + f;
    double v = 0.0;
    if (t < T) {
      int p = t * kHop + n - kNfft / 2;   // centred frame, reflect padding (np.pad(..., mode='reflect'))
      if (p < 0) p = -p;
      if (p >= L) p = 2 * (L - 1) - p;
      v = (0.5 - 0.5 * __ldg(&d_tw512[n].x)) * (double)x[p];
      if (v != 0.0) nonzero[f] = 1;
    }
    double* dst = reinterpret_cast<double*>(&sm.buf[f >> 1][n]);
    dst[f & 1] = v;
  }
  __syncthreads();
  {
    const int q = tid >> 6, j = tid & 63;
    double2 v[8];
#pragma unroll
    for (int stage = 0; stage < 3; ++stage) {
      const int Ns = stage == 0 ? 1 : (stage == 1 ? 8 : 64);
#pragma unroll
      for (int r = 0; r < 8; ++r) v[r] = sm.buf[q][j + r * 64];
      if (stage > 0) {
        const int k = (j & (Ns - 1)) * (64 / Ns);
#pragma unroll
        for (int r = 1; r < 8; ++r) v[r] = cmul(v[r], __ldg(&d_tw512[k * r]));
      }
      fft8(v);
      __syncthreads();
      const int base = (j / Ns) * Ns * 8 + (j & (Ns - 1));
#pragma unroll
      for (int r = 0; r < 8; ++r) sm.buf[q][base + r * Ns] = v[r];
      __syncthreads();
    }
  }
  // split Z = A + iB into the spectra of the two real frames, round to complex64, |.| and angle
  {
    float rm[5][2], rp[5][2];
#pragma unroll
    for (int it = 0; it < 5; ++it) {
      const int e = tid + it * 256;
      if (e < 4 * kBins) {
        int q = e / kBins, k = e - q * kBins;
        double2 zk = sm.buf[q][k], zn = sm.buf[q][(kNfft - k) & (kNfft - 1)];
        float are = (float)(0.5 * (zk.x + zn.x)), aim = (float)(0.5 * (zk.y - zn.y));
        float bre = (float)(0.5 * (zk.y + zn.y)), bim = (float)(-0.5 * (zk.x - zn.x));
        if (k == 0 || k == kBins - 1) aim = 0.f, bim = 0.f;   // a real FFT returns +0 there
        if (!nonzero[2 * q]) are = aim = 0.f;
        if (!nonzero[2 * q + 1]) bre = bim = 0.f;
        // np.abs(complex64) = hypotf; glibc evaluates it as the rounded double sqrt of the exact sum of squares
        rm[it][0] = (float)sqrt((double)are * are + (double)aim * aim);
        rm[it][1] = (float)sqrt((double)bre * bre + (double)bim * bim);
        if (phase) {
          rp[it][0] = atan2f(aim, are);
          rp[it][1] = atan2f(bim, bre);
        }
      }
    }
    __syncthreads();
#pragma unroll
    for (int it = 0; it < 5; ++it) {
      const int e = tid + it * 256;
      if (e < 4 * kBins) {
        int q = e / kBins, k = e - q * kBins;
        sm.o.mag[2 * q][k] = rm[it][0];
        sm.o.mag[2 * q + 1][k] = rm[it][1];
        if (phase) {
          sm.o.ph[2 * q][k] = rp[it][0];
          sm.o.ph[2 * q + 1][k] = rp[it][1];
        }
      }
    }
  }
  __syncthreads();
  const int64_t fo = foff[u];
  const int64_t mbase = (int64_t)kBins * fo;
  for (int e = tid; e < kBins * kTileFrames; e += 256) {
    int k = e >> 3, f = e & 7;
    int t = t0 + f;
    if (t < T) {
      if (mag) mag[mbase + (int64_t)k * T + t] = sm.o.mag[f][k];
      if (phase) phase[mbase + (int64_t)k * T + t] = sm.o.ph[f][k];
    }
  }
  if (band) {
    __syncthreads();
    // square in place (float32, as X[iT, k] ** 2), then the 64 bands of each frame
    for (int e = tid; e < kBins * kTileFrames; e += 256) {
      int f = e / kBins, k = e - f * kBins;
      float a = sm.o.mag[f][k];
      sm.o.mag[f][k] = __fmul_rn(a, a);
    }
    __syncthreads();
    for (int e = tid; e < kFeatBands * kTileFrames; e += 256) {
      int f = e >> 6, b = e & 63;
      int t = t0 + f;
      if (t < T) band[(fo + t) * kFeatBands + b] = band_power(band_energy(sm.o.mag[f], b), power, normalize);
    }
  }
}

// ---------------------------------------------------------------------------------------- IMCRA
constexpr int kImcraThreads = 288;
constexpr int kImcraTile = 32;
constexpr int kIS = 15, kU = 8, kV = 15;   // noise_est/imcra.py:181,192,194 (imcra_est passes IS = 15)

struct ImcraSmem {
  float tile[kBins][kImcraTile + 1];
  double P[kBins + 2], I[kBins + 2], IP[kBins + 2];
};

// fsmooth (noise_est/imcra.py:336-337) with w = 1: sym_hanning(3) = [0.5, 1, 0.5], truncated at the
// edges and normalised per row (:262-272).  `a` has one guard cell on each side (index k + 1 = bin k).
__device__ __forceinline__ double fsmooth3(const double* a, int k) {
  if (k == 0) return (1.0 / 1.5) * a[1] + (0.5 / 1.5) * a[2];
  if (k == kBins - 1) return (0.5 / 1.5) * a[k] + (1.0 / 1.5) * a[k + 1];
  return 0.25 * a[k] + 0.5 * a[k + 1] + 0.25 * a[k + 2];
}

__global__ void __launch_bounds__(kImcraThreads, 3) feat_imcra_kernel(const float* __restrict__ mag,
                                                                    const int64_t* __restrict__ foff,
                                                                    const int32_t* __restrict__ lens, float power,
                                                                    int normalize, float* __restrict__ band,
                                                                    float* __restrict__ psd) {
  __shared__ ImcraSmem sm;
  const int u = blockIdx.x, k = threadIdx.x, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const bool live = k < kBins;
  const int T = 1 + lens[u] / kHop;
  const int64_t fo = foff[u];
  const float* M = mag + (int64_t)kBins * fo;
  float* PSD = psd ? psd + (int64_t)kBins * fo : nullptr;

  // imcra defaults (:181-228), imcra_est defaults (:492)
  const double alpha_s = 0.9, alpha_d = 0.85, Bmin = 3.2, Gamma0 = 4.6, Gamma1 = 3.0, zeta0 = 1.67, beta = 1.47;
  const double alpha_dd = 0.92, xi_min = 0.056234132519034911 /* 10 ** (-25 / 20) */, p_up = 0.9;

  double G = 1.0, Gamma = 1.0, Lam = 1e-6;
  float Lam32 = 0.f;
  double S = 0, tS = 0, Smin = 0, tSmin = 0, Smin_sw = 0, tSmin_sw = 0, ovLam = 0;
  double store[kU], tstore[kU];
#pragma unroll
  for (int i = 0; i < kU; ++i) store[i] = tstore[i] = 0.0;
  int jcnt = 0, ucnt = 0;
  if (k < kBins + 2) sm.P[k] = sm.I[k] = sm.IP[k] = 0.0;

  for (int tb = 0; tb < T; tb += kImcraTile) {
    __syncthreads();
    for (int row = warp; row < kBins; row += kImcraThreads / 32) {
      int t = tb + lane;
      sm.tile[row][lane] = t < T ? M[(int64_t)row * T + t] : 0.f;
    }
    __syncthreads();
    const int nt = min(kImcraTile, T - tb);
    for (int tt = 0; tt < nt; ++tt) {
      const int l = tb + tt;
      float P32 = 0.f;
      double P = 0, xi = 0;
      if (live) {
        float a = sm.tile[k][tt];
        P32 = __fmul_rn(a, a);
        P = (double)P32;
        // decision-directed a-priori SNR (:544-557)
        double xi_G = G * G * Gamma;
        Gamma = P / Lam;
        xi = alpha_dd * xi_G + (1.0 - alpha_dd) * fmax(Gamma - 1.0, 1e-6);
        xi = fmax(xi, xi_min);
        G = xi / (1.0 + xi);
        sm.P[k + 1] = P;
      }
      __syncthreads();
      double Sf = 0;
      if (live) {
        Sf = fsmooth3(sm.P, k);
        if (l == 0) {   // init_params (:339-360)
          S = tS = Smin = tSmin = Smin_sw = tSmin_sw = Sf;
          ovLam = P;
          Lam32 = P32;
        }
        S = alpha_s * S + (1.0 - alpha_s) * Sf;
        Smin = fmin(Smin, S);
        Smin_sw = fmin(Smin_sw, S);
      }
      if (l < kIS) {
        // initial segment (:384-399): python scalars times float32 arrays -- a float32 recursion
        if (live) {
          Lam32 = __fadd_rn(__fmul_rn(0.85f, Lam32), __fmul_rn(0.15f, P32));
          Lam = (double)Lam32;
        }
        __syncthreads();   // the next frame overwrites sm.P (the other branch has its own barrier)
      } else {
        double Iv = 0;
        if (live) {
          const double rmin = 1.0 / (Bmin * Smin);                    // :411-414, one reciprocal for both ratios
          double Gmin = P * rmin, zeta = S * rmin;
          Iv = (Gmin < Gamma0 && zeta < zeta0) ? 1.0 : 0.0;
          sm.I[k + 1] = Iv;
          sm.IP[k + 1] = Iv * P;
        }
        __syncthreads();
        if (live) {
          double norm = fsmooth3(sm.I, k), tSf = fsmooth3(sm.IP, k);   // :420-422
          if (norm > 0) tSf = tSf / norm;
          tS = alpha_s * tS + (1.0 - alpha_s) * tSf;
          tSmin = fmin(tSmin, tS);
          tSmin_sw = fmin(tSmin_sw, tS);
          const double trmin = 1.0 / (Bmin * tSmin);                  // :429-430
          double tG = P * trmin, tz = S * trmin;
          double q = 0.0;
          if (tz < zeta0) {
            if (tG <= 1.0) q = 1.0;
            else if (tG < Gamma1) q = (Gamma1 - tG) / (Gamma1 - 1.0);
          }
          double p = 0.0;                                              // post_speech_prob (:23-38)
          if (q < 1.0) {
            const double nu = Gamma * G;   // Gamma xi / (1 + xi), G = xi / (1 + xi) from above
            p = q > 0.0 ? (1.0 - q) / ((1.0 - q) + q * (1.0 + xi) * exp(-nu)) : 1.0;
          }
          p = fmin(p, p_up);
          double ta = alpha_d + (1.0 - alpha_d) * p;                   // :443-447
          ovLam = ta * ovLam + (1.0 - ta) * P;
          Lam = beta * ovLam;
        }
        if (++jcnt == kV) {   // minimum store (:451-482); a ring: the minimum does not depend on the order
          if (live) {
            const int slot = ucnt & (kU - 1), cnt = min(ucnt + 1, kU);
            double m1 = INFINITY, m2 = INFINITY;
#pragma unroll
            for (int i = 0; i < kU; ++i) {
              if (i == slot) store[i] = Smin_sw, tstore[i] = tSmin_sw;
              if (i < cnt) m1 = fmin(m1, store[i]), m2 = fmin(m2, tstore[i]);
            }
            Smin = m1;
            tSmin = m2;
            Smin_sw = S;
            tSmin_sw = tS;
          }
          jcnt = 0;
          ++ucnt;
        }
      }
      if (live) sm.tile[k][tt] = (float)Lam;   // N_PSD is float32 (:531,568); the cell has been consumed
    }
    __syncthreads();
    if (PSD) {
      for (int row = warp; row < kBins; row += kImcraThreads / 32) {
        int t = tb + lane;
        if (t < T) PSD[(int64_t)row * T + t] = sm.tile[row][lane];
      }
      __syncthreads();
    }
    if (band) {
      // noise band energies of the tile's frames, all threads: compute_band_E(np.sqrt(estPSD)) squares the
      // float32 root again (audio_util.py:445-447)
      if (live)
        for (int tt = 0; tt < nt; ++tt) {
          float r = sqrtf(sm.tile[k][tt]);
          sm.tile[k][tt] = __fmul_rn(r, r);
        }
      __syncthreads();
      for (int e = threadIdx.x; e < nt * kFeatBands; e += kImcraThreads) {
        const int tt = e >> 6, b = e & 63;
        double acc = 0.0;
        if (b > 0)
          for (int kk = c_gmt[b - 1]; kk < c_gmt[b]; ++kk) acc += (double)__fmul_rn(__ldg(&d_wnext[kk]), sm.tile[kk][tt]);
        if (b < kFeatBands - 1)
          for (int kk = c_gmt[b]; kk < c_gmt[b + 1]; ++kk) acc += (double)__fmul_rn(__ldg(&d_wown[kk]), sm.tile[kk][tt]);
        band[(fo + tb + tt) * kFeatBands + b] = band_power((float)acc, power, normalize);
      }
    }
  }
}

}  // namespace

// the tables live in __device__ memory, i.e. per device: one upload per device this process uses
static void feat_tables_ready(cudaStream_t s) {
  static bool ready_dev[64] = {};
  int dev = 0;
  cudaGetDevice(&dev);
  const bool ready = dev >= 0 && dev < 64 && ready_dev[dev];
  if (!ready) {
    static const int gmt[kFeatBands] = {0,  3,  4,  5,  6,  7,  8,  9,  10,  11,  12,  13,  14,  15,  16,  17,
                                        18, 19, 20, 21, 22, 23, 24, 25, 26,  28,  30,  32,  34,  36,  38,  41,
                                        43, 46, 49, 52, 55, 58, 62, 66, 70,  74,  79,  83,  88,  93,  99,  105,
                                        111, 117, 124, 131, 139, 147, 156, 165, 174, 184, 195, 206, 218, 230, 243, 257};
    float wown[kBins], wnext[kBins];
    for (int i = 0; i < kFeatBands - 1; ++i) {
      const int size = gmt[i + 1] - gmt[i];
      for (int j = 0; j < size; ++j) {
        const double frac = (double)j / size;
        wown[gmt[i] + j] = (float)(1.0 - frac);
        wnext[gmt[i] + j] = (float)frac;
      }
    }
    cudaMemcpyToSymbolAsync(d_wown, wown, sizeof(wown), 0, cudaMemcpyHostToDevice, s);
    cudaMemcpyToSymbolAsync(d_wnext, wnext, sizeof(wnext), 0, cudaMemcpyHostToDevice, s);
    static double2 tw[kNfft];
    for (int m = 0; m < kNfft; ++m) {
      // exact octant symmetries so that the table is as symmetric as pocketfft's
      const double a = 2.0 * 3.14159265358979323846 * m / kNfft;
      tw[m] = make_double2(m == 128 || m == 384 ? 0.0 : cos(a), m == 0 || m == 256 ? 0.0 : -sin(a));
    }
    cudaMemcpyToSymbolAsync(d_tw512, tw, sizeof(tw), 0, cudaMemcpyHostToDevice, s);
    cudaStreamSynchronize(s);
    if (dev >= 0 && dev < 64) ready_dev[dev] = true;
  }
}

// ------------------------------------------------------------------------------------ resynthesis
// The in-loop boundary of a GAN sampling round (train_nele.py:303-314): the generator's
// band energy gains alpha2 [T][64] of one utterance -> interp_band_gain (audio_util.py:98-115) -> sqrt -> times the
// clean STFT (Resyn, :84-96) -> librosa.istft(hop 256, win 512) (:60-65) -> the PCM-16 rounding of
// sf.write(..., 'PCM_16') (train_nele.py:313) -> + noise (audio_util.py:196), written at the clean signal's offsets so
// that nele_score_batch can take (clean, degraded) as device inputs.  One CTA = 8 consecutive frames of one utterance
// (four 512-point complex FFTs in FP64, as feat_stft) = 7 output hops: every output sample is the windowed overlap-add
// of exactly two frames divided by the summed squared window.  Utterances are processed at their own length (reflect
// padding at their own ends), which is what the reference does one file at a time.
constexpr int kResynHops = kTileFrames - 1;

__device__ __forceinline__ void fft512_x4(double2 (*buf)[kNfft], int tid) {
  const int q = tid >> 6, j = tid & 63;
  double2 v[8];
#pragma unroll
  for (int stage = 0; stage < 3; ++stage) {
    const int Ns = stage == 0 ? 1 : (stage == 1 ? 8 : 64);
#pragma unroll
    for (int r = 0; r < 8; ++r) v[r] = buf[q][j + r * 64];
    if (stage > 0) {
      const int k = (j & (Ns - 1)) * (64 / Ns);
#pragma unroll
      for (int r = 1; r < 8; ++r) v[r] = cmul(v[r], __ldg(&d_tw512[k * r]));
    }
    fft8(v);
    __syncthreads();
    const int base = (j / Ns) * Ns * 8 + (j & (Ns - 1));
#pragma unroll
    for (int r = 0; r < 8; ++r) buf[q][base + r * Ns] = v[r];
    __syncthreads();
  }
}

// sqrt(interp_band_gain(E)[k]) (audio_util.py:98-115, :91): g[k] = (1 - frac) E[i] + frac E[i + 1] inside band i; the bins of
// the last band (243..256) keep g = 1 (the reference's loop stops at NB_BANDS - 1); bins 0, 1 and 256 are forced
__device__ __forceinline__ double resyn_gain(const float* __restrict__ E, int k) {
  if (k <= 1) return 1e-2;          // sqrt(1e-4)
  if (k == kBins - 1) return 0.1;   // sqrt(1e-2)
  int i = 0;
  while (i + 1 < kFeatBands && c_gmt[i + 1] <= k) ++i;      // band whose bins contain k
  if (i >= kFeatBands - 1) return 1.0;
  const double frac = (double)(k - c_gmt[i]) / (double)(c_gmt[i + 1] - c_gmt[i]);
  return sqrt((1.0 - frac) * (double)E[i] + frac * (double)E[i + 1]);
}

__global__ void __launch_bounds__(256) feat_resyn_kernel(const float* __restrict__ clean, const float* __restrict__ noise,
                                                          const int64_t* __restrict__ offs, const int32_t* __restrict__ lens,
                                                          const int64_t* __restrict__ foff, const int2* __restrict__ tiles,
                                                          const float* __restrict__ alpha2, int pcm16,
                                                          float* __restrict__ enh, float* __restrict__ deg) {
  __shared__ __align__(16) double2 buf[4][kNfft];
  const int tid = threadIdx.x;
  const int2 tile = tiles[blockIdx.x];
  const int u = tile.x, t0 = tile.y;
  const int L = lens[u];
  const int T = 1 + L / kHop;
  const float* x = clean + offs[u];
  for (int e = tid; e < kTileFrames * kNfft; e += 256) {
    const int f = e >> 9, n = e & 511, t = t0 + f;
    double v = 0.0;
    if (t < T) {
      int p = t * kHop + n - kNfft / 2;
      if (p < 0) p = -p;
      if (p >= L) p = 2 * (L - 1) - p;
      v = (0.5 - 0.5 * __ldg(&d_tw512[n].x)) * (double)x[p];
    }
    reinterpret_cast<double*>(&buf[f >> 1][n])[f & 1] = v;
  }
  __syncthreads();
  fft512_x4(buf, tid);
  // split Z = A + iB, round the two spectra to complex64 (what librosa.stft returns), apply the gains and put
  // conj(A' + i B') back: a second forward transform then gives N conj(a' + i b')
  for (int e = tid; e < 4 * kBins; e += 256) {
    const int q = e / kBins, k = e - q * kBins, kn = (kNfft - k) & (kNfft - 1);
    const double2 zk = buf[q][k], zn = buf[q][kn];
    float are = (float)(0.5 * (zk.x + zn.x)), aim = (float)(0.5 * (zk.y - zn.y));
    float bre = (float)(0.5 * (zk.y + zn.y)), bim = (float)(-0.5 * (zk.x - zn.x));
    if (k == 0 || k == kBins - 1) aim = 0.f, bim = 0.f;
    const int ta = t0 + 2 * q, tb = ta + 1;
    const double ga = ta < T ? resyn_gain(alpha2 + (foff[u] + ta) * kFeatBands, k) : 0.0;
    const double gb = tb < T ? resyn_gain(alpha2 + (foff[u] + tb) * kFeatBands, k) : 0.0;
    const double ar = ga * (double)are, ai = ga * (double)aim, br = gb * (double)bre, bi = gb * (double)bim;
    // Z'[k] = A' + i B' = (ar - bi) + i (ai + br); Z'[N - k] = conj(A') + i conj(B') = (ar + bi) + i (br - ai)
    buf[q][k] = make_double2(ar - bi, -(ai + br));
    if (k != 0 && k != kBins - 1) buf[q][kn] = make_double2(ar + bi, -(br - ai));
  }
  __syncthreads();
  fft512_x4(buf, tid);
  // overlap-add of frames h (second half) and h + 1 (first half), summed squared window, PCM-16, + noise
  const int64_t o = offs[u];
  for (int e = tid; e < kResynHops * kHop; e += 256) {
    const int hh = e >> 8, p = e & 255, h = t0 + hh;
    if (h >= T - 1) continue;
    const int fa = hh, fb = hh + 1;
    const double2 ra = buf[fa >> 1][p + kHop], rb = buf[fb >> 1][p];
    const double ya = ((fa & 1) ? -ra.y : ra.x) * (1.0 / kNfft), yb = ((fb & 1) ? -rb.y : rb.x) * (1.0 / kNfft);
    const double wa = 0.5 - 0.5 * __ldg(&d_tw512[p + kHop].x), wb = 0.5 - 0.5 * __ldg(&d_tw512[p].x);
    // librosa.istft accumulates window * frame in a float32 buffer and divides by the float32 window sum
    const float acc = (float)(wa * ya) + (float)(wb * yb);
    const float ss = (float)(wa * wa) + (float)(wb * wb);
    float y = acc / ss;
    const int64_t sidx = (int64_t)h * kHop + p;
    if (enh && !(pcm16 & 2)) enh[o + sidx] = y;
    if (pcm16 & 1) y = fminf(fmaxf(rintf(y * 32768.f), -32768.f), 32767.f) * (1.f / 32768.f);
    if (enh && (pcm16 & 2)) enh[o + sidx] = y;   // what the reference reads back from the PCM-16 file (dataloader.py:59)
    if (deg) deg[o + sidx] = y + noise[o + sidx];
  }
}

int resyn_run(const float* clean, const float* noise, const int64_t* offs, const int32_t* lens, const int64_t* foff,
              const int2* tiles, int ntiles, const float* alpha2, int pcm16, float* enh, float* deg, KernelTimer* kt,
              cudaStream_t s) {
  feat_tables_ready(s);
  kt_begin(kt, "feat_resyn", s);
  feat_resyn_kernel<<<ntiles, 256, 0, s>>>(clean, noise, offs, lens, foff, tiles, alpha2, pcm16, enh, deg);
  kt_end(kt, s);
  return 1;
}

int features_run(const float* wav, const int64_t* offs, const int32_t* lens, const int64_t* foff, const int2* tiles,
                 int n, int ntiles, bool noise, float power, bool normalize, float* band, float* mag, float* phase,
                 float* psd, KernelTimer* kt, cudaStream_t s) {
  feat_tables_ready(s);
  int launches = 0;
  kt_begin(kt, "feat_stft", s);
  feat_stft_kernel<<<ntiles, 256, 0, s>>>(wav, offs, lens, foff, tiles, power, normalize ? 1 : 0,
                                                         noise ? nullptr : band, mag, phase);
  kt_end(kt, s);
  ++launches;
  if (noise) {
    kt_begin(kt, "feat_imcra", s);
    feat_imcra_kernel<<<n, kImcraThreads, 0, s>>>(mag, foff, lens, power, normalize ? 1 : 0, band, psd);
    kt_end(kt, s);
    ++launches;
  }
  return launches;
}

}  // namespace nele
