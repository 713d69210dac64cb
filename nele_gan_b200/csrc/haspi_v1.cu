// HASPI version 1 on sm_100a: haspi(x, fx, y, fy, HL, alpha) of pyHASPI/pyhaspi2.py:109-157.
//
// Same ear model as version 2 (haspi.cu: prep, control, shift kernels); the back-end differs:
//
//   haspi_ear_v1_kernel    per pair warp, lane = band: the main pass of the ear model with the
//                          basilar-membrane output switched on (pyhaspi2.py:899, 996-998,
//                          1086-1087, 1075-1077), threshold noise (eb_BMaddnoise, :1091-1095),
//                          delay alignment (:1238-1241).  The envelope is reduced on the fly to
//                          the windowed 192-sample block sums eb_EnvSmooth needs (:673-703); the
//                          BM motion goes to HBM once, [t][32] per signal.
//   haspi_melcor_kernel    per pair CTA: 16 ms smoothed envelopes from the block sums, loudness
//                          selection, cepstral projection and correlation (eb_melcor, :706-751)
//   haspi_bmcov_kernel     per (pair, segment) CTA: Hann-windowed, mean-removed segments of the
//                          BM motion, mean squares and the +-1 ms normalised cross-covariance
//                          maximum (eb_BMcovary, :550-659), lane = band, 7 lags per warp
//   haspi_lev3_kernel      per pair CTA: loudness histogram, low / mid / high thirds and the
//                          per-band average covariance of each (eb_3LevelCovary, :418-547)
//   haspi_v1_score_kernel  logsig(-9.047 + 14.816 CepCorr + 4.616 cov3[high]) (:146-155)
#include "host_tables.hpp"
#include "kernels.h"
#include "philox.cuh"

namespace nele {

constexpr int kSegWin = 384;    // round(16 ms * 24 kHz), even          (pyhaspi2.py:674-677)
constexpr int kSegHalf = 192;
constexpr int kMaxLag = 24;     // round(1 ms * 24 kHz)                  (pyhaspi2.py:554-555)

__constant__ float c1_win[kSegWin];         // np.hanning(384)
__constant__ float c1_lagw[kMaxLag + 1];    // 1 / xcorr(window, window)[|lag|]      (the table at :564)
__constant__ float c1_lagh[kMaxLag + 1];    // 1 / xcorr(halfwindow, halfwindow)     (the table at :571)
__constant__ float c1_norm[4];              // 1/sum(w), 1/sum(hw), 1/sum(w^2), 1/sum(hw^2)
__constant__ float c1_cepm[kBands * kNumCep];
__constant__ IhcConst c1_ihc;
__constant__ float c1_fsync5[kBands];        // eb_AveCovary2 fsync[4]: sqrt(fc^10 / (fc^10 + cf^10)), fc = 3500 Hz

__host__ __device__ inline int v1_nseg(int n24) {  // pyhaspi2.py:686-687
  if (n24 < kSegHalf) return 0;
  return 1 + n24 / kSegWin + (n24 - kSegHalf) / kSegWin;
}

// ------------------------------------------------------------------ ear
constexpr int kEar1Warps = 4;
constexpr int kEar1Chunk = 576;  // three 192-sample blocks; ring of two chunks per warp and signal

template <typename T>
__global__ void __launch_bounds__(kEar1Warps * 32) haspi_ear_v1_kernel(PairGeom g, HaspiBuffers b, HaspiV1Buffers v,
                                                                       int n_pairs) {
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const int pair = blockIdx.x * kEar1Warps + wib;
  if (pair >= n_pairs) return;
  const float* __restrict__ midx = b.mid + g.off24[pair];
  const float* __restrict__ midy = b.mid + b.tot24 + g.off24[pair];
  const int N = g.n24[pair];
  float* __restrict__ bmx = v.bm + g.off24[pair] * kBands;
  float* __restrict__ bmy = v.bm + (b.tot24 + g.off24[pair]) * kBands;
  const int64_t ob = g.offblk[pair];
  float* __restrict__ srx = v.segsum + (0 * v.totblk + ob) * kBands;  // rising-half sums of x
  float* __restrict__ sfx = v.segsum + (1 * v.totblk + ob) * kBands;  // falling-half sums of x
  float* __restrict__ sry = v.segsum + (2 * v.totblk + ob) * kBands;
  float* __restrict__ sfy = v.segsum + (3 * v.totblk + ob) * kBands;
  extern __shared__ __align__(16) unsigned char s_ring_raw[];
  T* ringx = reinterpret_cast<T*>(s_ring_raw) + (size_t)wib * 4 * kEar1Chunk;
  T* ringy = ringx + 2 * kEar1Chunk;

  EarLane<T> Lx, Ly;
  Carrier<T> car;
  int shift;
  {
    const BandConst bc = b.bands[lane];
    shift = b.shift[(int64_t)pair * kBands + lane];
    car.init(bc.cf);
    Lx.init(bc, 0, b.bw[((int64_t)pair * 2 + 0) * kBands + lane], c1_ihc);
    Ly.init(bc, 1, b.bw[((int64_t)pair * 2 + 1) * kBands + lane], c1_ihc);
  }
  const float gn = 1.7782794100389228e-4f;  // 10^((-10 - 65) / 20)  (pyhaspi2.py:1092, 1231)
  const bool noisy = !b.no_dither;
  const uint64_t gp = (uint64_t)(b.pair_base + pair);
  int rp = (shift == 0) ? 0 : 2 * kEar1Chunk - shift;
  double pw_x = 0.0, pw_y = 0.0;  // HASQI: sum of the squared signal-path envelope (before compression)
  const int nchunks = (N + kEar1Chunk - 1) / kEar1Chunk;
  for (int c = 0; c < nchunks; ++c) {
    const int i0 = c * kEar1Chunk;
    T* hx = ringx + (c & 1) * kEar1Chunk;
    T* hy = ringy + (c & 1) * kEar1Chunk;
    __syncwarp();
#pragma unroll
    for (int k = 0; k < kEar1Chunk / 32; ++k) {
      const int t = i0 + k * 32 + lane;
      hx[k * 32 + lane] = (t < N) ? (T)midx[t] : (T)0;
      hy[k * 32 + lane] = (t < N) ? (T)midy[t] : (T)0;
    }
    __syncwarp();
    {
      const int t = i0 - shift;
      if (t >= 0) car.seed_before(t);
      else if (t + kEar1Chunk > 0) car.seed_before(0);
    }
    for (int blk = 0; blk < 3; ++blk) {
      const int ib = i0 + blk * kSegHalf;
      if (ib >= N) break;
      float ar_x = 0.f, af_x = 0.f, ar_y = 0.f, af_y = 0.f, pwp_x = 0.f, pwp_y = 0.f;
#pragma unroll 4
      for (int p = 0; p < kSegHalf; ++p) {
        const int i = ib + p;
        const T xs = ringx[rp], ys = ringy[rp];
        rp = (rp + 1 == 2 * kEar1Chunk) ? 0 : rp + 1;
        float vx = 0.f, vy = 0.f, bx = 0.f, by = 0.f;
        if (i >= shift && i < N) {
          car.advance();
          float px, py;
          vx = Lx.sample_bm(xs * car.c, xs * car.s, car.c, car.s, bx, px);
          vy = Ly.sample_bm(ys * car.c, ys * car.s, car.c, car.s, by, py);
          pwp_x += px;
          pwp_y += py;
          if (noisy) {  // added before the delay compensation: the zero-filled head stays silent
            float z0, z1;
            philox_normal2(b.seed ^ 0x9e3779b97f4a7c15ull, gp, (uint32_t)(i - shift), (uint32_t)lane, z0, z1);
            bx = fmaf(gn, z0, bx);
            by = fmaf(gn, z1, by);
          }
        }
        if (i < N) {
          bmx[(int64_t)i * kBands + lane] = bx;
          bmy[(int64_t)i * kBands + lane] = by;
        }
        const float wr = c1_win[p], wf = c1_win[kSegHalf + p];
        ar_x = fmaf(wr, vx, ar_x);
        af_x = fmaf(wf, vx, af_x);
        ar_y = fmaf(wr, vy, ar_y);
        af_y = fmaf(wf, vy, af_y);
      }
      const int64_t m = (int64_t)(ib / kSegHalf) * kBands + lane;
      srx[m] = ar_x;
      sfx[m] = af_x;
      sry[m] = ar_y;
      sfy[m] = af_y;
      pw_x += (double)pwp_x;
      pw_y += (double)pwp_y;
    }
  }
  if (v.hasqi) {
    // eb_EarModel averages the envelope over all N samples of the band (pyhaspi2.py:1211-1212);
    // the delayed time axis of this lane stopped `shift` samples short: run the signal path on.
    float tx = 0.f, ty = 0.f;
    for (int t = N - shift; t < N; ++t) {
      car.advance();
      const T xs = (T)midx[t], ys = (T)midy[t];
      tx += (float)Lx.fs.step(Lx.ks, xs * car.c, xs * car.s);
      ty += (float)Ly.fs.step(Ly.ks, ys * car.c, ys * car.s);
    }
    pw_x += (double)tx;
    pw_y += (double)ty;
    v.ave[((int64_t)pair * 2 + 0) * kBands + lane] = (double)Lx.ks.gain * sqrt(pw_x / (double)N);
    v.ave[((int64_t)pair * 2 + 1) * kBands + lane] = (double)Ly.ks.gain * sqrt(pw_y / (double)N);
  }
}

// smoothed envelope of segment s from the block sums (pyhaspi2.py:692-700): segment 0 is the
// falling half window over block 0, segment nseg - 1 the rising half over its own block, the
// others rise over block s and fall over block s + 1
__device__ __forceinline__ float seg_value(const float* __restrict__ rise, const float* __restrict__ fall, int s,
                                           int nseg, int lane) {
  if (s == 0) return fall[lane] * c1_norm[1];
  if (s == nseg - 1) return rise[(int64_t)s * kBands + lane] * c1_norm[1];
  return (rise[(int64_t)s * kBands + lane] + fall[(int64_t)(s + 1) * kBands + lane]) * c1_norm[0];
}

// --------------------------------------------------------------- melcor
constexpr int kMelThreads = 256;

__global__ void __launch_bounds__(kMelThreads) haspi_melcor_kernel(PairGeom g, HaspiV1Buffers v) {
  const int pair = blockIdx.x, tid = threadIdx.x, lane = tid & 31, wib = tid >> 5;
  constexpr int NW = kMelThreads / 32;
  const int nseg = v1_nseg(g.n24[pair]);
  const int64_t ob = g.offblk[pair];
  const float* __restrict__ srx = v.segsum + (0 * v.totblk + ob) * kBands;
  const float* __restrict__ sfx = v.segsum + (1 * v.totblk + ob) * kBands;
  const float* __restrict__ sry = v.segsum + (2 * v.totblk + ob) * kBands;
  const float* __restrict__ sfy = v.segsum + (3 * v.totblk + ob) * kBands;
  __shared__ double s_acc[NW][5 * kNumCep + 1];
  float cm[kNumCep];
#pragma unroll
  for (int j = 0; j < kNumCep; ++j) cm[j] = c1_cepm[lane * kNumCep + j];
  double acc[5 * kNumCep];
#pragma unroll
  for (int j = 0; j < 5 * kNumCep; ++j) acc[j] = 0.0;
  int cnt = 0;
  for (int s = wib; s < nseg; s += NW) {
    const float xs = seg_value(srx, sfx, s, nseg, lane), ys = seg_value(sry, sfy, s, nseg, lane);
    const float lin = warp_sum(undb20(xs));
    if (!(db20(lin * (1.0f / kBands)) > 2.5f)) continue;  // pyhaspi2.py:716-720
    ++cnt;
#pragma unroll
    for (int j = 0; j < kNumCep; ++j) {
      const double px = (double)warp_sum(xs * cm[j]), py = (double)warp_sum(ys * cm[j]);
      acc[5 * j + 0] += px;
      acc[5 * j + 1] += py;
      acc[5 * j + 2] += px * px;
      acc[5 * j + 3] += py * py;
      acc[5 * j + 4] += px * py;
    }
  }
  if (lane == 0) {
    for (int j = 0; j < 5 * kNumCep; ++j) s_acc[wib][j] = acc[j];
    s_acc[wib][5 * kNumCep] = (double)cnt;
  }
  __syncthreads();
  if (tid == 0) {
    double t[5 * kNumCep + 1];
    for (int j = 0; j <= 5 * kNumCep; ++j) {
      t[j] = 0.0;
      for (int w = 0; w < NW; ++w) t[j] += s_acc[w][j];
    }
    const double n = t[5 * kNumCep];
    if (n <= 1.0) {  // pyhaspi2.py:722-723 raises
      v.cepcorr[pair] = nan("");
      v.status[pair] = 1;
    } else {
      double m1 = 0.0;
      for (int j = 0; j < kNumCep; ++j) {
        const double xs = t[5 * j + 2] - t[5 * j] * t[5 * j] / n, ys = t[5 * j + 3] - t[5 * j + 1] * t[5 * j + 1] / n;
        const double xy = t[5 * j + 4] - t[5 * j] * t[5 * j + 1] / n;
        m1 += (xs < 1.0e-30 || ys < 1.0e-30) ? 0.0 : fabs(xy / sqrt(xs * ys));
      }
      v.cepcorr[pair] = m1 / (double)kNumCep;
      v.status[pair] = 0;
    }
  }
}

// ---------------------------------------------------------------- BM covariance
constexpr int kCovLagWarps = 7;                         // 7 warps x 7 lags = 49
constexpr int kCovThreads = (kCovLagWarps + 1) * 32;    // + one warp for the mean squares
constexpr int kCovPadRows = kSegWin + 2 * kMaxLag;      // x is zero padded by the lag reach on both sides

__global__ void __launch_bounds__(kCovThreads) haspi_bmcov_kernel(PairGeom g, HaspiBuffers b, HaspiV1Buffers v) {
  const int pair = blockIdx.y, seg = blockIdx.x, tid = threadIdx.x, lane = tid & 31, wib = tid >> 5;
  constexpr int NW = kCovThreads / 32;
  const int N = g.n24[pair];
  const int nseg = v1_nseg(N);
  if (seg >= nseg) return;
  extern __shared__ __align__(16) float s_cov[];
  float* sx = s_cov;                          // [kCovPadRows][32], row r <-> sample r - 24
  float* sy = s_cov + kCovPadRows * kBands;   // [kSegWin][32]
  __shared__ float s_part[NW][2][kBands];
  __shared__ float s_max[kCovLagWarps][kBands];
  // segment geometry (pyhaspi2.py:586-640): first = falling half window over [0, 192), last =
  // rising half over its 192 samples, the others the full window over [192 s, 192 s + 384)
  const bool first = seg == 0, last = seg == nseg - 1, half = first || last;
  const int start = first ? 0 : seg * kSegHalf, len = half ? kSegHalf : kSegWin;
  const int woff = first ? kSegHalf : 0;
  const float* __restrict__ gx = v.bm + (g.off24[pair] + start) * kBands;
  const float* __restrict__ gy = v.bm + (b.tot24 + g.off24[pair] + start) * kBands;
  // windowed load, 128-bit; partial sums for the means
  float4 px = make_float4(0.f, 0.f, 0.f, 0.f), py = px;
  for (int idx = tid; idx < kCovPadRows * (kBands / 4); idx += kCovThreads) {
    const int row = idx / (kBands / 4) - kMaxLag, c4 = idx % (kBands / 4);
    float4 a = make_float4(0.f, 0.f, 0.f, 0.f), c = a;
    if (row >= 0 && row < len) {
      const float w = c1_win[woff + row];
      a = *reinterpret_cast<const float4*>(gx + (int64_t)row * kBands + 4 * c4);
      c = *reinterpret_cast<const float4*>(gy + (int64_t)row * kBands + 4 * c4);
      a.x *= w; a.y *= w; a.z *= w; a.w *= w;
      c.x *= w; c.y *= w; c.z *= w; c.w *= w;
      px.x += a.x; px.y += a.y; px.z += a.z; px.w += a.w;
      py.x += c.x; py.y += c.y; py.z += c.z; py.w += c.w;
    }
    *reinterpret_cast<float4*>(sx + (row + kMaxLag) * kBands + 4 * c4) = a;
    if (row >= 0 && row < kSegWin) *reinterpret_cast<float4*>(sy + row * kBands + 4 * c4) = c;
  }
  // a thread always visits the same column group c4 = tid % 8 (kCovThreads is a multiple of 8):
  // reduce the per-thread partial sums over the threads that share c4
  __shared__ float s_colsum[2][kCovThreads][4];
  s_colsum[0][tid][0] = px.x; s_colsum[0][tid][1] = px.y; s_colsum[0][tid][2] = px.z; s_colsum[0][tid][3] = px.w;
  s_colsum[1][tid][0] = py.x; s_colsum[1][tid][1] = py.y; s_colsum[1][tid][2] = py.z; s_colsum[1][tid][3] = py.w;
  __syncthreads();
  __shared__ float s_mean[2][kBands];
  if (tid < 2 * kBands) {
    const int q = tid / kBands, band = tid % kBands, c4 = band / 4, e = band % 4;
    float s = 0.f;
    for (int t = c4; t < kCovThreads; t += kBands / 4) s += s_colsum[q][t][e];
    s_mean[q][band] = s / (float)len;
  }
  __syncthreads();
  // remove the means (inside the segment only: the padding stays zero) and sum the squares
  {
    const float mx = s_mean[0][lane], my = s_mean[1][lane];
    float qx = 0.f, qy = 0.f;
    for (int row = wib; row < len; row += NW) {
      const float a = sx[(row + kMaxLag) * kBands + lane] - mx, c = sy[row * kBands + lane] - my;
      sx[(row + kMaxLag) * kBands + lane] = a;
      sy[row * kBands + lane] = c;
      qx = fmaf(a, a, qx);
      qy = fmaf(c, c, qy);
    }
    s_part[wib][0][lane] = qx;
    s_part[wib][1][lane] = qy;
  }
  __syncthreads();
  // lag products: warp w covers lags -24 + 7 w .. -24 + 7 w + 6, lane = band.
  // corr[l] = sum_n x[n + l] y[n]  (np.correlate(segx, segy, 'full') centred, pyhaspi2.py:596-599)
  if (wib < kCovLagWarps) {
    const int l0 = 7 * wib;  // row offset of x for r = 0 (lag l0 - 24, padded by 24)
    float acc[7] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    float w[7];
    const float* xp = sx + l0 * kBands + lane;
    const float* yp = sy + lane;
#pragma unroll
    for (int r = 0; r < 6; ++r) w[r] = xp[r * kBands];
    int n = 0;
    for (; n + 7 <= len; n += 7) {
#pragma unroll
      for (int u = 0; u < 7; ++u) {
        // logical window element r lives in w[(u + r) % 7]
        w[(u + 6) % 7] = xp[(n + u + 6) * kBands];
        const float yv = yp[(n + u) * kBands];
#pragma unroll
        for (int r = 0; r < 7; ++r) acc[r] = fmaf(w[(u + r) % 7], yv, acc[r]);
      }
    }
    for (; n < len; ++n) {  // remainder (len = 384 = 7 * 54 + 6, 192 = 7 * 27 + 3); layout is identity here
      w[6] = xp[(n + 6) * kBands];
      const float yv = yp[n * kBands];
#pragma unroll
      for (int r = 0; r < 7; ++r) acc[r] = fmaf(w[r], yv, acc[r]);
#pragma unroll
      for (int r = 0; r < 6; ++r) w[r] = w[r + 1];
    }
    float m = 0.f;
#pragma unroll
    for (int r = 0; r < 7; ++r) {
      const int lag = l0 + r - kMaxLag, al = lag < 0 ? -lag : lag;
      m = fmaxf(m, fabsf(acc[r] * (half ? c1_lagh[al] : c1_lagw[al])));
    }
    s_max[wib][lane] = m;
  }
  __syncthreads();
  if (wib == 0) {
    float qx = 0.f, qy = 0.f, m = 0.f;
#pragma unroll
    for (int w = 0; w < NW; ++w) {
      qx += s_part[w][0][lane];
      qy += s_part[w][1][lane];
    }
#pragma unroll
    for (int w = 0; w < kCovLagWarps; ++w) m = fmaxf(m, s_max[w][lane]);
    const float nrm = half ? c1_norm[3] : c1_norm[2];
    const float msx = qx * nrm, msy = qy * nrm;
    float cv = (msx > 1.0e-30f && msy > 1.0e-30f) ? m * rsqrtf(msx) * rsqrtf(msy) : 0.f;
    cv = fminf(fmaxf(cv, 0.f), 1.f);
    const int64_t o = (g.offblk[pair] + seg) * kBands + lane;
    v.cov[o] = cv;
    v.msx[o] = 2.0f * msx;
  }
}

// --------------------------------------------------------- three-level covariance
constexpr int kLevThreads = 256;
constexpr int kLevMaxBins = 1024;

__global__ void __launch_bounds__(kLevThreads) haspi_lev3_kernel(PairGeom g, HaspiV1Buffers v) {
  const int pair = blockIdx.x, tid = threadIdx.x, lane = tid & 31, wib = tid >> 5;
  constexpr int NW = kLevThreads / 32;
  const int nseg = v1_nseg(g.n24[pair]);
  const int64_t ob = g.offblk[pair];
  const float* __restrict__ cov = v.cov + ob * kBands;
  const float* __restrict__ msx = v.msx + ob * kBands;
  double* __restrict__ xsum = v.xsum + ob;
  __shared__ double red[32];
  __shared__ int s_hist[kLevMaxBins];
  __shared__ double s_edge[2];
  __shared__ double s_ss[NW][3][kBands], s_ws[NW][3][kBands];
  // loudness of every segment (pyhaspi2.py:423-428); -inf marks "not selected"
  double lo = 1.0e300, hi = -1.0e300;
  int cnt = 0;
  for (int s = wib; s < nseg; s += NW) {
    const double rms = sqrt((double)msx[(int64_t)s * kBands + lane]);
    const double lin = warp_sum(pow(10.0, rms / 20.0));
    const double xs = 20.0 * log10(lin / (double)kBands);
    const bool sel = xs > 2.5;
    if (lane == 0) xsum[s] = sel ? xs : -1.0e300;
    if (sel) {
      lo = fmin(lo, xs);
      hi = fmax(hi, xs);
      ++cnt;
    }
  }
  for (int i = tid; i < kLevMaxBins; i += kLevThreads) s_hist[i] = 0;
  const double dbmax = block_max(hi, red);
  const double dbmin = -block_max(-lo, red);
  const int nsel = (int)(block_sum((double)cnt, red) / 32.0 + 0.5);  // every lane of a warp counted
  if (nsel <= 1) {  // pyhaspi2.py:430-431 raises
    if (tid == 0) {
      v.cov3[3 * pair + 0] = v.cov3[3 * pair + 1] = v.cov3[3 * pair + 2] = nan("");
      v.status[pair] = 1;
    }
    return;
  }
  // histogram with 0.5 dB bins centred on dbmin + 0.5 j (pyhaspi2.py:448-458)
  int nbins = (int)ceil((dbmax + 0.5 - dbmin) / 0.5);  // len(np.arange(dBmin, dBmax + dBstep, dBstep))
  nbins = min(max(nbins, 1), kLevMaxBins);
  __syncthreads();
  for (int s = tid; s < nseg; s += kLevThreads) {
    const double xs = xsum[s];
    if (xs < -1.0e299) continue;
    int j = (int)floor((xs - dbmin) / 0.5 + 0.5);
    j = min(max(j, 0), nbins - 1);
    // exact edges: bin j = [e_j, e_{j+1}), e_j = (bins[j-1] + bins[j]) / 2
    while (j + 1 <= nbins - 1 && xs >= ((dbmin + 0.5 * j) + (dbmin + 0.5 * (j + 1))) * 0.5) ++j;
    while (j >= 1 && xs < ((dbmin + 0.5 * (j - 1)) + (dbmin + 0.5 * j)) * 0.5) --j;
    atomicAdd(&s_hist[j], 1);
  }
  __syncthreads();
  if (tid == 0) {  // cumulative histogram -> boundaries of the thirds (pyhaspi2.py:461-475)
    double e0 = 0.0, e1 = 0.0;
    int cum = 0;
    for (int j = 0; j < nbins; ++j) {
      cum += s_hist[j];
      const double c = (double)cum / (double)nsel;
      if (c < 0.333) e0 = dbmin + 0.5 * j;
      if (c < 0.667) e1 = dbmin + 0.5 * j;
    }
    s_edge[0] = e0;
    s_edge[1] = e1;
  }
  __syncthreads();
  const double e0 = s_edge[0], e1 = s_edge[1];
  double ss[3] = {0.0, 0.0, 0.0}, ws[3] = {0.0, 0.0, 0.0};
  double f5 = 0.0, s5 = 0.0;  // eb_AveCovary2 (pyhaspi2.py:161-218): fsync-weighted covariance of the audible cells
  const double fw = (double)c1_fsync5[lane];
  for (int s = wib; s < nseg; s += NW) {
    const double xs = xsum[s];
    if (xs < -1.0e299) continue;
    const int grp = (xs < e0) ? 0 : (xs < e1) ? 1 : 2;
    const double rms = sqrt((double)msx[(int64_t)s * kBands + lane]);
    if (rms > 2.5) {  // pyhaspi2.py:486-488
      const double c = (double)cov[(int64_t)s * kBands + lane];
      f5 += fw * c;
      s5 += fw;
#pragma unroll
      for (int k = 0; k < 3; ++k)
        if (k == grp) {
          ss[k] += c;
          ws[k] += 1.0;
        }
    }
  }
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    s_ss[wib][k][lane] = ss[k];
    s_ws[wib][k][lane] = ws[k];
  }
  __syncthreads();
  if (wib < 3) {
    double a = 0.0, w = 0.0;
    for (int u = 0; u < NW; ++u) {
      a += s_ss[u][wib][lane];
      w += s_ws[u][wib][lane];
    }
    const double ave = (w != 0.0) ? a / w : 0.0;
    const double tot = warp_sum(ave);
    const double nc = warp_sum((w != 0.0) ? 1.0 : 0.0);
    if (lane == 0) v.cov3[3 * pair + wib] = tot / nc;  // 0 / 0 = NaN like the reference
  }
  if (v.hasqi) {
    const double a = block_sum(f5, red), c = block_sum(s5, red);
    if (tid == 0) v.sync5[pair] = a / c;
  }
}

__global__ void haspi_v1_score_kernel(HaspiV1Buffers v, int n, double* intel, double* raw10, int32_t* status) {
  const int pair = blockIdx.x * blockDim.x + threadIdx.x;
  if (pair >= n) return;
  double* raw = raw10 + (int64_t)pair * kNumMod;
  for (int m = 0; m < kNumMod; ++m) raw[m] = nan("");
  if (v.status[pair] != 0) {
    intel[pair] = nan("");
    status[pair] = 1;
    return;
  }
  const double cep = v.cepcorr[pair];
  const double c0 = v.cov3[3 * pair + 0], c1 = v.cov3[3 * pair + 1], c2 = v.cov3[3 * pair + 2];
  const double arg = -9.047 + 14.816 * cep + 4.616 * c2;  // pyhaspi2.py:146-149
  intel[pair] = 1.0 / (1.0 + exp(-arg));                   // alpha = -1 (:152)
  raw[0] = cep;
  raw[1] = c0;
  raw[2] = c1;
  raw[3] = c2;
  status[pair] = 0;
}

// HASQI version 2 (pyhaspi2.py:32-74): one warp per pair, lane = band
__global__ void __launch_bounds__(128) hasqi_score_kernel(HaspiBuffers b, HaspiV1Buffers v, int n, double* out, double* raw10,
                                                           int32_t* status) {
  const int pair = blockIdx.x * 4 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (pair >= n) return;
  double* raw = raw10 + (int64_t)pair * kNumMod;
  if (v.status[pair] != 0) {
    if (lane == 0) {
      for (int m = 0; m < kNumMod; ++m) raw[m] = nan("");
      out[pair] = nan("");
      status[pair] = 1;
    }
    return;
  }
  const BandConst bc = b.bands[lane];
  double lin[2];
#pragma unroll
  for (int q = 0; q < 2; ++q) {  // eb_aveSL (pyhaspi2.py:1135-1152); itype = 2: x gets the audiogram of y (:1162-1165)
    const double env = v.ave[((int64_t)pair * 2 + q) * kBands + lane], ctl = b.cave[((int64_t)pair * 2 + q) * kBands + lane];
    double le = 65.0 + 20.0 * log10(fmax(ctl, 1.0e-30));
    le = fmin(fmax(le, bc.lowknee[1]), 100.0);
    const double gain = -bc.attn_ohc[1] - (le - bc.lowknee[1]) * (1.0 - 1.0 / bc.cr[1]);
    double ls = fmax(65.0 + 20.0 * log10(fmax(env, 1.0e-30)), 0.0);
    const double sl = fmax(ls + gain - bc.attn_ihc[1], 0.0);
    lin[q] = pow(10.0, sl / 20.0);
  }
  // eb_SpectDiff (pyhaspi2.py:220-251): dloud[1] = nbands std(x - y), dslope[1] = nbands std of the first differences
  const double x = lin[0] / warp_sum(lin[0]), y = lin[1] / warp_sum(lin[1]);
  const double d = x - y;
  const double md = warp_sum(d) / kBands;
  const double dloud = kBands * sqrt(warp_sum((d - md) * (d - md)) / kBands);
  const double dn = __shfl_down_sync(0xffffffffu, d, 1) - d;  // (x[i+1] - x[i]) - (y[i+1] - y[i]), lanes 0..30
  const double e = (lane < kBands - 1) ? dn : 0.0;
  const double me = warp_sum(e) / (kBands - 1);
  const double dslope = kBands * sqrt(warp_sum((lane < kBands - 1) ? (e - me) * (e - me) : 0.0) / (kBands - 1));
  if (lane == 0) {
    const double cep = v.cepcorr[pair], sync5 = v.sync5[pair];
    const double Dloud = fmin(fmax(1.0 - dloud / 2.5, 0.0), 1.0), Dslope = fmin(fmax(1.0 - dslope, 0.0), 1.0);
    const double nonlin = cep * cep * sync5, linear = 0.579 * Dloud + 0.421 * Dslope;
    out[pair] = nonlin * linear;
    raw[0] = cep;
    raw[1] = sync5;
    raw[2] = Dloud;
    raw[3] = Dslope;
    raw[4] = nonlin;
    raw[5] = linear;
    for (int m = 6; m < kNumMod; ++m) raw[m] = nan("");
    status[pair] = 0;
  }
}

// ------------------------------------------------------------- launchers
void haspi_v1_upload_tables(const float* cepm, cudaStream_t s) {
  double w[kSegWin];
  float wf[kSegWin];
  for (int k = 0; k < kSegWin; ++k) {
    w[k] = 0.5 - 0.5 * cos(2.0 * host::kPi * k / (kSegWin - 1));  // np.hanning(384)
    wf[k] = (float)w[k];
  }
  float lagw[kMaxLag + 1], lagh[kMaxLag + 1], norm[4];
  for (int l = 0; l <= kMaxLag; ++l) {
    double a = 0.0, h = 0.0;
    for (int k = 0; k + l < kSegWin; ++k) a += w[k] * w[k + l];
    for (int k = kSegHalf; k + l < kSegWin; ++k) h += w[k] * w[k + l];
    lagw[l] = (float)(1.0 / a);
    lagh[l] = (float)(1.0 / h);
  }
  double sw = 0.0, sh = 0.0, sw2 = 0.0, sh2 = 0.0;
  for (int k = 0; k < kSegWin; ++k) {
    sw += w[k];
    sw2 += w[k] * w[k];
    if (k >= kSegHalf) {
      sh += w[k];
      sh2 += w[k] * w[k];
    }
  }
  norm[0] = (float)(1.0 / sw);
  norm[1] = (float)(1.0 / sh);
  norm[2] = (float)(1.0 / sw2);
  norm[3] = (float)(1.0 / sh2);
  cudaMemcpyToSymbolAsync(c1_win, wf, sizeof(wf), 0, cudaMemcpyHostToDevice, s);
  cudaMemcpyToSymbolAsync(c1_lagw, lagw, sizeof(lagw), 0, cudaMemcpyHostToDevice, s);
  cudaMemcpyToSymbolAsync(c1_lagh, lagh, sizeof(lagh), 0, cudaMemcpyHostToDevice, s);
  cudaMemcpyToSymbolAsync(c1_norm, norm, sizeof(norm), 0, cudaMemcpyHostToDevice, s);
  cudaMemcpyToSymbolAsync(c1_cepm, cepm, sizeof(float) * kBands * kNumCep, 0, cudaMemcpyHostToDevice, s);
  const IhcConst ih = make_ihc_const();
  cudaMemcpyToSymbolAsync(c1_ihc, &ih, sizeof(ih), 0, cudaMemcpyHostToDevice, s);
  double cf[kBands];
  host::center_freqs(cf);
  float fs5[kBands];
  for (int k = 0; k < kBands; ++k) {
    const double fc10 = pow(3500.0, 10.0);
    fs5[k] = (float)sqrt(fc10 / (fc10 + pow(cf[k], 10.0)));
  }
  cudaMemcpyToSymbolAsync(c1_fsync5, fs5, sizeof(fs5), 0, cudaMemcpyHostToDevice, s);
  cudaFuncSetAttribute(haspi_ear_v1_kernel<double>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                       (int)(kEar1Warps * 4 * kEar1Chunk * sizeof(double)));
  cudaFuncSetAttribute(haspi_bmcov_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                       (int)((kCovPadRows + kSegWin) * kBands * sizeof(float)));
  cudaStreamSynchronize(s);  // sources are stack temporaries
}

int haspi_v1_run(const PairGeom& g, const HaspiBuffers& b, const HaspiV1Buffers& v, int n, int max_n24, bool f64,
                 KernelTimer* kt, cudaStream_t s) {
  int launches = haspi_run_front(g, b, n, f64, kt, s);
  const int ctas = (n + kEar1Warps - 1) / kEar1Warps;
  kt_begin(kt, "haspi_ear_v1", s);
  if (f64) haspi_ear_v1_kernel<double><<<ctas, kEar1Warps * 32, kEar1Warps * 4 * kEar1Chunk * sizeof(double), s>>>(g, b, v, n);
  else haspi_ear_v1_kernel<float><<<ctas, kEar1Warps * 32, kEar1Warps * 4 * kEar1Chunk * sizeof(float), s>>>(g, b, v, n);
  kt_end(kt, s);
  ++launches;
  kt_begin(kt, "haspi_melcor", s);
  haspi_melcor_kernel<<<n, kMelThreads, 0, s>>>(g, v);
  kt_end(kt, s);
  ++launches;
  const int max_seg = v1_nseg(max_n24);
  if (max_seg > 0) {
    kt_begin(kt, "haspi_bmcov", s);
    haspi_bmcov_kernel<<<dim3(max_seg, n), kCovThreads, (kCovPadRows + kSegWin) * kBands * sizeof(float), s>>>(g, b, v);
    kt_end(kt, s);
    ++launches;
  }
  return launches;
}

int haspi_v1_finish(const PairGeom& g, const HaspiBuffers& b, const HaspiV1Buffers& v, int n, double* intel, double* raw10,
                    int32_t* status, KernelTimer* kt, cudaStream_t s) {
  kt_begin(kt, "haspi_lev3", s);
  haspi_lev3_kernel<<<n, kLevThreads, 0, s>>>(g, v);
  kt_end(kt, s);
  if (v.hasqi) {
    kt_begin(kt, "hasqi_score", s);
    hasqi_score_kernel<<<(n + 3) / 4, 128, 0, s>>>(b, v, n, intel, raw10, status);
    kt_end(kt, s);
  } else {
    kt_begin(kt, "haspi_v1_score", s);
    haspi_v1_score_kernel<<<(n + 127) / 128, 128, 0, s>>>(v, n, intel, raw10, status);
    kt_end(kt, s);
  }
  return 2;
}

}  // namespace nele
