// ESTOI on sm_100a: batched restatement of pystoi.stoi(x, y, fs, extended=True)
// (reference call sites intel.py:126,133; algorithm: oracle/pystoi_np.py).
//
//   estoi_resample_kernel  per (output tile, pair, signal) CTA: Octave-style polyphase
//                          resampler to 10 kHz (resample_poly(x, up, down, window=h)),
//                          inputs staged in shared memory, FP64 accumulation
//   estoi_vad_kernel       per pair CTA: Hann-windowed frame energies of x, 40 dB dynamic
//                          range mask, ordered compaction -> list of kept frames
//   estoi_tob_kernel       per (pair, STFT frame) warp: rebuilds the frame of the
//                          silence-removed signal from the three kept source frames that
//                          overlap it (the overlap-add is never materialised), one complex
//                          512-point FFT for x + i y in shared memory, one-third-octave
//                          band magnitudes [nfr][15] for both signals
//   estoi_corr_kernel      per pair CTA, one thread per 30-frame segment: row then column
//                          normalisation and the correlation sum
#include <algorithm>

#include <stdlib.h>

#include "kernels.h"

namespace nele {

constexpr int kStoiFrame = 256, kStoiHop = 128, kStoiFft = 512, kStoiBands = 15, kStoiSeg = 30;

__device__ float g_stoi_win[kStoiFrame];       // MATLAB hanning(256); staged in shared memory by its users
__constant__ int c_stoi_lo[kStoiBands], c_stoi_hi[kStoiBands];
__device__ float2 g_stoi_tw[kStoiFft / 2];     // exp(-2 pi i k / 512)

// ------------------------------------------------------------- resample
constexpr int kRsThreads = 256, kRsJ = 4;

// A thread owns kRsJ outputs of one polyphase branch, m, m + up, ... (same taps, inputs `down`
// apart): every tap is fetched once for kRsJ independent FP64 FMA chains.  A CTA covers
// tile = kRsJ * up * groups consecutive outputs with groups * up threads.
__global__ void __launch_bounds__(kRsThreads) estoi_resample_kernel(EstoiGeom g, EstoiBuffers b, int span, int groups,
                                                                    int taps_in_smem) {
  const int pair = blockIdx.y, q = blockIdx.z, tid = threadIdx.x;
  const int n_out = g.n10[pair];
  const int tile = kRsJ * b.up * groups;
  const int m0 = blockIdx.x * tile;
  if (m0 >= n_out) return;
  const float* __restrict__ src = (q == 0 ? b.ref : b.deg) + g.off16[pair];
  const int L = g.len16[pair];
  float* __restrict__ dst = b.x10 + (int64_t)q * b.tot10 + g.off10[pair];
  extern __shared__ __align__(16) unsigned char s_rs_raw[];
  if (b.up == 1 && b.down == 1) {
    for (int m = m0 + tid; m < min(m0 + tile, n_out); m += kRsThreads) dst[m] = src[m];
    return;
  }
  const int nt = 2 * b.K + 1;
  double* s_tap = reinterpret_cast<double*>(s_rs_raw);
  float* s_in = reinterpret_cast<float*>(s_rs_raw + (taps_in_smem ? (size_t)b.up * nt * sizeof(double) : 0));
  // staged index u <-> input sample base + u, base = n(m0) - K
  const int base = (int)(((int64_t)m0 * b.down) / b.up) - b.K;
  for (int u = tid; u < span; u += kRsThreads) {
    const int j = base + u;
    s_in[u] = (j >= 0 && j < L) ? src[j] : 0.f;
  }
  if (taps_in_smem)
    for (int u = tid; u < b.up * nt; u += kRsThreads) s_tap[u] = b.taps[u];
  __syncthreads();
  const int nthr = groups * b.up;  // outputs m, m + nthr, m + 2 nthr, ...: same branch (nthr is a multiple of up),
  if (tid >= nthr) return;         // and consecutive lanes read consecutive-ish inputs (no bank conflicts)
  const int m = m0 + tid;
  const int64_t pos = (int64_t)m * b.down;
  const int n = (int)(pos / b.up), r = (int)(pos % b.up);
  const double* __restrict__ tp = (taps_in_smem ? s_tap : b.taps) + (size_t)r * nt;
  const int sj = b.down * groups;  // input distance between the outputs of one thread
  const float* xs = s_in + (n - base) + b.K;  // xs[j sj - k'] with k' = k + K -> x[n + j sj - k]
  double acc[kRsJ];
#pragma unroll
  for (int j = 0; j < kRsJ; ++j) acc[j] = 0.0;
#pragma unroll 2
  for (int k = 0; k < nt; ++k) {
    const double t = tp[k];
#pragma unroll
    for (int j = 0; j < kRsJ; ++j) acc[j] = fma(t, (double)xs[j * sj - k], acc[j]);
  }
#pragma unroll
  for (int j = 0; j < kRsJ; ++j) {
    const int mj = m + j * nthr;
    if (mj < n_out) dst[mj] = (float)acc[j];
  }
}

// 16 -> 10 kHz fast path (up = 5, down = 8, 2K + 1 <= 128 taps per branch; K = 59 for pystoi's
// filter).  Warp = output phase (t mod 5, i.e. one polyphase branch: the tap of a step is one
// broadcast read), lane = group of 20 outputs; a thread owns the four outputs t = 20 v + phi + 5 j
// of its group, whose inputs sit 8 apart, so the inputs slide through a 32-register ring: per tap
// one tap load, one new input and four FP64 FMAs (the generic kernel re-reads and re-converts four
// inputs per tap).  Inputs are staged as FP64 with a (u + u / 32) skew -- lanes 32 samples apart
// hit distinct banks -- and the taps are zero padded to 128 so the ring indices are static.
constexpr int kRfG = 32, kRfThreads = 5 * kRfG, kRfTile = 20 * kRfG, kRfNt = 128;
constexpr int kRfSpan = 32 * kRfG + 126;
__constant__ double c_rf_tap[5][kRfNt];   // zero padded branches (estoi_upload_polytaps); a warp reads one entry per step
__constant__ float c_rf_tapf[5][kRfNt];   // the same in FP32 (the default; NELE_RESAMPLE_F32=0 selects FP64)

constexpr int kRfTilesPerCta = 4;
constexpr int kRfPre = (kRfSpan + kRfThreads - 1) / kRfThreads;   // staged inputs per thread and tile

// T = float (default): FP32 throughout -- 119 taps of FP32 accumulation leave ~1e-6 relative error in the 10 kHz
// signal (measured on a B200: ESTOI moves by 3.5e-7, 4.5 -> 3.2 ms per 4096 x 3 s).  T = double
// (NELE_RESAMPLE_F32=0): FP64 products and sums, kept for A/B checks.
template <typename T>
__global__ void __launch_bounds__(kRfThreads, 4) estoi_resample58_kernel(EstoiGeom g, EstoiBuffers b) {
  const int pair = blockIdx.y, q = blockIdx.z, tid = threadIdx.x;
  const int n_out = g.n10[pair];
  const int Tfirst = blockIdx.x * (kRfTilesPerCta * kRfTile);
  if (Tfirst >= n_out) return;
  const float* __restrict__ src = (q == 0 ? b.ref : b.deg) + g.off16[pair];
  const int L = g.len16[pair];
  float* __restrict__ dst = b.x10 + (int64_t)q * b.tot10 + g.off10[pair];
  __shared__ T s_in[kRfSpan + kRfSpan / 32 + 2];
  __shared__ float s_out[kRfTile];
  const int phi = tid >> 5, v = tid & 31;
  const T* __restrict__ tp;
  if constexpr (sizeof(T) == 8) tp = c_rf_tap[(8 * phi) % 5];
  else tp = c_rf_tapf[(8 * phi) % 5];
  const int u0 = 32 * v + (8 * phi) / 5 + (kRfNt - 1);   // staged index of x[n + K] for the first output
  // staged index i <-> x[inbase + i]; output t with n = floor(8 t / 5) reads x[n + K - kk], kk = 0..127.
  // The inputs of the next tile are fetched into registers while the current one is computed.
  float pre[kRfPre];
  auto fetch = [&](int T0) {
    const int inbase = (T0 / 5) * 8 + b.K - (kRfNt - 1);
#pragma unroll
    for (int c = 0; c < kRfPre; ++c) {
      const int j = inbase + tid + c * kRfThreads;
      pre[c] = (j >= 0 && j < L) ? __ldg(src + j) : 0.f;
    }
  };
  fetch(Tfirst);
  for (int tile = 0; tile < kRfTilesPerCta; ++tile) {
    const int T0 = Tfirst + tile * kRfTile;
    if (T0 >= n_out) break;
#pragma unroll
    for (int c = 0; c < kRfPre; ++c) {
      const int i = tid + c * kRfThreads;
      if (i < kRfSpan) s_in[i + (i >> 5)] = (T)pre[c];
    }
    __syncthreads();
    if (tile + 1 < kRfTilesPerCta && T0 + kRfTile < n_out) fetch(T0 + kRfTile);
    T w[32];
#pragma unroll
    for (int i = 1; i <= 24; ++i) {
      const int u = u0 + i;
      w[32 - i] = s_in[u + (u >> 5)];
    }
    T a0 = 0, a1 = 0, a2 = 0, a3 = 0;
#pragma unroll 1
    for (int kk0 = 0; kk0 < kRfNt; kk0 += 32) {
#pragma unroll
      for (int sft = 0; sft < 32; ++sft) {
        const int u = u0 - kk0 - sft;
        w[sft] = s_in[u + (u >> 5)];
        const T t = tp[kk0 + sft];
        a0 = fma(t, w[sft], a0);
        a1 = fma(t, w[(sft + 24) & 31], a1);   // loaded 8 steps ago: x 8 samples later
        a2 = fma(t, w[(sft + 16) & 31], a2);
        a3 = fma(t, w[(sft + 8) & 31], a3);
      }
    }
    const int tl = 20 * v + phi;
    s_out[tl] = (float)a0;
    s_out[tl + 5] = (float)a1;
    s_out[tl + 10] = (float)a2;
    s_out[tl + 15] = (float)a3;
    __syncthreads();   // also: every read of s_in is done before the next tile overwrites it
    for (int t = tid; t < kRfTile && T0 + t < n_out; t += kRfThreads) dst[T0 + t] = s_out[t];
  }
}

// ------------------------------------------------------------------ VAD
constexpr int kVadThreads = 256;

__global__ void __launch_bounds__(kVadThreads) estoi_vad_kernel(EstoiGeom g, EstoiBuffers b) {
  const int pair = blockIdx.x, tid = threadIdx.x, lane = tid & 31, wib = tid >> 5;
  constexpr int NW = kVadThreads / 32;
  const int nfa = g.nfa[pair];
  const float* __restrict__ x = b.x10 + g.off10[pair];
  double* __restrict__ en = b.energy + g.offfr[pair];
  int32_t* __restrict__ kept = b.kept + g.offfr[pair];
  __shared__ double red[32];
  __shared__ int s_cnt[NW];
  __shared__ int s_base;
  __shared__ float s_win[kStoiFrame];
  if (nfa <= 0) {
    if (tid == 0) b.nkept[pair] = 0;
    return;
  }
  s_win[tid] = g_stoi_win[tid];
  __syncthreads();
  double mx = -1.0e300;
  for (int f = wib; f < nfa; f += NW) {
    const float* fr = x + (int64_t)f * kStoiHop;
    double ss = 0.0;
#pragma unroll
    for (int k = 0; k < kStoiFrame / 32; ++k) {
      const int i = k * 32 + lane;
      const double v = (double)s_win[i] * (double)fr[i];
      ss += v * v;
    }
    ss = warp_sum(ss);
    const double e = 20.0 * log10(sqrt(ss) + 2.220446049250313e-16);
    if (lane == 0) en[f] = e;
    mx = fmax(mx, e);
  }
  mx = block_max(mx, red);
  const double thr = mx - 40.0;
  if (tid == 0) s_base = 0;
  __syncthreads();
  for (int f0 = 0; f0 < nfa; f0 += kVadThreads) {
    const int f = f0 + tid;
    const int keep = (f < nfa && en[f] > thr) ? 1 : 0;
    int inc = keep;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int t = __shfl_up_sync(0xffffffffu, inc, o);
      if (lane >= o) inc += t;
    }
    if (lane == 31) s_cnt[wib] = inc;
    __syncthreads();
    int woff = 0, tot = 0;
    for (int w = 0; w < NW; ++w) {
      if (w < wib) woff += s_cnt[w];
      tot += s_cnt[w];
    }
    const int base = s_base;
    if (keep) kept[base + woff + inc - 1] = f;
    __syncthreads();
    if (tid == 0) s_base = base + tot;
    __syncthreads();
  }
  if (tid == 0) b.nkept[pair] = s_base;
}

// ------------------------------------------------- STFT + one-third octaves
constexpr int kTobWarps = 8;

__device__ __forceinline__ float2 cmul(float2 a, float2 w) {
  return make_float2(a.x * w.x - a.y * w.y, a.x * w.y + a.y * w.x);
}

__global__ void __launch_bounds__(kTobWarps * 32) estoi_tob_kernel(EstoiGeom g, EstoiBuffers b) {
  const int pair = blockIdx.y, lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const int nk = b.nkept[pair];
  const int nfr = nk - 1;  // STFT frames of the silence-removed signal
  __shared__ float2 s_z[kTobWarps][kStoiFft];
  __shared__ float2 s_tw[kStoiFft / 2];
  __shared__ float s_win[kStoiFrame];
  for (int k = threadIdx.x; k < kStoiFft / 2; k += kTobWarps * 32) {
    s_tw[k] = g_stoi_tw[k];
    s_win[k] = g_stoi_win[k];
  }
  __syncthreads();
  const int m = blockIdx.x * kTobWarps + wib;
  if (m >= nfr) return;
  const float* __restrict__ x = b.x10 + g.off10[pair];
  const float* __restrict__ y = b.x10 + b.tot10 + g.off10[pair];
  const int32_t* __restrict__ kept = b.kept + g.offfr[pair];
  float2* z = s_z[wib];
  const int64_t sc = (int64_t)kept[m] * kStoiHop;
  const int64_t sp = (m > 0) ? (int64_t)kept[m - 1] * kStoiHop : -1;
  const int64_t sn = (int64_t)kept[m + 1] * kStoiHop;
  // frame of the overlap-added signal, windowed again, bit-reversed into z
#pragma unroll
  for (int k = 0; k < kStoiFrame / 32; ++k) {
    const int i = k * 32 + lane;
    const float w = s_win[i];
    float vx = w * x[sc + i], vy = w * y[sc + i];
    if (i < kStoiHop) {
      if (sp >= 0) {
        const float w2 = s_win[i + kStoiHop];
        vx = fmaf(w2, x[sp + i + kStoiHop], vx);
        vy = fmaf(w2, y[sp + i + kStoiHop], vy);
      }
    } else {
      const float w2 = s_win[i - kStoiHop];
      vx = fmaf(w2, x[sn + i - kStoiHop], vx);
      vy = fmaf(w2, y[sn + i - kStoiHop], vy);
    }
    const int br = (int)(__brev((unsigned)i) >> 23);  // 9-bit reversal
    z[br] = make_float2(w * vx, w * vy);
    z[br | 1] = make_float2(0.f, 0.f);                // i + 256 reverses to br + 1
  }
  __syncwarp();
  // in-place radix-2 decimation-in-time, 9 stages, 8 butterflies per lane per stage
#pragma unroll
  for (int s = 0; s < 9; ++s) {
    const int half = 1 << s;
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const int j = u * 32 + lane;
      const int pos = j & (half - 1);
      const int i0 = ((j >> s) << (s + 1)) + pos, i1 = i0 + half;
      const float2 t = cmul(z[i1], s_tw[pos << (8 - s)]);
      const float2 a = z[i0];
      z[i0] = make_float2(a.x + t.x, a.y + t.y);
      z[i1] = make_float2(a.x - t.x, a.y - t.y);
    }
    __syncwarp();
  }
  // spectra of the two real signals from Z = FFT(x + i y)
  float px[8], py[8];
#pragma unroll
  for (int u = 0; u < 8; ++u) {
    const int k = u * 32 + lane;
    const float2 a = z[k], c = z[(kStoiFft - k) & (kStoiFft - 1)];
    const float xr = a.x + c.x, xi = a.y - c.y;   // 2 X[k]
    const float yr = a.y + c.y, yi = c.x - a.x;   // 2 Y[k]
    px[u] = 0.25f * (xr * xr + xi * xi);
    py[u] = 0.25f * (yr * yr + yi * yi);
  }
  __syncwarp();
  float* pw = reinterpret_cast<float*>(z);
#pragma unroll
  for (int u = 0; u < 8; ++u) {
    pw[u * 32 + lane] = px[u];
    pw[256 + u * 32 + lane] = py[u];
  }
  __syncwarp();
  if ((lane & 15) < kStoiBands) {
    const int band = lane & 15, sig = lane >> 4;
    const float* p = pw + sig * 256;
    float acc = 0.f;
    for (int k = c_stoi_lo[band]; k < c_stoi_hi[band]; ++k) acc += p[k];
    b.tob[((int64_t)sig * b.totfr + g.offfr[pair] + m) * kStoiBands + band] = sqrtf(acc);
  }
}

// ------------------------------------------------------ segment correlation
constexpr int kCorrThreads = 128;

// classic STOI (pystoi.stoi(..., extended=False)): per segment and band, y is scaled to the
// energy of x, clipped at x (1 + 10^(15/20)) and correlated with x over the 30 frames
__device__ __forceinline__ float stoi_classic_segment(const float* __restrict__ xs, const float* __restrict__ ys) {
  const float kEpsF = 2.220446049250313e-16f, kClip = 1.0f + 5.623413251903491f;  // BETA = -15 dB
  float acc = 0.f;
#pragma unroll 1
  for (int k = 0; k < kStoiBands; ++k) {
    float xv[kStoiSeg], yv[kStoiSeg];
    float ex = 0.f, ey = 0.f;
#pragma unroll
    for (int t = 0; t < kStoiSeg; ++t) {
      xv[t] = __ldg(xs + t * kStoiBands + k);
      yv[t] = __ldg(ys + t * kStoiBands + k);
      ex = fmaf(xv[t], xv[t], ex);
      ey = fmaf(yv[t], yv[t], ey);
    }
    const float nrm = sqrtf(ex / (ey + kEpsF));
    float mx = 0.f, my = 0.f;
#pragma unroll
    for (int t = 0; t < kStoiSeg; ++t) {
      yv[t] = fminf(yv[t] * nrm, xv[t] * kClip);
      mx += xv[t];
      my += yv[t];
    }
    mx *= (1.0f / kStoiSeg);
    my *= (1.0f / kStoiSeg);
    float qx = 0.f, qy = 0.f, xy = 0.f;
#pragma unroll
    for (int t = 0; t < kStoiSeg; ++t) {
      const float dx = xv[t] - mx, dy = yv[t] - my;
      qx = fmaf(dx, dx, qx);
      qy = fmaf(dy, dy, qy);
      xy = fmaf(dx, dy, xy);
    }
    acc += xy / ((sqrtf(qx) + kEpsF) * (sqrtf(qy) + kEpsF));
  }
  return acc * (1.0f / kStoiBands);
}

__global__ void __launch_bounds__(kCorrThreads) estoi_corr_kernel(EstoiGeom g, EstoiBuffers b, int classic) {
  const int pair = blockIdx.x, tid = threadIdx.x;
  const int nfr = b.nkept[pair] - 1;
  __shared__ double red[32];
  if (nfr < kStoiSeg) {  // pystoi: RuntimeWarning + sentinel
    if (tid == 0) {
      b.score[pair] = 1.0e-5;
      b.status[pair] = 2;
    }
    return;
  }
  const int J = nfr - kStoiSeg + 1;
  const float* __restrict__ X = b.tob + (g.offfr[pair]) * kStoiBands;
  const float* __restrict__ Y = b.tob + (b.totfr + g.offfr[pair]) * kStoiBands;
  double total = 0.0;
  for (int m = tid; m < J; m += kCorrThreads) {
    const float* xs = X + (int64_t)m * kStoiBands;
    const float* ys = Y + (int64_t)m * kStoiBands;
    if (classic) {
      total += (double)stoi_classic_segment(xs, ys);
      continue;
    }
    float mux[kStoiBands], ivx[kStoiBands], muy[kStoiBands], ivy[kStoiBands];
#pragma unroll
    for (int k = 0; k < kStoiBands; ++k) {
      float sx = 0.f, sy = 0.f;
      for (int t = 0; t < kStoiSeg; ++t) {
        sx += __ldg(xs + t * kStoiBands + k);
        sy += __ldg(ys + t * kStoiBands + k);
      }
      sx *= (1.0f / kStoiSeg);
      sy *= (1.0f / kStoiSeg);
      float qx = 0.f, qy = 0.f;
      for (int t = 0; t < kStoiSeg; ++t) {
        const float dx = __ldg(xs + t * kStoiBands + k) - sx, dy = __ldg(ys + t * kStoiBands + k) - sy;
        qx = fmaf(dx, dx, qx);
        qy = fmaf(dy, dy, qy);
      }
      mux[k] = sx;
      muy[k] = sy;
      ivx[k] = 1.0f / sqrtf(qx);
      ivy[k] = 1.0f / sqrtf(qy);
    }
    float acc = 0.f;
    for (int t = 0; t < kStoiSeg; ++t) {
      float vx[kStoiBands], vy[kStoiBands];
      float cx = 0.f, cy = 0.f;
#pragma unroll
      for (int k = 0; k < kStoiBands; ++k) {
        vx[k] = (__ldg(xs + t * kStoiBands + k) - mux[k]) * ivx[k];
        vy[k] = (__ldg(ys + t * kStoiBands + k) - muy[k]) * ivy[k];
        cx += vx[k];
        cy += vy[k];
      }
      cx *= (1.0f / kStoiBands);
      cy *= (1.0f / kStoiBands);
      float nx = 0.f, ny = 0.f, xy = 0.f;
#pragma unroll
      for (int k = 0; k < kStoiBands; ++k) {
        const float dx = vx[k] - cx, dy = vy[k] - cy;
        nx = fmaf(dx, dx, nx);
        ny = fmaf(dy, dy, ny);
        xy = fmaf(dx, dy, xy);
      }
      acc += xy / sqrtf(nx * ny);
    }
    total += (double)acc * (1.0 / kStoiSeg);
  }
  total = block_sum(total, red);
  if (tid == 0) {
    b.score[pair] = total / (double)J;
    b.status[pair] = 0;
  }
}

// ------------------------------------------------------------- launchers
void estoi_upload_tables(const float* win, const int* lo, const int* hi, const float* tw /*[256][2]*/, cudaStream_t s) {
  cudaMemcpyToSymbolAsync(g_stoi_win, win, sizeof(float) * kStoiFrame, 0, cudaMemcpyHostToDevice, s);
  cudaMemcpyToSymbolAsync(c_stoi_lo, lo, sizeof(int) * kStoiBands, 0, cudaMemcpyHostToDevice, s);
  cudaMemcpyToSymbolAsync(c_stoi_hi, hi, sizeof(int) * kStoiBands, 0, cudaMemcpyHostToDevice, s);
  cudaMemcpyToSymbolAsync(g_stoi_tw, tw, sizeof(float) * kStoiFft, 0, cudaMemcpyHostToDevice, s);
  cudaStreamSynchronize(s);
}

// taps [up][2K + 1] of the polyphase resampler (host_tables.hpp make_estoi_polytaps): the 16 -> 10 kHz
// fast path keeps them zero padded in constant memory
static bool g_rf_taps_ok = false;
void estoi_upload_polytaps(const double* taps, int up, int K, cudaStream_t s) {
  g_rf_taps_ok = false;
  const int nt = 2 * K + 1;
  if (up != 5 || nt > kRfNt) return;
  static double padded[5][kRfNt];
  for (int r = 0; r < 5; ++r)
    for (int k = 0; k < kRfNt; ++k) padded[r][k] = k < nt ? taps[r * nt + k] : 0.0;
  cudaMemcpyToSymbolAsync(c_rf_tap, padded, sizeof(padded), 0, cudaMemcpyHostToDevice, s);
  static float paddedf[5][kRfNt];
  for (int r = 0; r < 5; ++r)
    for (int k = 0; k < kRfNt; ++k) paddedf[r][k] = (float)padded[r][k];
  cudaMemcpyToSymbolAsync(c_rf_tapf, paddedf, sizeof(paddedf), 0, cudaMemcpyHostToDevice, s);
  cudaStreamSynchronize(s);
  g_rf_taps_ok = true;
}

int estoi_run(const EstoiGeom& g, const EstoiBuffers& b, int n, int max_n10, int max_nfa, bool classic, KernelTimer* kt,
              cudaStream_t s) {
  int launches = 0;
  const int groups = std::max(1, kRsThreads / b.up);       // (an `up` above 256 never occurs for audio rates)
  const int tile = kRsJ * b.up * groups;
  const int span = (int)(((int64_t)tile * b.down) / b.up) + 2 * b.K + 4;
  const size_t tap_bytes = (size_t)b.up * (2 * b.K + 1) * sizeof(double);
  const int taps_in_smem = tap_bytes <= 64 * 1024;
  const size_t smem = (size_t)span * sizeof(float) + (taps_in_smem ? tap_bytes : 0);
  if (smem > 48 * 1024)
    cudaFuncSetAttribute(estoi_resample_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  static const bool generic_only = getenv("NELE_ESTOI_RESAMPLE_GENERIC") != nullptr;   // A/B switch
  kt_begin(kt, "estoi_resample", s);
  if (b.up == 5 && b.down == 8 && g_rf_taps_ok && !generic_only)
  {
    static const bool f32 = [] { const char* p = getenv("NELE_RESAMPLE_F32"); return !(p && p[0] == '0'); }();
    const dim3 grid((max_n10 + kRfTilesPerCta * kRfTile - 1) / (kRfTilesPerCta * kRfTile), n, 2);
    if (f32) estoi_resample58_kernel<float><<<grid, kRfThreads, 0, s>>>(g, b);
    else estoi_resample58_kernel<double><<<grid, kRfThreads, 0, s>>>(g, b);
  }
  else
    estoi_resample_kernel<<<dim3((max_n10 + tile - 1) / tile, n, 2), kRsThreads, smem, s>>>(g, b, span, groups, taps_in_smem);
  kt_end(kt, s);
  ++launches;
  kt_begin(kt, "estoi_vad", s);
  estoi_vad_kernel<<<n, kVadThreads, 0, s>>>(g, b);
  kt_end(kt, s);
  ++launches;
  if (max_nfa > 1) {
    kt_begin(kt, "estoi_tob", s);
    estoi_tob_kernel<<<dim3((max_nfa - 1 + kTobWarps - 1) / kTobWarps, n), kTobWarps * 32, 0, s>>>(g, b);
    kt_end(kt, s);
    ++launches;
  }
  kt_begin(kt, "estoi_corr", s);
  estoi_corr_kernel<<<n, kCorrThreads, 0, s>>>(g, b, classic ? 1 : 0);
  kt_end(kt, s);
  ++launches;
  return launches;
}

}  // namespace nele
