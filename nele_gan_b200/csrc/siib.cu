// SIIB^Gauss on sm_100a: batched restatement of intel.py:57-100 (wrapper VAD + tiling to
// >= 25 s) and pysiib.SIIB(x, y, fs, gauss=True) (algorithm: oracle/pysiib_np.py).
//
//   siib_wrapvad_kernel  per pair CTA: the wrapper's VAD on the untiled clean signal
//                        (intel.py:62-75) -> tiling factor M.  The tiled signal is never
//                        materialised: frame f of it is x[(200 f + i) mod L].
//   siib_vad_kernel      per pair CTA: means, frame powers of the tiled mean-removed x,
//                        99.9th percentile by k-fold arg-max, ordered compaction of the
//                        active frames (intel.py:37-50 semantics inside pysiib)
//   siib_spec_kernel     per (pair, active frame) warp: Hann frame of x + i y, 400-point
//                        FFT (fft400.cuh), power spectra, 28 gammatone band energies, log
//                        -> logspec [t][32] (time-major, 28 bands + 4 zero lanes)
//   siib_mask_kernel     per (pair, signal) warp, lane = band: forward temporal masking
//                        (16 frames, sequential in time) and mean removal
//   siib_cov_kernel      per (pair, lag block) warp: FP64 28x28 lag products
//                        D_d = sum_t a_{t+max(0,-d)} b_{t+max(0,d)}^T for xx, yy (d = 0..14)
//                        and xy (d = -14..14).  The 420 x 420 stacked covariances are
//                        block-Toeplitz up to 14 edge terms per entry, so 59 blocks replace
//                        three 420 x 420 x Nf products (8x fewer flops, exact).
//   siib_expand_kernel   per pair CTA: walks every block diagonal with the one-term-out,
//                        one-term-in recurrence -> centred scatter matrices Sxx (FP64),
//                        Sxy, Syy (FP32, folded onto the lower triangle: only u^T S u is ever needed)
//   siib_chol_kernel     per pair CTA: diagonally pivoted left-looking Cholesky of Sxx in
//                        FP64, stops at the numerical rank r -> L (FP32 copy, r columns)
//   siib_jacobi_kernel   per pair CTA: one-sided (Hestenes) Jacobi on the r columns of L;
//                        at convergence column j = sqrt(lambda_j) u_j, i.e. the KLT basis of
//                        cov(X) without accumulating rotations.  Column blocks of 4 are held
//                        in registers and paired by a round-robin tournament.
//   qf::quadform_kernel  (siib_klt.cu) rho_j = u^T Sxy u / sqrt(lambda_j u^T Syy u), the
//                        Gaussian channel capacity sum and the final score
//
// Rank-deficient inputs (a tiled signal whose length is a multiple of the 200-sample hop
// repeats its frames exactly) stop the Cholesky early; the null space contributes zero
// information, which is the exact-arithmetic value of the reference formula.
#include <cooperative_groups.h>

#include <algorithm>

#include "fft400.cuh"
#include <stdlib.h>

#include "kernels.h"

namespace cg = cooperative_groups;

namespace nele {

constexpr int kSWin = 400, kSHop = 200, kSBins = 201, kSBands = 28, kSLanes = 32, kSStack = 15;
constexpr int kSDim = kSBands * kSStack;  // 420
constexpr int kSLd = 448;                 // padded column length of L (14 x 32)
constexpr int kSBlocks = 59;              // xx d=0..14, yy d=0..14, xy d=-14..14
constexpr int kSMaskT = 16;               // floor(0.2 s * 80 frames/s)
constexpr double kEps = 2.220446049250313e-16;

__device__ float g_siib_win[kSWin];       // periodic Hann(400); staged in shared memory by its users
                                          // (lane-indexed reads of __constant__ data serialise)
__constant__ float c_siib_decay[kSMaskT];  // log(d + 1) / log(16)
__device__ float g_siib_g2t[kSBins * kSLanes];  // squared gammatone responses, [bin][band]
__device__ cpx g_siib_tw[kSWin];

// ---------------------------------------------------------------- VAD helpers
// energy of Hann frame f of (x - mean), summed over the warp; wrap = tiled signal, else zero padded.  win = the window
// as doubles in shared memory.  A frame that neither wraps nor runs past the end (all but one per tile) takes the loop
// without per-sample index checks; the dB conversion is left to the caller, one frame per thread (frame_db) instead of
// one per warp: the VAD kernels are issue bound and log10 in FP64 was a fifth of their instructions.
__device__ __forceinline__ double frame_energy(const float* __restrict__ x, int L, double mean, int64_t f, bool wrap,
                                               int lane, const double* __restrict__ win) {
  // f * 200 < 2^31 for every frame count the engine admits (F <= 400 000)
  const int s0 = (int)f * kSHop;
  const int base = wrap ? (s0 % L) : s0;
  double ss = 0.0;
  if (base + kSWin <= L) {
    const float* __restrict__ xf = x + base;
#pragma unroll
    for (int k = 0; k < (kSWin + 31) / 32; ++k) {
      const int i = k * 32 + lane;
      if (i < kSWin) {
        const double v = ((double)xf[i] - mean) * win[i];
        ss = fma(v, v, ss);
      }
    }
  } else {
#pragma unroll 1
    for (int i = lane; i < kSWin; i += 32) {
      int idx = base + i;
      double v;
      if (wrap) {
        if (idx >= L) idx = (L >= kSWin) ? idx - L : idx % L;
        v = (double)x[idx] - mean;
      } else {
        v = (idx < L) ? (double)x[idx] - mean : 0.0;
      }
      v *= win[i];
      ss = fma(v, v, ss);
    }
  }
  return warp_sum(ss);
}
__device__ __forceinline__ double frame_db(double ss) { return 10.0 * log10(ss / (double)kSWin + kEps); }

// k-th largest (counting multiplicity) of v[0..n), k >= 1: k rounds of arg-max in the
// total order (value descending, index ascending).  All threads get the result.
__device__ double kth_largest(const double* __restrict__ v, int64_t n, int k, double* red, int64_t* redi) {
  double pv = 1.0e300;
  int64_t pi = -1;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = blockDim.x >> 5;
  for (int r = 0; r < k; ++r) {
    double bv = -1.0e300;
    int64_t bi = -1;
    for (int64_t i = threadIdx.x; i < n; i += blockDim.x) {
      const double c = v[i];
      const bool after = (c < pv) || (c == pv && i > pi);
      if (after && (c > bv || (c == bv && i < bi) || bi < 0)) {
        bv = c;
        bi = i;
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const double ov = __shfl_xor_sync(0xffffffffu, bv, o);
      const int64_t oi = __shfl_xor_sync(0xffffffffu, bi, o);
      if (oi >= 0 && (bi < 0 || ov > bv || (ov == bv && oi < bi))) {
        bv = ov;
        bi = oi;
      }
    }
    __syncthreads();
    if (lane == 0) {
      red[w] = bv;
      redi[w] = bi;
    }
    __syncthreads();
    bv = red[0];
    bi = redi[0];
    for (int j = 1; j < nw; ++j) {
      const double ov = red[j];
      const int64_t oi = redi[j];
      if (oi >= 0 && (bi < 0 || ov > bv || (ov == bv && oi < bi))) {
        bv = ov;
        bi = oi;
      }
    }
    pv = bv;
    pi = bi;
  }
  return pv;
}

__device__ __forceinline__ int percentile_rank(int64_t n) {  // k such that the k-th largest is sorted[ind]
  int64_t ind = (int64_t)rint((double)n * 0.999) - 1;       // int(round(len * 0.999) - 1), round-half-even
  if (ind < 0) ind += n;                                     // numpy negative index
  if (ind < 0) ind = 0;
  if (ind > n - 1) ind = n - 1;
  return (int)(n - ind);
}

constexpr int kVadThreads = 256;
// the tiled-signal VAD re-reads its waveform M times: 1024-thread CTAs keep the number of
// resident waveforms (2 per SM x 192 KB) inside the L2
constexpr int kVad2Threads = 1024;

__global__ void __launch_bounds__(kVadThreads) siib_wrapvad_kernel(SiibGeom g, SiibBuffers b, int no_tile) {
  const int pair = blockIdx.x, tid = threadIdx.x, lane = tid & 31, wib = tid >> 5;
  constexpr int NW = kVadThreads / 32;
  const float* __restrict__ x = b.ref + g.off16[pair];
  const int L = g.len16[pair];
  const int Lp = max(L, kSWin + 1);
  const int F1 = (Lp - kSWin + kSHop - 1) / kSHop;
  double* __restrict__ db = b.wrapdb + g.offW[pair];
  __shared__ double red[32];
  __shared__ int64_t redi[32];
  __shared__ double s_win[kSWin];
  for (int i = tid; i < kSWin; i += kVadThreads) s_win[i] = (double)g_siib_win[i];
  __syncthreads();
  for (int f = wib; f < F1; f += NW) {
    const double d = frame_energy(x, L, 0.0, f, false, lane, s_win);
    if (lane == 0) db[f] = d;
  }
  __syncthreads();
  for (int f = tid; f < F1; f += kVadThreads) db[f] = frame_db(db[f]);
  __syncthreads();
  const double sel = kth_largest(db, F1, percentile_rank(F1), red, redi);
  const double thr = sel - 40.0;
  int cnt = 0;
  for (int f = tid; f < F1; f += kVadThreads) cnt += (db[f] > thr) ? 1 : 0;
  const int active = (int)block_sum((double)cnt, red);
  if (tid == 0) {
    int M = 1;
    if (!no_tile && (double)active / 80.0 < 20.0) M = (active > 0) ? (int)floor(25.0 / ((double)active / 80.0)) : 0;
    b.M[pair] = M;
    b.wrap_active[pair] = active;
  }
}

__global__ void __launch_bounds__(kVad2Threads) siib_vad_kernel(SiibGeom g, SiibBuffers b) {
  const int pair = b.pair_lo + blockIdx.x, tid = threadIdx.x, lane = tid & 31, wib = tid >> 5;
  constexpr int NW = kVad2Threads / 32;
  const float* __restrict__ x = b.ref + g.off16[pair];
  const float* __restrict__ y = b.deg + g.off16[pair];
  const int L = g.len16[pair];
  const int64_t F = g.F[pair];
  double* __restrict__ db = b.xdb + g.offF[pair];
  int32_t* __restrict__ act = b.act + g.offF[pair];
  int32_t* __restrict__ aidx = b.aidx + g.offF[pair];
  int32_t* __restrict__ src = b.src + g.offF[pair];
  __shared__ double red[32];
  __shared__ int64_t redi[32];
  __shared__ int s_cnt[NW];
  __shared__ int s_base;
  __shared__ double s_win[kSWin];
  if (F <= 0) {
    if (tid == 0) b.Fa[pair] = 0;
    return;
  }
  for (int i = tid; i < kSWin; i += kVad2Threads) s_win[i] = (double)g_siib_win[i];
  double sx = 0.0, sy = 0.0;
  for (int i = tid; i < L; i += kVad2Threads) {
    sx += (double)x[i];
    sy += (double)y[i];
  }
  sx = block_sum(sx, red);
  sy = block_sum(sy, red);
  const double mx = sx / (double)L, my = sy / (double)L;
  if (tid == 0) {
    b.mean[2 * pair] = mx;
    b.mean[2 * pair + 1] = my;
  }
  // Frame f of the tiled signal starts at sample (200 f) mod L: frames one period apart
  // (per = L / gcd(L, 200) frames) are identical, so only the first period is analysed -- here and
  // in the spectrum kernel -- and the rest copied (a 3.0 s utterance tiled 8 times has 240
  // distinct frames out of 1920).
  int gcd = L, r200 = kSHop;
  while (r200) {
    const int t = gcd % r200;
    gcd = r200;
    r200 = t;
  }
  const int64_t per = L / gcd;
  const int64_t Fu = (per < F) ? per : F;
  for (int64_t f = wib; f < Fu; f += NW) {
    const double d = frame_energy(x, L, mx, f, true, lane, s_win);
    if (lane == 0) db[f] = d;
  }
  __syncthreads();
  for (int64_t f = tid; f < Fu; f += kVad2Threads) db[f] = frame_db(db[f]);
  __syncthreads();
  for (int64_t f = Fu + tid; f < F; f += kVad2Threads) db[f] = db[f % per];
  __syncthreads();
  const double sel = kth_largest(db, F, percentile_rank(F), red, redi);
  const double thr = sel - 40.0;
  if (tid == 0) s_base = 0;
  __syncthreads();
  for (int64_t f0 = 0; f0 < F; f0 += kVad2Threads) {
    const int64_t f = f0 + tid;
    const int keep = (f < F && db[f] > thr) ? 1 : 0;
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
    if (keep) {
      act[base + woff + inc - 1] = (int32_t)f;
      if (f < Fu) aidx[f] = base + woff + inc - 1;
    }
    __syncthreads();
    if (tid == 0) s_base = base + tot;
    __syncthreads();
  }
  const int Fa = s_base;
  if (tid == 0) b.Fa[pair] = Fa;
  // row of the raw spectra that active frame t reads: its own, or that of its first occurrence
  for (int t = tid; t < Fa; t += kVad2Threads) {
    const int64_t f = act[t];
    src[t] = (f < Fu) ? t : aidx[f % per];
  }
  if (tid == 0) {  // active frames per period of the tiled signal (0: the signal does not repeat within F frames)
    int P = 0;
    if (per < F) {
      int lo = 0, hi = Fa;
      while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (act[mid] < per) lo = mid + 1;
        else hi = mid;
      }
      P = lo;
    }
    b.Pact[pair] = P;
  }
}

// ------------------------------------------------------------- spectra
constexpr int kSpecWarps = 8;
struct SpecSmem {
  cpx twa[kSWin];            // fft400.cuh: phase-A twiddles [k1][n2]
  cpx tw25[32];              // phase-B twiddles (25 used)
  float win[kSWin];
  float g2t[kSBins * kSLanes];
  cpx z[kSpecWarps][kSWin];  // one buffer per warp: the FFT is in place (fft400.cuh)
  float2 pw[kSpecWarps][kSBins + 1];  // power spectra (px, py) in bin order: two bins per 128-bit broadcast read
};

constexpr int kSpecIter = 8;  // frames per warp: the 29 KB of tables are staged once per 64 frames, not once per 8

// 68 KB of shared memory and <= 80 registers: three CTAs (24 warps) per SM.  Round 2 started from 2970 instructions
// per frame at 68 % of the issue slots with 16 warps; unrolled band sums, constant FFT twiddles and constant offsets
// brought that to 1985, at which point the shared-memory pipe became the limit (ncu: 93 % busy) -- hence the
// conflict-free twiddle table and the 128-bit reads of the power spectra.
__global__ void __launch_bounds__(kSpecWarps * 32, 3) siib_spec_kernel(SiibGeom g, SiibBuffers b) {
  const int pair = b.pair_lo + blockIdx.y, lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const int Fa = b.Fa[pair];
  const int tbase = blockIdx.x * kSpecWarps * kSpecIter;
  if (tbase >= Fa) return;
  {
    // nothing to do when every frame of this CTA is a copy of an earlier one (siib_vad_kernel):
    // leave before the 29 KB of tables are staged
    int need = 0;
    for (int it = 0; it < kSpecIter; ++it) {
      const int tw = tbase + it * kSpecWarps + wib;
      need |= (tw < Fa) && (b.src[g.offF[pair] + tw] == tw);
    }
    if (!__syncthreads_or(need)) return;
  }
  extern __shared__ __align__(16) unsigned char smem_raw[];
  SpecSmem& sm = *reinterpret_cast<SpecSmem*>(smem_raw);
  for (int k = threadIdx.x; k < kSWin; k += kSpecWarps * 32) {
    sm.twa[k] = g_siib_tw[(k % 25) * (k / 25)];   // n2 k1 <= 360
    sm.win[k] = g_siib_win[k];
  }
  if (threadIdx.x < 25) sm.tw25[threadIdx.x] = g_siib_tw[16 * threadIdx.x];
  for (int k = threadIdx.x; k < kSBins * kSLanes; k += kSpecWarps * 32) sm.g2t[k] = g_siib_g2t[k];
  __syncthreads();
  const float* __restrict__ x = b.ref + g.off16[pair];
  const float* __restrict__ y = b.deg + g.off16[pair];
  const int L = g.len16[pair];
  const float mx = (float)b.mean[2 * pair], my = (float)b.mean[2 * pair + 1];
  cpx* z = sm.z[wib];
  const float* gl = sm.g2t + lane;
  for (int it = 0; it < kSpecIter; ++it) {
    const int t = tbase + it * kSpecWarps + wib;
    if (t >= Fa) break;
    if (b.src[g.offF[pair] + t] != t) continue;  // a copy of an earlier frame (siib_vad_kernel)
    const int64_t f = b.act[g.offF[pair] + t];
    const int base = (int)((f * kSHop) % L);
    __syncwarp();
    if (base + kSWin <= L) {
      // all but one frame per tile: no wrap, constant offsets from one pointer per signal
      const float* __restrict__ xf = x + base + lane;
      const float* __restrict__ yf = y + base + lane;
      const float* wl = sm.win + lane;
      cpx* zl = z + lane;
#pragma unroll
      for (int k = 0; k < (kSWin + 31) / 32; ++k)
        if (k * 32 + 32 <= kSWin || lane < kSWin - k * 32) {
          const float w = wl[k * 32];
          zl[k * 32] = {w * (xf[k * 32] - mx), w * (yf[k * 32] - my)};
        }
    } else {
#pragma unroll 1
      for (int i = lane; i < kSWin; i += 32) {
        int idx = base + i;  // < 2 L
        if (idx >= L) {
          idx -= L;
          if (idx >= L) idx %= L;  // utterances shorter than a frame
        }
        const float w = sm.win[i];
        z[i] = {w * (x[idx] - mx), w * (y[idx] - my)};
      }
    }
    __syncwarp();
    if (lane < 25) fft400_phase_a(lane, z, sm.twa);
    __syncwarp();
    if (lane < 16) fft400_phase_b(lane, z, sm.tw25);
    __syncwarp();
    // power spectra of the two real signals.  k = lane + 32 it: fft400_pos(k) = pos(lane) + 2 it and
    // fft400_pos(400 - k) = pos(400 - lane) - 2 it (k > 0), so the unrolled loop reads both with constant offsets
    float2* pw = sm.pw[wib];
    {
      const cpx* za = z + fft400_pos(lane);
      const cpx* zc = z + fft400_pos(kSWin - lane);  // lane 0: position of "bin 400", replaced by bin 0 below
#pragma unroll
      for (int it = 0; it < (kSBins + 31) / 32; ++it)
        if (it * 32 + 32 <= kSBins || lane < kSBins - it * 32) {
          const cpx a = za[2 * it];
          const cpx c = (it == 0 && lane == 0) ? a : zc[-2 * it];
          const float xr = a.x + c.x, xi = a.y - c.y, yr = a.y + c.y, yi = c.x - a.x;
          pw[it * 32 + lane] = make_float2(0.25f * (xr * xr + xi * xi), 0.25f * (yr * yr + yi * yi));
        }
    }
    __syncwarp();
    // band energies: the kernel is bound by the shared-memory pipe (93 % busy), and these sums are most of its
    // wavefronts -- one per lane-indexed filter weight, one per TWO bins of the broadcast power spectra
    float ex = 0.f, ey = 0.f;
    const float4* pw4 = reinterpret_cast<const float4*>(pw);
#pragma unroll
    for (int k2 = 0; k2 < kSBins / 2; ++k2) {
      const float4 p = pw4[k2];
      const float g0 = gl[(2 * k2) * kSLanes], g1 = gl[(2 * k2 + 1) * kSLanes];
      ex = fmaf(g0, p.x, ex);
      ey = fmaf(g0, p.y, ey);
      ex = fmaf(g1, p.z, ex);
      ey = fmaf(g1, p.w, ey);
    }
    {
      const float2 p = pw[kSBins - 1];
      const float gk = gl[(kSBins - 1) * kSLanes];
      ex = fmaf(gk, p.x, ex);
      ey = fmaf(gk, p.y, ey);
    }
    const int64_t row = g.offF[pair] + t;
    b.lograw[row * kSLanes + lane] = (lane < kSBands) ? logf(ex + (float)kEps) : 0.f;
    b.lograw[(b.totF + row) * kSLanes + lane] = (lane < kSBands) ? logf(ey + (float)kEps) : 0.f;
  }
}

// --------------------------------------------------- forward masking + de-mean
__global__ void __launch_bounds__(128) siib_mask_kernel(SiibGeom g, SiibBuffers b, int n_items) {
  const int lane = threadIdx.x & 31;
  const int item = blockIdx.x * 4 + (threadIdx.x >> 5);
  if (item >= n_items) return;
  const int pair = b.pair_lo + (item >> 1), q = item & 1;
  const int Fa = b.Fa[pair];
  float* __restrict__ X = b.logspec + ((int64_t)q * b.totF + g.offF[pair]) * kSLanes + lane;
  const float* __restrict__ R = b.lograw + ((int64_t)q * b.totF + g.offF[pair]) * kSLanes + lane;
  const int32_t* __restrict__ src = b.src + g.offF[pair];
  // Every pass below takes the frames eight at a time with the loads issued first: the row of a frame is a
  // two-level gather (first-occurrence map, then the row), and one frame per iteration leaves the warp
  // waiting on that latency for most of its life.
  float fl = 3.0e38f;
  for (int t0 = 0; t0 < Fa; t0 += 8) {
    float raw[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) raw[u] = R[(int64_t)src[min(t0 + u, Fa - 1)] * kSLanes];
#pragma unroll
    for (int u = 0; u < 8; ++u) fl = fminf(fl, raw[u]);
  }
  float hx[kSMaskT - 1], he[kSMaskT - 1];  // masked level and (level - floor) of the previous 15 frames
#pragma unroll
  for (int d = 0; d < kSMaskT - 1; ++d) {
    hx[d] = -3.0e38f;
    he[d] = 0.f;
  }
  double sum = 0.0;
  for (int t0 = 0; t0 < Fa; t0 += 8) {
    float raw[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) raw[u] = R[(int64_t)src[min(t0 + u, Fa - 1)] * kSLanes];
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const int t = t0 + u;
      if (t < Fa) {
        float v = raw[u];
#pragma unroll
        for (int d = 0; d < kSMaskT - 1; ++d) v = fmaxf(v, fmaf(-c_siib_decay[d + 1], he[d], hx[d]));
#pragma unroll
        for (int d = kSMaskT - 2; d > 0; --d) {
          hx[d] = hx[d - 1];
          he[d] = he[d - 1];
        }
        hx[0] = v;
        he[0] = v - fl;
        X[(int64_t)t * kSLanes] = v;
        sum += (double)v;
      }
    }
  }
  const float mu = (Fa > 0) ? (float)(sum / (double)Fa) : 0.f;
  for (int t0 = 0; t0 < Fa; t0 += 8) {
    float cur[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) cur[u] = X[(int64_t)min(t0 + u, Fa - 1) * kSLanes];
#pragma unroll
    for (int u = 0; u < 8; ++u)
      if (t0 + u < Fa) X[(int64_t)(t0 + u) * kSLanes] = (lane < kSBands) ? cur[u] - mu : 0.f;
  }
  // The raw frames of a tiled signal repeat with period P; the masked ones do so only once the
  // start-up transient of the recurrence has died out.  Verify (bit-exact) that everything from
  // the second period on repeats: the lag-product kernels then sum two periods instead of all.
  const int P = b.Pact[pair];
  int ok = (P > 0 && 2 * P + kSStack <= Fa) ? 1 : 0;
  if (ok)
    for (int t0 = P; t0 + P < Fa; t0 += 8) {
      float p0[8], p1[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const int t = min(t0 + u, Fa - P - 1);
        p0[u] = X[(int64_t)t * kSLanes];
        p1[u] = X[(int64_t)(t + P) * kSLanes];
      }
#pragma unroll
      for (int u = 0; u < 8; ++u) ok &= (p0[u] == p1[u]) ? 1 : 0;
    }
  ok = __all_sync(0xffffffffu, ok);
  if (lane == 0) b.perflag[2 * pair + q] = ok;
}

// ---------------------------------------------------------- lag products
// block id -> (type, lag): 0..14 xx d=id; 15..29 yy d=id-15; 30..58 xy d=id-44
__device__ __forceinline__ void block_desc(int blk, int& ta, int& tb, int& d) {
  if (blk < 15) { ta = 0; tb = 0; d = blk; }
  else if (blk < 30) { ta = 1; tb = 1; d = blk - 15; }
  else { ta = 0; tb = 1; d = blk - 44; }
}

// Summation plan of the lag products.  Plain: t = 0 .. Nf - 1, weight 1.  Periodic (frames
// repeat with period P from t = P on, verified by siib_mask_kernel): with Nf - P = q P + rem the
// sum over t >= P is q + 1 times the first rem frames of the period plus q times the rest, so
// only t < 2 P is visited, the A row of frame t weighted by w(t).
struct CovSpan {
  int ne, P, rem, q;
  __device__ __forceinline__ double w(int t) const { return (t < P) ? 1.0 : (t < P + rem) ? (double)(q + 1) : (double)q; }
};
__device__ __forceinline__ CovSpan cov_span(const SiibBuffers& b, int pair, int Nf, bool need_y) {
  CovSpan s;
  s.ne = Nf;
  s.P = Nf;  // plain: every frame below P, weight 1
  s.rem = 0;
  s.q = 0;
  const int P = b.Pact[pair];
  const bool periodic = b.perflag[2 * pair] && (!need_y || b.perflag[2 * pair + 1]);
  if (periodic && P > 0 && Nf >= 2 * P) {
    s.P = P;
    s.q = (Nf - P) / P;
    s.rem = (Nf - P) % P;
    s.ne = 2 * P;
  }
  return s;
}
// Pairs whose x and y features are both verified periodic take the projection route: the 2 P
// distinct stacked frames are projected on the eigenvectors and rho comes from the weighted
// sample moments of the two KLT-domain series (siib_projquad_kernel), which is what the
// quadratic forms u^T Sxy u, u^T Syy u evaluate -- so the yy / xy lag products and the expanded
// Sxy, Syy are never formed for them.
__device__ __forceinline__ bool siib_projected(const SiibBuffers& b, int pair, int Nf) {
  const int P = b.Pact[pair];
  return !b.no_proj && b.perflag[2 * pair] && b.perflag[2 * pair + 1] && P > 0 && Nf >= 2 * P;
}

constexpr int kCovWarps = 8;
constexpr int kCovTile = 64;                        // frames per staged tile
constexpr int kCovRows = kCovTile + kSStack;        // + 15 frames of lag reach (two lags per warp)
constexpr int kCovTasks = 31;                       // warp tasks per pair: two consecutive lags each
constexpr int kCovTasks64 = 8;                      // the xx tasks: FP64 (the rank decision of the Cholesky needs
                                                    // exact duplicates to stay exact); yy / xy / yx run in FP32
constexpr int kCov32Warps = 12;                     // 23 FP32 tasks = two CTAs of 12 warps per pair

// task -> (type of A rows, type of B rows, first lag e >= 0, number of lags 1 or 2, transposed store)
//   0..7   xx lags (0,1) (2,3) ... (12,13) (14)        8..15  yy likewise
//   16..23 xy lags >= 0 likewise                        24..30 yx lags (1,2) ... (13,14): stored
//   transposed as the xy blocks of negative lag (D_xy,-e[jx][jy] = D_yx,e[jy][jx])
__device__ __forceinline__ void cov_task(int task, int& ta, int& tb, int& e, int& nl, bool& tr) {
  tr = false;
  if (task < 24) {
    const int grp = task >> 3, k = task & 7;
    ta = (grp == 1) ? 1 : 0;
    tb = (grp == 0) ? 0 : 1;
    e = 2 * k;
    nl = (k == 7) ? 1 : 2;
  } else {
    ta = 1;
    tb = 0;
    e = 1 + 2 * (task - 24);
    nl = 2;
    tr = true;
  }
}

// Pairs whose x features do not repeat (siib_mask_kernel's check) skip the Cholesky rank decision and go to the FP32
// tridiagonalisation of Sxx: their xx lag products need no more than FP32 either and join the FP32 kernel (the FP64
// kernel runs at the FP64 pipe's rate: 2.8 ms per 1024 pairs for 8 of the 31 tasks, the FP32 one 4.5 ms for 23).
__device__ __forceinline__ bool cov_xx32(const SiibBuffers& b, int pair) { return b.xx32 && !b.perflag[2 * pair]; }

// The eight warps of a CTA walk the time axis together: each tile of frames is converted to
// FP64 once into shared memory and every warp accumulates two 32 x 32 lag blocks (28 x 28 used)
// that share their A rows; the B rows of lag e + 1 at frame t are the B rows of lag e at frame
// t + 1, so one new B row per frame feeds 64 FP64 FMAs per lane (6 LDS.128 per 64 DFMA).
__global__ void __launch_bounds__(kCovWarps * 32) siib_cov_kernel(SiibGeom g, SiibBuffers b) {
  const int lp = blockIdx.y, pair = b.pair_lo + lp, lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const int task = blockIdx.x * kCovWarps + wib;
  const int Fa = b.Fa[pair];
  const int Nf = Fa - (kSStack - 1);
  if (Nf < 1 || cov_xx32(b, pair)) return;
  __shared__ __align__(16) double s_x[2][kCovRows][kSLanes];  // [0] rows weighted (A side), [1] plain (B side)
  const CovSpan span = cov_span(b, pair, Nf, false);
  const bool active = task < kCovTasks64;
  int ta = 0, tb = 0, e = 0, nl = 1;
  bool tr = false;
  if (active) cov_task(task, ta, tb, e, nl, tr);
  const int rg = lane >> 3, cg = lane & 7;  // rows 8 rg .. 8 rg + 7 of A, columns 4 cg .. 4 cg + 3 of B
  const float* __restrict__ X = b.logspec + (g.offF[pair]) * kSLanes;
  const float* __restrict__ Y = b.logspec + (b.totF + g.offF[pair]) * kSLanes;
  double acc0[8][4], acc1[8][4];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc0[i][j] = acc1[i][j] = 0.0;
  for (int t0 = 0; t0 < span.ne; t0 += kCovTile) {
    __syncthreads();
    for (int idx = threadIdx.x; idx < kCovRows * (kSLanes / 4); idx += kCovWarps * 32) {  // xx only: X rows
      const int row = idx / (kSLanes / 4), c4 = idx % (kSLanes / 4);
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (t0 + row < Fa) v = *reinterpret_cast<const float4*>(X + (int64_t)(t0 + row) * kSLanes + 4 * c4);
      const double w = span.w(t0 + row);  // integer weights: exact in FP64
      double* dw = &s_x[0][row][4 * c4];
      double* dp = &s_x[1][row][4 * c4];
      dp[0] = (double)v.x;
      dp[1] = (double)v.y;
      dp[2] = (double)v.z;
      dp[3] = (double)v.w;
      dw[0] = w * (double)v.x;
      dw[1] = w * (double)v.y;
      dw[2] = w * (double)v.z;
      dw[3] = w * (double)v.w;
    }
    __syncthreads();
    if (!active) continue;
    const int nt = min(kCovTile, span.ne - t0);
    const double* A = &s_x[0][0][8 * rg];
    const double* B = &s_x[1][e][4 * cg];
    double2 c01 = *reinterpret_cast<const double2*>(B), c23 = *reinterpret_cast<const double2*>(B + 2);
#pragma unroll 2
    for (int t = 0; t < nt; ++t) {
      const double2 a01 = *reinterpret_cast<const double2*>(A + t * kSLanes);
      const double2 a23 = *reinterpret_cast<const double2*>(A + t * kSLanes + 2);
      const double2 a45 = *reinterpret_cast<const double2*>(A + t * kSLanes + 4);
      const double2 a67 = *reinterpret_cast<const double2*>(A + t * kSLanes + 6);
      const double2 n01 = *reinterpret_cast<const double2*>(B + (t + 1) * kSLanes);      // lag e + 1 now,
      const double2 n23 = *reinterpret_cast<const double2*>(B + (t + 1) * kSLanes + 2);  // lag e next frame
      const double a[8] = {a01.x, a01.y, a23.x, a23.y, a45.x, a45.y, a67.x, a67.y};
      const double c[4] = {c01.x, c01.y, c23.x, c23.y};
      const double n[4] = {n01.x, n01.y, n23.x, n23.y};
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          acc0[i][j] = fma(a[i], c[j], acc0[i][j]);
          acc1[i][j] = fma(a[i], n[j], acc1[i][j]);
        }
      c01 = n01;
      c23 = n23;
    }
  }
  if (!active) return;
  // block ids of siib_expand_kernel: xx lag d -> d, yy -> 15 + d, xy lag d (-14..14) -> 44 + d
#pragma unroll
  for (int l = 0; l < 2; ++l) {
    if (l >= nl) break;
    const int lag = e + l;
    const int blk = tr ? (44 - lag) : (ta == 0 && tb == 0) ? lag : (ta == 1 && tb == 1) ? 15 + lag : 44 + lag;
    double* __restrict__ out = b.base + ((int64_t)lp * kSBlocks + blk) * (kSLanes * kSLanes);
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const double v = l ? acc1[i][j] : acc0[i][j];
        const int ra = 8 * rg + i, cb = 4 * cg + j;
        if (tr) out[cb * kSLanes + ra] = v;
        else out[ra * kSLanes + cb] = v;
      }
  }
}

// FP32 twin of the kernel above for the yy, xy and yx tasks (8..30).  Syy and Sxy are stored
// in FP32 anyway; the sums run in FP32 per 64-frame tile and the tile sums are added to a second
// FP32 accumulator (two-level summation: ~4e-7 relative error over 1900 frames, an order below
// what moves the score by 1e-4).
__global__ void __launch_bounds__(kCov32Warps * 32) siib_cov32_kernel(SiibGeom g, SiibBuffers b) {
  const int lp = blockIdx.y, pair = b.pair_lo + lp, lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  // tasks 8..30 in two CTAs, or 0..30 in three when the xx products are taken here as well
  const bool xx = cov_xx32(b, pair);
  if (!xx && blockIdx.x == 2) return;
  const int task = (xx ? 0 : kCovTasks64) + blockIdx.x * kCov32Warps + wib;
  const int Fa = b.Fa[pair];
  const int Nf = Fa - (kSStack - 1);
  if (Nf < 1 || siib_projected(b, pair, Nf)) return;
  __shared__ __align__(16) float s_x[4][kCovRows][kSLanes];  // [0], [1]: x, y rows weighted (A side); [2], [3]: plain
  const CovSpan span = cov_span(b, pair, Nf, true);
  const bool active = task < kCovTasks;
  int ta = 0, tb = 0, e = 0, nl = 1;
  bool tr = false;
  if (active) cov_task(task, ta, tb, e, nl, tr);
  const int rg = lane >> 3, cg = lane & 7;
  const float* __restrict__ X = b.logspec + (g.offF[pair]) * kSLanes;
  const float* __restrict__ Y = b.logspec + (b.totF + g.offF[pair]) * kSLanes;
  // accumulators as f32x2 pairs of rows (fma.rn.f32x2: the kernel sits at 63 % issue slots and 54 % of the FMA pipe, so
  // halving the FMA instructions is time; the arithmetic per component is unchanged)
  F2 tot0[4][4], tot1[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) tot0[i][j] = tot1[i][j] = f2_pack(0.f, 0.f);
  for (int t0 = 0; t0 < span.ne; t0 += kCovTile) {
    __syncthreads();
    for (int idx = threadIdx.x; idx < 2 * kCovRows * (kSLanes / 4); idx += kCov32Warps * 32) {
      const int q = idx / (kCovRows * (kSLanes / 4)), rem = idx % (kCovRows * (kSLanes / 4));
      const int row = rem / (kSLanes / 4), c4 = rem % (kSLanes / 4);
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (t0 + row < Fa) v = *reinterpret_cast<const float4*>((q ? Y : X) + (int64_t)(t0 + row) * kSLanes + 4 * c4);
      *reinterpret_cast<float4*>(&s_x[2 + q][row][4 * c4]) = v;
      const float w = (float)span.w(t0 + row);
      v.x *= w;
      v.y *= w;
      v.z *= w;
      v.w *= w;
      *reinterpret_cast<float4*>(&s_x[q][row][4 * c4]) = v;
    }
    __syncthreads();
    if (!active) continue;
    const int nt = min(kCovTile, span.ne - t0);
    const float* A = &s_x[ta][0][8 * rg];
    const float* B = &s_x[2 + tb][e][4 * cg];
    F2 acc0[4][4], acc1[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) acc0[i][j] = acc1[i][j] = f2_pack(0.f, 0.f);
    float4 cc = *reinterpret_cast<const float4*>(B);
#pragma unroll 2
    for (int t = 0; t < nt; ++t) {
      const float4 a0 = *reinterpret_cast<const float4*>(A + t * kSLanes);
      const float4 a1 = *reinterpret_cast<const float4*>(A + t * kSLanes + 4);
      const float4 nn = *reinterpret_cast<const float4*>(B + (t + 1) * kSLanes);  // lag e + 1 now, lag e next frame
      const F2 a[4] = {f2_pack(a0.x, a0.y), f2_pack(a0.z, a0.w), f2_pack(a1.x, a1.y), f2_pack(a1.z, a1.w)};
      const F2 c[4] = {f2_pack(cc.x, cc.x), f2_pack(cc.y, cc.y), f2_pack(cc.z, cc.z), f2_pack(cc.w, cc.w)};
      const F2 n[4] = {f2_pack(nn.x, nn.x), f2_pack(nn.y, nn.y), f2_pack(nn.z, nn.z), f2_pack(nn.w, nn.w)};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          acc0[i][j] = f2_fma(a[i], c[j], acc0[i][j]);
          acc1[i][j] = f2_fma(a[i], n[j], acc1[i][j]);
        }
      cc = nn;
    }
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        tot0[i][j] = f2_add(tot0[i][j], acc0[i][j]);
        tot1[i][j] = f2_add(tot1[i][j], acc1[i][j]);
      }
  }
  if (!active) return;
#pragma unroll
  for (int l = 0; l < 2; ++l) {
    if (l >= nl) break;
    const int lag = e + l;
    const int blk = tr ? (44 - lag) : (ta == 0 && tb == 0) ? lag : (ta == 1 && tb == 1) ? 15 + lag : 44 + lag;
    double* __restrict__ out = b.base + ((int64_t)lp * kSBlocks + blk) * (kSLanes * kSLanes);
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        float lo, hi;
        f2_unpack(l ? tot1[i >> 1][j] : tot0[i >> 1][j], lo, hi);
        const double v = (double)((i & 1) ? hi : lo);
        const int ra = 8 * rg + i, cb = 4 * cg + j;
        if (tr) out[cb * kSLanes + ra] = v;
        else out[ra * kSLanes + cb] = v;
      }
  }
}

// ------------------------------------------------------------------ expand
constexpr int kExpThreads = 256;

__global__ void __launch_bounds__(kExpThreads) siib_expand_kernel(SiibGeom g, SiibBuffers b) {
  const int lp = blockIdx.x, pair = b.pair_lo + lp, tid = threadIdx.x, lane = tid & 31, wib = tid >> 5;
  const int Fa = b.Fa[pair];
  const int Nf = Fa - (kSStack - 1);
  if (Nf < 2) return;
  __shared__ double s_rs[2][kSDim];  // row sums of the stacked matrices
  __shared__ float s_edge[2][2 * kSStack][kSLanes];   // frames 0..14 and Nf..Nf+14 of x and y: all the sliding terms
  __shared__ double s_part[kExpThreads / 32][2][kSLanes];
  const float* __restrict__ X = b.logspec + (g.offF[pair]) * kSLanes;
  const float* __restrict__ Y = b.logspec + (b.totF + g.offF[pair]) * kSLanes;
  for (int idx = tid; idx < 2 * 2 * kSStack * kSLanes; idx += kExpThreads) {
    const int q = idx / (2 * kSStack * kSLanes), rr = (idx / kSLanes) % (2 * kSStack), j = idx % kSLanes;
    const int fr = rr < kSStack ? rr : Nf + (rr - kSStack);
    s_edge[q][rr][j] = (fr < Fa) ? (q ? Y : X)[(int64_t)fr * kSLanes + j] : 0.f;
  }
  // sums over t < Nf of every band (k = 0): a warp reads whole 128-byte frame rows, lane = band
  {
    double sx = 0.0, sy = 0.0;
    constexpr int NWp = kExpThreads / 32;
    for (int t0 = wib; t0 < Nf; t0 += 8 * NWp) {   // sixteen independent loads in flight per lane
      float vx[8], vy[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const int t = t0 + u * NWp;
        vx[u] = t < Nf ? X[(int64_t)t * kSLanes + lane] : 0.f;
        vy[u] = t < Nf ? Y[(int64_t)t * kSLanes + lane] : 0.f;
      }
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        sx += (double)vx[u];
        sy += (double)vy[u];
      }
    }
    s_part[wib][0][lane] = sx;
    s_part[wib][1][lane] = sy;
  }
  __syncthreads();
  if (tid < 2 * kSLanes) {
    const int q = tid / kSLanes, j = tid % kSLanes;
    if (j < kSBands) {
      double s = 0.0;
#pragma unroll
      for (int w = 0; w < kExpThreads / 32; ++w) s += s_part[w][q][j];
      s_rs[q][j] = s;
      for (int k = 1; k < kSStack; ++k) {   // then slide by one frame per stack offset
        s += (double)s_edge[q][kSStack + k - 1][j] - (double)s_edge[q][k - 1][j];
        s_rs[q][k * kSBands + j] = s;
      }
    }
  }
  __syncthreads();
  const double inv_nf = 1.0 / (double)Nf;
  double* __restrict__ Sxx = b.Sxx + (int64_t)lp * kSDim * kSDim;
  float* __restrict__ Sxy = b.Sxy + (int64_t)lp * kSDim * kSDim;
  float* __restrict__ Syy = b.Syy + (int64_t)lp * kSDim * kSDim;
  const double* __restrict__ base = b.base + (int64_t)lp * kSBlocks * (kSLanes * kSLanes);
  // Sxx (FP64, full): every (block, j1, j2) diagonal is walked twice, once with the lanes along j2, writing row
  // segments of the upper block triangle, once with the lanes along j1, writing the mirrored element into the lower
  // one -- so both stores are 28 consecutive values of a row (a lane-transposed store costs one L1 wavefront per
  // element, which bound this kernel).
  for (int it = tid; it < 2 * 15 * kSBands * kSBands; it += kExpThreads) {
    const bool mirror = it >= 15 * kSBands * kSBands;
    const int idx = mirror ? it - 15 * kSBands * kSBands : it;
    const int d = idx / (kSBands * kSBands), rem = idx % (kSBands * kSBands);
    const int j1 = mirror ? rem % kSBands : rem / kSBands, j2 = mirror ? rem / kSBands : rem % kSBands;
    if (mirror && d == 0) continue;   // a diagonal block is its own mirror image
    const float (*A)[kSLanes] = s_edge[0];
    int k1 = 0, k2 = d;
    double r = base[d * (kSLanes * kSLanes) + j1 * kSLanes + j2];
    const int steps = kSStack - d;
    for (int m = 0; m < steps; ++m) {
      const int a = k1 * kSBands + j1, c = k2 * kSBands + j2;
      const double v = r - s_rs[0][a] * s_rs[0][c] * inv_nf;
      if (!mirror) Sxx[(int64_t)a * kSDim + c] = v;
      else Sxx[(int64_t)c * kSDim + a] = v;
      if (m + 1 == steps) break;
      // slide the summation window by one frame
      r += (double)A[kSStack + k1][j1] * (double)A[kSStack + k2][j2] - (double)A[k1][j1] * (double)A[k2][j2];
      ++k1;
      ++k2;
    }
  }
  if (siib_projected(b, pair, Nf)) return;  // siib_projquad_kernel needs neither Syy nor Sxy
  // Syy and Sxy enter the score only through the quadratic forms u^T S u (qf::quadform_kernel), which see the symmetric
  // part of S: they are stored FOLDED onto the lower triangle -- F[c][a] = S[c][a] + S[a][c] for c > a, F[a][a] = S[a][a],
  // the upper triangle is never written nor read -- so the products visit half the matrix.  Element a = (k1, j1),
  // c = (k1 + d, j2), d >= 0 (d = 0: j2 >= j1): Syy[a][c] = Syy[c][a] comes from the yy block of lag d; Sxy[a][c] from
  // the xy block of lag d and Sxy[c][a] from the one of lag -d, each with its own sliding window.  Lanes along j1: the
  // stores are consecutive values of row c.
  for (int it = tid; it < 2 * 15 * kSBands * kSBands; it += kExpThreads) {
    const bool xy = it >= 15 * kSBands * kSBands;
    const int idx = xy ? it - 15 * kSBands * kSBands : it;
    const int d = idx / (kSBands * kSBands), rem = idx % (kSBands * kSBands);
    const int j1 = rem % kSBands, j2 = rem / kSBands;
    if (d == 0 && j2 < j1) continue;
    const bool diag = d == 0 && j1 == j2;
    int k1 = 0, k2 = d;
    const int steps = kSStack - d;
    if (!xy) {
      const float (*B)[kSLanes] = s_edge[1];
      double r = base[(15 + d) * (kSLanes * kSLanes) + j1 * kSLanes + j2];
      for (int m = 0; m < steps; ++m) {
        const int a = k1 * kSBands + j1, c = k2 * kSBands + j2;
        const double v = r - s_rs[1][a] * s_rs[1][c] * inv_nf;
        Syy[(int64_t)c * kSDim + a] = (float)(diag ? v : 2.0 * v);
        if (m + 1 == steps) break;
        r += (double)B[kSStack + k1][j1] * (double)B[kSStack + k2][j2] - (double)B[k1][j1] * (double)B[k2][j2];
        ++k1;
        ++k2;
      }
    } else {
      const float (*A)[kSLanes] = s_edge[0];
      const float (*B)[kSLanes] = s_edge[1];
      // r1: x_a y_c (x stack k1, band j1; y stack k2, band j2), r2: x_c y_a (x stack k2, band j2; y stack k1, band j1)
      double r1 = base[(44 + d) * (kSLanes * kSLanes) + j1 * kSLanes + j2];
      double r2 = base[(44 - d) * (kSLanes * kSLanes) + j2 * kSLanes + j1];
      for (int m = 0; m < steps; ++m) {
        const int a = k1 * kSBands + j1, c = k2 * kSBands + j2;
        const double v1 = r1 - s_rs[0][a] * s_rs[1][c] * inv_nf;
        const double v2 = r2 - s_rs[0][c] * s_rs[1][a] * inv_nf;
        Sxy[(int64_t)c * kSDim + a] = (float)(diag ? v1 : v1 + v2);
        if (m + 1 == steps) break;
        r1 += (double)A[kSStack + k1][j1] * (double)B[kSStack + k2][j2] - (double)A[k1][j1] * (double)B[k2][j2];
        r2 += (double)A[kSStack + k2][j2] * (double)B[kSStack + k1][j1] - (double)A[k2][j2] * (double)B[k1][j1];
        ++k1;
        ++k2;
      }
    }
  }
}

// --------------------------------------------------------------- Cholesky
// Diagonally pivoted Cholesky in FP64, blocked right-looking (LAPACK dpstrf's scheme): the
// current panel of 32 columns of L lives in shared memory, each elimination step costs one
// block barrier, one coalesced row of the trailing matrix and a <= 31-term dot product from
// shared memory, and the trailing matrix is updated once per panel (420^2 * 8 B read + write
// per panel instead of re-reading every earlier column of L at every step).  Rows are never
// permuted: thread a owns original row a; only the order in which pivots are taken defines the
// columns of L, and the one-sided Jacobi that follows does not care about the row order.
constexpr int kCholThreads = 448;
constexpr int kCholW = 32;

__global__ void __launch_bounds__(kCholThreads) siib_chol_kernel(SiibGeom g, SiibBuffers b, int skip_nonperiodic) {
  const int lp = blockIdx.x, pair = b.pair_lo + lp, tid = threadIdx.x, lane = tid & 31, wib = tid >> 5;
  constexpr int NW = kCholThreads / 32;
  const int Nf = b.Fa[pair] - (kSStack - 1);
  float* __restrict__ G = b.G + (int64_t)lp * kSDim * kSLd;
  if (Nf < 2) {
    if (tid == 0) b.rank[pair] = 0;
    return;
  }
  if (skip_nonperiodic && !b.perflag[2 * pair]) {
    // A tiled signal that does not repeat its frames has a full-rank covariance (almost surely):
    // such pairs go straight to the tridiagonalisation path (siib_eig.cu), which needs no factor
    // and treats a numerically singular matrix the same way (eigenvalues <= 1e-10 lambda_max).
    if (tid == 0) b.rank[pair] = kSDim;
    return;
  }
  const double* __restrict__ A0 = b.Sxx + (int64_t)lp * kSDim * kSDim;
  double* __restrict__ W = b.Lc + (int64_t)lp * kSDim * kSDim;  // trailing matrix after the first panel
  extern __shared__ __align__(16) double s_panel[];              // Lp[m][a], m < 32, a < 420
  __shared__ double s_rv[2][NW];
  __shared__ double s_mx[NW];
  __shared__ int s_ri[2][NW];
  __shared__ unsigned char s_done[kSDim];
  const bool own = tid < kSDim;
  bool done = false;
  double mydg = own ? A0[(int64_t)tid * kSDim + tid] : -1.0e300;
  if (own) s_done[tid] = 0;
  double tol = 0.0;
  int rank = kSDim;
  bool stop = false;
  for (int k0 = 0; k0 < kSDim && !stop; k0 += kCholW) {
    const double* __restrict__ As = (k0 == 0) ? A0 : W;
    const int kb = min(kCholW, kSDim - k0);
    int j = 0;
    for (; j < kb; ++j) {
      const int k = k0 + j;
      double v = (own && !done) ? mydg : -1.0e300;
      int vi = tid;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        const double ov = __shfl_xor_sync(0xffffffffu, v, o);
        const int oi = __shfl_xor_sync(0xffffffffu, vi, o);
        if (ov > v || (ov == v && oi < vi)) {
          v = ov;
          vi = oi;
        }
      }
      if (lane == 0) {
        s_rv[k & 1][wib] = v;
        s_ri[k & 1][wib] = vi;
      }
      __syncthreads();  // also orders the previous step's panel writes before this step's reads
      double piv = s_rv[k & 1][0];
      int astar = s_ri[k & 1][0];
#pragma unroll
      for (int w = 1; w < NW; ++w) {
        const double ov = s_rv[k & 1][w];
        const int oi = s_ri[k & 1][w];
        if (ov > piv || (ov == piv && oi < astar)) {
          piv = ov;
          astar = oi;
        }
      }
      if (k == 0) tol = piv * 1.0e-10;
      if (!(piv > tol)) {
        rank = k;
        stop = true;
        break;
      }
      const double lkk = sqrt(piv);
      double l = 0.0;
      if (own && !done) {
        if (tid == astar) {
          l = lkk;
          done = true;
          s_done[tid] = 1;
        } else {
          // after the first panel only the lower triangle of the work matrix is maintained: row astar up
          // to the diagonal, column astar below it
          double s0 = (k0 == 0 || tid <= astar) ? As[(int64_t)astar * kSDim + tid] : As[(int64_t)tid * kSDim + astar], s1 = 0.0;
          int m = 0;
          for (; m + 2 <= j; m += 2) {
            s0 = fma(-s_panel[m * kSDim + tid], s_panel[m * kSDim + astar], s0);
            s1 = fma(-s_panel[(m + 1) * kSDim + tid], s_panel[(m + 1) * kSDim + astar], s1);
          }
          if (m < j) s0 = fma(-s_panel[m * kSDim + tid], s_panel[m * kSDim + astar], s0);
          l = (s0 + s1) / lkk;
          mydg -= l * l;
        }
      }
      if (own) s_panel[j * kSDim + tid] = l;
      G[(int64_t)k * kSLd + tid] = (float)l;  // rows 420..447 (and eliminated rows) are zero
    }
    __syncthreads();
    if (stop || k0 + kb >= kSDim) break;
    {
      // The diagonal is maintained step by step: if no admissible pivot is left the factorisation
      // is complete (numerical rank = k0 + kb) and the trailing update would be wasted work.
      double v = (own && !done) ? mydg : -1.0e300;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
      if (lane == 0) s_mx[wib] = v;
      __syncthreads();
      double vm = s_mx[0];
#pragma unroll
      for (int w = 1; w < NW; ++w) vm = fmax(vm, s_mx[w]);
      if (!(vm > tol)) {
        rank = k0 + kb;
        break;
      }
    }
    // trailing update with the finished panel: W[a][c] = As[a][c] - sum_m Lp[m][a] Lp[m][c], c <= a (the matrix is
    // symmetric: half the flops and half the traffic of the full update)
    for (int a0 = 4 * wib; a0 < kSDim; a0 += 4 * NW) {
      bool live[4];
      bool any = false;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        live[i] = !s_done[a0 + i];
        any |= live[i];
      }
      if (!any) continue;
      // 4 rows x 2 columns per lane: 2 + 2 shared-memory loads (the row values as two 128-bit
      // broadcasts) per 8 FP64 FMAs
      for (int cb = 0; cb <= a0 + 3; cb += 64) {   // lower triangle: columns <= row
        const int c = cb + lane, c1 = c + 32;
        const bool h0 = c <= a0 + 3 && !s_done[c], h1 = (c1 <= a0 + 3) && !s_done[c1];
        if (!h0 && !h1) continue;
        const int cs = min(c, kSDim - 1), c1s = min(c1, kSDim - 1);
        double acc0[4] = {0.0, 0.0, 0.0, 0.0}, acc1[4] = {0.0, 0.0, 0.0, 0.0};
        // the old values are requested before the product loop, so their latency hides behind it
        double old0[4], old1[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          old0[i] = (live[i] && h0 && c <= a0 + i) ? As[(int64_t)(a0 + i) * kSDim + c] : 0.0;
          old1[i] = (live[i] && h1 && c1 <= a0 + i) ? As[(int64_t)(a0 + i) * kSDim + c1] : 0.0;
        }
#pragma unroll 8
        for (int m = 0; m < kCholW; ++m) {
          const double l0 = s_panel[m * kSDim + cs], l1 = s_panel[m * kSDim + c1s];
          const double2 r01 = *reinterpret_cast<const double2*>(&s_panel[m * kSDim + a0]);
          const double2 r23 = *reinterpret_cast<const double2*>(&s_panel[m * kSDim + a0 + 2]);
          const double rr[4] = {r01.x, r01.y, r23.x, r23.y};
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            acc0[i] = fma(rr[i], l0, acc0[i]);
            acc1[i] = fma(rr[i], l1, acc1[i]);
          }
        }
#pragma unroll
        for (int i = 0; i < 4; ++i)
          if (live[i]) {
            if (h0 && c <= a0 + i) W[(int64_t)(a0 + i) * kSDim + c] = old0[i] - acc0[i];
            if (h1 && c1 <= a0 + i) W[(int64_t)(a0 + i) * kSDim + c1] = old1[i] - acc1[i];
          }
      }
    }
    __syncthreads();
  }
  if (tid == 0) b.rank[pair] = rank;
}

// ----------------------------------------------------------------- Jacobi
constexpr int kJacH = 4;                  // columns per block
constexpr int kJacE = kSLd / 32;          // 14 elements per lane per column
constexpr int kJacMaxSweeps = 15;
constexpr float kJacTol = 1.5e-6f;

// A block of four columns in registers.  Columns are held *scaled*: true column = scl * v.  A
// Hestenes rotation (p, q) <- (c (p - t q), c (q + t p)) then costs two FMAs per element pair
// instead of four operations -- the common factor c goes into the scales -- and the squared
// norms are carried along by the update formulas |p|^2 -= t p.q, |q|^2 += t p.q.  Scales and
// norms live next to the columns in shared memory (s_meta) for the length of a round; they are
// folded back / recomputed exactly once per round (<= 111 rotations per column: no underflow,
// norm drift ~1e-5).
struct ColBlock {
  float v[kJacH][kJacE];
  float nrm[kJacH];
  float scl[kJacH];
};

__device__ __forceinline__ float fast_rcp(float x) {
  float r;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}
__device__ __forceinline__ float fast_sqrt(float x) {
  float r;
  asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}
__device__ __forceinline__ float fast_rsqrt(float x) {
  float r;
  asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}
// Rotation parameters for one column pair.  gs = dot product of the scaled columns; a = |p|^2,
// b = |q|^2 (true, updated in place); dp, dq the scales (updated in place); tp, tq the factors of
// the scaled update p~ -= tp q~, q~ += tq p~.  Branch-free (every lane holds the same values; a
// branch here costs a convergence barrier per rotation): a pair that is already orthogonal to
// working precision gets t = 0.  Hardware approximations (rcp / sqrt / rsqrt, 1-2 ulp) are enough:
// the angle is re-estimated from exact dot products the next time the pair meets.
__device__ __forceinline__ void rot_params(float gs, float& a, float& b, float& dp, float& dq, float& tp, float& tq,
                                           int& nrot) {
  const float g = gs * (dp * dq);
  const bool rot = (g * g > (kJacTol * kJacTol) * (a * b)) && (a > 0.f) && (b > 0.f);
  const float zeta = (b - a) * fast_rcp(2.0f * g);             // +-inf / NaN when g == 0: masked below
  const float az = fabsf(zeta);
  float t = fast_rcp(az + fast_sqrt(fmaf(zeta, zeta, 1.0f)));  // 1 / (|z| + sqrt(1 + z^2)); 0 for huge |z|
  t = rot ? copysignf(t, zeta) : 0.0f;
  const float c = fast_rsqrt(fmaf(t, t, 1.0f));
  const float tg = rot ? t * g : 0.0f;
  a -= tg;
  b += tg;
  const float rpq = dq * fast_rcp(dp);
  tp = t * rpq;
  tq = t * fast_rcp(rpq);
  dp *= c;
  dq *= c;
  nrot += rot ? 1 : 0;
}
__device__ __forceinline__ float dot14(const float (&p)[kJacE], const float (&q)[kJacE]) {
  float g0 = 0.f, g1 = 0.f;
#pragma unroll
  for (int e = 0; e < kJacE; e += 2) {
    g0 = fmaf(p[e], q[e], g0);
    g1 = fmaf(p[e + 1], q[e + 1], g1);
  }
  return g0 + g1;
}
__device__ __forceinline__ void apply_rot(float (&p)[kJacE], float (&q)[kJacE], float tp, float tq) {
#pragma unroll
  for (int e = 0; e < kJacE; ++e) {
    const float a = p[e], bq = q[e];
    p[e] = fmaf(-tp, bq, a);
    q[e] = fmaf(tq, a, bq);
  }
}
// Warp sums of four values, every lane gets all four: two halving exchanges leave each lane with
// one partial, three butterfly steps finish it, four broadcasts hand the results out
// (10 SHFL + 6 FADD instead of 20 + 20).  The summation order is the same on every lane.
__device__ __forceinline__ void warp_sum4(float& g0, float& g1, float& g2, float& g3, int lane) {
  const bool h16 = lane & 16, h8 = lane & 8;
  const float s0 = h16 ? g0 : g2, s1 = h16 ? g1 : g3;
  float k0 = h16 ? g2 : g0, k1 = h16 ? g3 : g1;
  k0 += __shfl_xor_sync(0xffffffffu, s0, 16);
  k1 += __shfl_xor_sync(0xffffffffu, s1, 16);
  const float s = h8 ? k0 : k1;
  float k = h8 ? k1 : k0;
  k += __shfl_xor_sync(0xffffffffu, s, 8);
  k += __shfl_xor_sync(0xffffffffu, k, 4);
  k += __shfl_xor_sync(0xffffffffu, k, 2);
  k += __shfl_xor_sync(0xffffffffu, k, 1);
  g0 = __shfl_sync(0xffffffffu, k, 0);
  g1 = __shfl_sync(0xffffffffu, k, 8);
  g2 = __shfl_sync(0xffffffffu, k, 16);
  g3 = __shfl_sync(0xffffffffu, k, 24);
}
// four mutually independent rotations (p0,q0) .. (p3,q3), issued together so that the four
// scalar chains overlap.  Arguments per column: values, squared norm, scale.
#define NELE_ROT4_(P0, NP0, DP0, Q0, NQ0, DQ0, P1, NP1, DP1, Q1, NQ1, DQ1, P2, NP2, DP2, Q2, NQ2, DQ2, P3, NP3, DP3, \
                  Q3, NQ3, DQ3)                                                                                      \
  {                                                                                                                  \
    float g0 = dot14(P0, Q0), g1 = dot14(P1, Q1), g2 = dot14(P2, Q2), g3 = dot14(P3, Q3);                            \
    warp_sum4(g0, g1, g2, g3, lane);                                                                                 \
    float tp0, tq0, tp1, tq1, tp2, tq2, tp3, tq3;                                                                    \
    rot_params(g0, NP0, NQ0, DP0, DQ0, tp0, tq0, nrot);                                                              \
    rot_params(g1, NP1, NQ1, DP1, DQ1, tp1, tq1, nrot);                                                              \
    rot_params(g2, NP2, NQ2, DP2, DQ2, tp2, tq2, nrot);                                                              \
    rot_params(g3, NP3, NQ3, DP3, DQ3, tp3, tq3, nrot);                                                              \
    if (tp0 != 0.f || tp1 != 0.f || tp2 != 0.f || tp3 != 0.f) { /* one warp-uniform branch */                        \
      apply_rot(P0, Q0, tp0, tq0);                                                                                   \
      apply_rot(P1, Q1, tp1, tq1);                                                                                   \
      apply_rot(P2, Q2, tp2, tq2);                                                                                   \
      apply_rot(P3, Q3, tp3, tq3);                                                                                   \
    }                                                                                                                \
  }
#define NELE_COL(B, h) B.v[h], B.nrm[h], B.scl[h]
#define NELE_ROT4(...) NELE_ROT4_(__VA_ARGS__)

// Shared-memory / cluster version of the tournament.  The r columns are cut into S = 2 CL
// super-blocks; every CTA of a CL-wide thread-block cluster keeps two super-blocks in shared
// memory (2 x 56 columns x 1792 B = 196 KB) and rotates every column pair inside and between
// them with the columns held in registers (4 + 4 per warp, four independent rotations in
// flight).  Between the S - 1 super-rounds of a sweep the CTAs swap super-blocks through
// global memory (L2) and meet at a cluster barrier; CL = 1 (r <= 112) never leaves the SM.
constexpr int kJ2Warps = 12;       // cluster kernel
constexpr int kJ2WarpsSmall = 12;  // single-CTA kernel
constexpr int kJ2MaxSb = 56;       // columns per super-block (14 blocks of 4)

__device__ __forceinline__ void load_block_s(ColBlock& c, const float* slot, const float2* meta, int first, int lane) {
#pragma unroll
  for (int h = 0; h < kJacH; ++h) {
    const float* col = slot + (first + h) * kSLd + lane;
#pragma unroll
    for (int e = 0; e < kJacE; ++e) c.v[h][e] = col[e * 32];
    const float2 m = meta[first + h];
    c.nrm[h] = m.x;
    c.scl[h] = m.y;
  }
}
__device__ __forceinline__ void store_block_s(const ColBlock& c, float* slot, float2* meta, int first, int lane) {
#pragma unroll
  for (int h = 0; h < kJacH; ++h) {
    float* col = slot + (first + h) * kSLd + lane;
#pragma unroll
    for (int e = 0; e < kJacE; ++e) col[e * 32] = c.v[h][e];
    if (lane == 0) meta[first + h] = make_float2(c.nrm[h], c.scl[h]);
  }
}
__device__ __forceinline__ void rotate_within(ColBlock& P, ColBlock& Q, int& nrot, int lane) {
  NELE_ROT4(NELE_COL(P, 0), NELE_COL(P, 1), NELE_COL(P, 2), NELE_COL(P, 3),
            NELE_COL(Q, 0), NELE_COL(Q, 1), NELE_COL(Q, 2), NELE_COL(Q, 3));
  NELE_ROT4(NELE_COL(P, 0), NELE_COL(P, 2), NELE_COL(P, 1), NELE_COL(P, 3),
            NELE_COL(Q, 0), NELE_COL(Q, 2), NELE_COL(Q, 1), NELE_COL(Q, 3));
  NELE_ROT4(NELE_COL(P, 0), NELE_COL(P, 3), NELE_COL(P, 1), NELE_COL(P, 2),
            NELE_COL(Q, 0), NELE_COL(Q, 3), NELE_COL(Q, 1), NELE_COL(Q, 2));
}
__device__ __forceinline__ void rotate_cross(ColBlock& P, ColBlock& Q, int& nrot, int lane) {
  NELE_ROT4(NELE_COL(P, 0), NELE_COL(Q, 0), NELE_COL(P, 1), NELE_COL(Q, 1),
            NELE_COL(P, 2), NELE_COL(Q, 2), NELE_COL(P, 3), NELE_COL(Q, 3));
  NELE_ROT4(NELE_COL(P, 0), NELE_COL(Q, 1), NELE_COL(P, 1), NELE_COL(Q, 2),
            NELE_COL(P, 2), NELE_COL(Q, 3), NELE_COL(P, 3), NELE_COL(Q, 0));
  NELE_ROT4(NELE_COL(P, 0), NELE_COL(Q, 2), NELE_COL(P, 1), NELE_COL(Q, 3),
            NELE_COL(P, 2), NELE_COL(Q, 0), NELE_COL(P, 3), NELE_COL(Q, 1));
  NELE_ROT4(NELE_COL(P, 0), NELE_COL(Q, 3), NELE_COL(P, 1), NELE_COL(Q, 0),
            NELE_COL(P, 2), NELE_COL(Q, 1), NELE_COL(P, 3), NELE_COL(Q, 2));
}
// circle-method pairing of `np` players (np even), round `round`, table k -> (i, j)
__device__ __forceinline__ void circle_pair(int np, int round, int k, int& i, int& j) {
  if (k == 0) {
    i = np - 1;
    j = round;
  } else {
    i = (round + k) % (np - 1);
    j = (round - k + (np - 1)) % (np - 1);
  }
}

template <int CL, int NWARP>
__global__ void __launch_bounds__(NWARP * 32) siib_jacobi2_kernel(SiibGeom g, SiibBuffers b, int rank_lo, int rank_hi) {
  cg::cluster_group cluster = cg::this_cluster();
  const int crank = (CL == 1) ? 0 : (int)cluster.block_rank();
  const int lp = blockIdx.x / CL, pair = b.pair_lo + lp, lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const int r = b.rank[pair];
  if (r < rank_lo || r > rank_hi) return;  // same decision in every CTA of the cluster
  float* __restrict__ G = b.G + (int64_t)lp * kSDim * kSLd;
  constexpr int S = 2 * CL;
  const int sbp = 4 * ((((r + S - 1) / S) + 3) / 4);  // columns per super-block, multiple of 4, <= 56
  const int nblk = sbp / kJacH;                        // 4-column blocks per super-block
  const int nbe = (nblk + 1) & ~1;                     // padded to even for the inner tournament
  extern __shared__ __align__(16) float s_slot[];      // [2][sbp][448]
  float* slotA = s_slot;
  float* slotB = s_slot + sbp * kSLd;
  __shared__ float2 s_meta[2][kJ2MaxSb];               // (squared norm, scale) of every resident column
  __shared__ int s_rot;
  int32_t* __restrict__ grot = b.sweep_rot + (int64_t)pair * 16;

  auto load_super = [&](float* slot, int u) {
    // columns [u sbp, (u + 1) sbp) of G, zero beyond r; 128-bit copies
    const int c0 = u * sbp;
    for (int idx = threadIdx.x; idx < sbp * (kSLd / 4); idx += NWARP * 32) {
      const int c = idx / (kSLd / 4), q4 = idx % (kSLd / 4);
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (c0 + c < r) v = *reinterpret_cast<const float4*>(G + (int64_t)(c0 + c) * kSLd + 4 * q4);
      *reinterpret_cast<float4*>(slot + c * kSLd + 4 * q4) = v;
    }
  };
  auto store_super = [&](const float* slot, int u) {
    const int c0 = u * sbp;
    for (int idx = threadIdx.x; idx < sbp * (kSLd / 4); idx += NWARP * 32) {
      const int c = idx / (kSLd / 4), q4 = idx % (kSLd / 4);
      if (c0 + c < r)
        *reinterpret_cast<float4*>(G + (int64_t)(c0 + c) * kSLd + 4 * q4) = *reinterpret_cast<const float4*>(slot + c * kSLd + 4 * q4);
    }
  };
  // fold the scales into the columns (fresh = false) and recompute the norms exactly
  auto refresh_meta = [&](bool fresh) {
    for (int c = wib; c < 2 * sbp; c += NWARP) {
      const int sl = c >= sbp, cc = sl ? c - sbp : c;
      float* col = (sl ? slotB : slotA) + cc * kSLd + lane;
      const float d = fresh ? 1.0f : s_meta[sl][cc].y;
      float ss = 0.f;
#pragma unroll
      for (int e = 0; e < kJacE; ++e) {
        const float v = col[e * 32] * d;
        col[e * 32] = v;
        ss = fmaf(v, v, ss);
      }
      ss = warp_sum(ss);
      if (lane == 0) s_meta[sl][cc] = make_float2(ss, 1.0f);
    }
  };

  int ua = 0, ub = 1;
  if (CL == 1) {
    load_super(slotA, 0);
    load_super(slotB, 1);
    __syncthreads();
  }
  int sweeps = 0;
  for (; sweeps < kJacMaxSweeps; ++sweeps) {
    int nrot = 0;
    for (int rho = 0; rho < S - 1; ++rho) {
      if (CL > 1) {
        circle_pair(S, rho, crank, ua, ub);
        load_super(slotA, ua);
        load_super(slotB, ub);
        __syncthreads();
      }
      refresh_meta(CL > 1 || sweeps == 0);
      __syncthreads();
      if (rho == 0) {
        // every pair inside super-block A and inside super-block B (once per sweep: in round 0 the
        // S super-blocks are spread over the CL CTAs, two each)
        for (int lr = 0; lr < nbe - 1; ++lr) {
          for (int item = wib; item < nbe; item += NWARP) {
            const int sl = (item < nbe / 2) ? 0 : 1;
            float* slot = sl ? slotB : slotA;
            float2* meta = s_meta[sl];
            int bi, bj;
            circle_pair(nbe, lr, item % (nbe / 2), bi, bj);
            if (bi > bj) {
              const int t = bi;
              bi = bj;
              bj = t;
            }
            const bool pad = bj >= nblk;  // partner is the padding block
            if (pad && lr != 0) continue;
            ColBlock P, Q;
            load_block_s(P, slot, meta, bi * kJacH, lane);
            if (!pad) load_block_s(Q, slot, meta, bj * kJacH, lane);
            else {
#pragma unroll
              for (int h = 0; h < kJacH; ++h) {
                Q.nrm[h] = 0.f;
                Q.scl[h] = 1.f;
#pragma unroll
                for (int e = 0; e < kJacE; ++e) Q.v[h][e] = 0.f;
              }
            }
            if (lr == 0) rotate_within(P, Q, nrot, lane);
            if (!pad) rotate_cross(P, Q, nrot, lane);
            store_block_s(P, slot, meta, bi * kJacH, lane);
            if (!pad) store_block_s(Q, slot, meta, bj * kJacH, lane);
          }
          __syncthreads();
        }
      }
      // every pair between super-block A and super-block B.  Pair k = sh * nblk + a joins block a
      // of A with block (a + sh) mod nblk of B; any window of fewer than nblk consecutive k is
      // conflict free (it spans at most two shifts), so NWARP warps take NWARP pairs per barrier
      // even when nblk is not a multiple of NWARP.
      {
        const int npairs = nblk * nblk;
        const int per = (nblk <= NWARP) ? nblk : NWARP;
        for (int k0 = 0; k0 < npairs; k0 += per) {
          const int k = k0 + wib;
          if (wib < per && k < npairs) {
            const int a = k % nblk, bq = (a + k / nblk) % nblk;
            ColBlock P, Q;
            load_block_s(P, slotA, s_meta[0], a * kJacH, lane);
            load_block_s(Q, slotB, s_meta[1], bq * kJacH, lane);
            rotate_cross(P, Q, nrot, lane);
            store_block_s(P, slotA, s_meta[0], a * kJacH, lane);
            store_block_s(Q, slotB, s_meta[1], bq * kJacH, lane);
          }
          __syncthreads();
        }
      }
      if (CL > 1) {
        refresh_meta(false);  // the global copy holds true (unscaled) columns
        __syncthreads();
        store_super(slotA, ua);
        store_super(slotB, ub);
        __threadfence();
        cluster.sync();
      }
    }
    // rotations of this sweep over the whole cluster
    if (threadIdx.x == 0) s_rot = 0;
    __syncthreads();
    if (lane == 0 && nrot) atomicAdd(&s_rot, nrot);
    __syncthreads();
    int tot = s_rot;
    if (CL > 1) {
      if (threadIdx.x == 0 && tot) atomicAdd(&grot[sweeps], tot);
      __threadfence();
      cluster.sync();
      tot = *reinterpret_cast<volatile int32_t*>(&grot[sweeps]);
      cluster.sync();
    } else {
      if (threadIdx.x == 0) grot[sweeps] = tot;
      __syncthreads();
    }
    // Quadratic convergence: a sweep that still found only a handful of pairs above the
    // threshold (out of r (r - 1) / 2) leaves nothing for the next one; skip the empty
    // verification sweep.
    if (tot <= r / 8) {
      ++sweeps;
      break;
    }
  }
  if (CL == 1) {
    refresh_meta(false);
    __syncthreads();
    store_super(slotA, 0);
    store_super(slotB, 1);
  }
  if (threadIdx.x == 0 && crank == 0) b.sweeps[pair] = sweeps;
}

// -------------------------------------------------- projection route (periodic pairs)
// rho_j of a pair whose stacked frames repeat (siib_projected): with the summation plan of
// cov_span only the first 2 P stacked frames are distinct, so
//   u^T Sxy u = sum_t w(t) xk(t) yk(t) - (sum_t w xk)(sum_t w yk) / Nf,   xk(t) = u^T xs(t),
// and likewise for xx and yy: 2 P x r dot products of length 420 per signal instead of two
// r x 420 x 420 quadratic forms over matrices that then need not exist.  Stacked frame t is the
// contiguous slab logspec[t .. t + 14][28], so for one band j a thread that owns 8 consecutive
// frames slides a 22-value register window past the 15 stack offsets: 59 shared-memory loads
// feed 960 FMAs.
//   CTA = pair, 128 threads: thread (tc, tt) = components 4 tc .. 4 tc + 3 of a 32-component tile,
//   frames 8 tt .. 8 tt + 7 of a 128-frame tile.  sU [420][32] holds the unit eigenvectors with
//   the 16-byte chunks XOR-swizzled by the row (the fill writes along a column), the slabs
//   [28][142] are stored with an (i + i / 8) skew so that lanes 8 frames apart hit distinct banks.
constexpr int kPqThreads = 128, kPqC = 32, kPqT = 128, kPqF = 8;
constexpr int kPqRows = kPqT + kSStack - 1;               // 142 feature rows per time tile
constexpr int kPqLd = kPqRows + (kPqRows >> 3) + 2;        // 161: skewed row length, odd
constexpr int kPqSmem = (kSDim * kPqC + 2 * kSBands * kPqLd) * (int)sizeof(float);

__global__ void __launch_bounds__(kPqThreads) siib_projquad_kernel(SiibGeom g, SiibBuffers b) {
  const int lp = blockIdx.x, pair = b.pair_lo + lp, tid = threadIdx.x, lane = tid & 31, wib = tid >> 5;
  const int Fa = b.Fa[pair];
  const int Nf = Fa - (kSStack - 1);
  if (b.M[pair] <= 0 || (double)Fa / 80.0 < 20.0 || Nf < 2) return;  // quad_finish_kernel reports these
  if (!siib_projected(b, pair, Nf)) return;
  const CovSpan span = cov_span(b, pair, Nf, true);
  const int r = b.rank[pair];
  extern __shared__ __align__(16) float s_pq[];
  float* sU = s_pq;                       // [420][32], chunk c / 4 of row i at chunk (c / 4) ^ (i & 7)
  float* sX = sU + kSDim * kPqC;          // [28][kPqLd]
  float* sY = sX + kSBands * kPqLd;
  __shared__ float s_nrm[kPqC];
  __shared__ double s_info[kSDim];
  const float* __restrict__ G = b.G + (int64_t)lp * kSDim * kSLd;
  const float* __restrict__ X = b.logspec + (g.offF[pair]) * kSLanes;
  const float* __restrict__ Y = b.logspec + (b.totF + g.offF[pair]) * kSLanes;
  float* __restrict__ lam_out = b.lambda + (int64_t)pair * kSDim;
  float* __restrict__ rho_out = b.rho + (int64_t)pair * kSDim;
  const int tt = tid & 15, tc = tid >> 4;
  for (int c0 = 0; c0 < r; c0 += kPqC) {
    __syncthreads();
    for (int c = wib; c < kPqC; c += kPqThreads / 32) {
      float ss = 0.f;
      if (c0 + c < r) {
        const float* col = G + (int64_t)(c0 + c) * kSLd;
        for (int i = lane; i < kSDim; i += 32) ss = fmaf(col[i], col[i], ss);
      }
      ss = warp_sum(ss);
      if (lane == 0) s_nrm[c] = ss;
    }
    __syncthreads();
    for (int idx = tid; idx < kSDim * kPqC; idx += kPqThreads) {
      const int c = idx / kSDim, i = idx % kSDim;
      const float nn = s_nrm[c];
      const float v = (c0 + c < r && nn > 0.f) ? G[(int64_t)(c0 + c) * kSLd + i] * rsqrtf(nn) : 0.f;
      sU[i * kPqC + ((((c >> 2) ^ (i & 7)) << 2) | (c & 3))] = v;
    }
    double m1x[4], m1y[4], mxx[4], myy[4], mxy[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) m1x[i] = m1y[i] = mxx[i] = myy[i] = mxy[i] = 0.0;
    for (int t0 = 0; t0 < span.ne; t0 += kPqT) {
      __syncthreads();
      for (int idx = tid; idx < kPqRows * kSLanes; idx += kPqThreads) {
        const int row = idx / kSLanes, j = idx % kSLanes;
        if (j < kSBands) {
          const bool ok = t0 + row < Fa;
          const int at = j * kPqLd + row + (row >> 3);
          sX[at] = ok ? X[(int64_t)(t0 + row) * kSLanes + j] : 0.f;
          sY[at] = ok ? Y[(int64_t)(t0 + row) * kSLanes + j] : 0.f;
        }
      }
      __syncthreads();
      float ax[4][kPqF], ay[4][kPqF];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int f = 0; f < kPqF; ++f) ax[i][f] = ay[i][f] = 0.f;
#pragma unroll 1
      for (int j = 0; j < kSBands; ++j) {
        // frames 8 tt + o, o = 0..21, live at 9 tt + o + o / 8 of the skewed row
        const float* xr = sX + j * kPqLd + 9 * tt;
        const float* yr = sY + j * kPqLd + 9 * tt;
        float xw[kPqF + kSStack - 1], yw[kPqF + kSStack - 1];
#pragma unroll
        for (int o = 0; o < kPqF + kSStack - 1; ++o) {
          xw[o] = xr[o + (o >> 3)];
          yw[o] = yr[o + (o >> 3)];
        }
#pragma unroll
        for (int k = 0; k < kSStack; ++k) {
          const int row = k * kSBands + j;
          const float4 uu = *reinterpret_cast<const float4*>(sU + row * kPqC + ((tc ^ (row & 7)) << 2));
          const float uc[4] = {uu.x, uu.y, uu.z, uu.w};
#pragma unroll
          for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int f = 0; f < kPqF; ++f) {
              ax[i][f] = fmaf(uc[i], xw[k + f], ax[i][f]);
              ay[i][f] = fmaf(uc[i], yw[k + f], ay[i][f]);
            }
        }
      }
      // weighted moments of the 8 frames of this thread, then of the 16 frame groups
      float wt[kPqF];
#pragma unroll
      for (int f = 0; f < kPqF; ++f) {
        const int t = t0 + kPqF * tt + f;
        wt[f] = (t < span.ne) ? (float)span.w(t) : 0.f;
      }
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        float p1x = 0.f, p1y = 0.f, pxx = 0.f, pyy = 0.f, pxy = 0.f;
#pragma unroll
        for (int f = 0; f < kPqF; ++f) {
          const float wx = wt[f] * ax[i][f], wy = wt[f] * ay[i][f];
          p1x += wx;
          p1y += wy;
          pxx = fmaf(wx, ax[i][f], pxx);
          pyy = fmaf(wy, ay[i][f], pyy);
          pxy = fmaf(wx, ay[i][f], pxy);
        }
        double d1x = (double)p1x, d1y = (double)p1y, dxx = (double)pxx, dyy = (double)pyy, dxy = (double)pxy;
#pragma unroll
        for (int o = 8; o > 0; o >>= 1) {
          d1x += __shfl_xor_sync(0xffffffffu, d1x, o);
          d1y += __shfl_xor_sync(0xffffffffu, d1y, o);
          dxx += __shfl_xor_sync(0xffffffffu, dxx, o);
          dyy += __shfl_xor_sync(0xffffffffu, dyy, o);
          dxy += __shfl_xor_sync(0xffffffffu, dxy, o);
        }
        m1x[i] += d1x;
        m1y[i] += d1y;
        mxx[i] += dxx;
        myy[i] += dyy;
        mxy[i] += dxy;
      }
    }
    if (tt == 0) {
      const double inv_nf = 1.0 / (double)Nf;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int c = c0 + 4 * tc + i;
        if (c >= r) continue;
        const double cxx = mxx[i] - m1x[i] * m1x[i] * inv_nf, cyy = myy[i] - m1y[i] * m1y[i] * inv_nf;
        const double cxy = mxy[i] - m1x[i] * m1y[i] * inv_nf;
        double rho = 0.0;
        if (cxx > 0.0 && cyy > 0.0) rho = cxy / sqrt(cxx * cyy);
        if (rho > 1.0) rho = 1.0;
        if (rho < -1.0) rho = -1.0;
        const double pr = 0.75 * rho;
        s_info[c] = -0.5 * log2(1.0 - pr * pr);
        lam_out[c] = s_nrm[4 * tc + i];
        rho_out[c] = (float)rho;
      }
    }
  }
  __syncthreads();
  for (int j = r + tid; j < kSDim; j += kPqThreads) {
    lam_out[j] = 0.f;
    rho_out[j] = 0.f;
  }
  if (tid == 0) {
    double info = 0.0;
    for (int c = 0; c < r; ++c) info += s_info[c];
    const double R = 1.0 / 200.0 * 16000.0;
    const double v = R / (double)kSStack * info;
    b.score[pair] = v > 0.0 ? v : 0.0;
    b.status[pair] = 0x100;  // periodic pair, rank <= 112: null space dropped (NELE_INFO_SIIB_NULLSPACE)
  }
}

// ------------------------------------------------------------- launchers
void siib_upload_tables(const float* win, const float* decay, const float* g2t, const float* tw, cudaStream_t s) {
  cudaMemcpyToSymbolAsync(g_siib_win, win, sizeof(float) * kSWin, 0, cudaMemcpyHostToDevice, s);
  cudaMemcpyToSymbolAsync(c_siib_decay, decay, sizeof(float) * kSMaskT, 0, cudaMemcpyHostToDevice, s);
  cudaMemcpyToSymbolAsync(g_siib_g2t, g2t, sizeof(float) * kSBins * kSLanes, 0, cudaMemcpyHostToDevice, s);
  cudaMemcpyToSymbolAsync(g_siib_tw, tw, sizeof(float) * 2 * kSWin, 0, cudaMemcpyHostToDevice, s);
  cudaFuncSetAttribute(siib_spec_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(SpecSmem));
  cudaFuncSetAttribute(siib_chol_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kCholW * kSDim * (int)sizeof(double));
  cudaFuncSetAttribute(siib_projquad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kPqSmem);
  cudaFuncSetAttribute(siib_jacobi2_kernel<1, kJ2WarpsSmall>, cudaFuncAttributeMaxDynamicSharedMemorySize, 2 * kJ2MaxSb * kSLd * (int)sizeof(float));
  cudaFuncSetAttribute(siib_jacobi2_kernel<4, kJ2Warps>, cudaFuncAttributeMaxDynamicSharedMemorySize, 2 * kJ2MaxSb * kSLd * (int)sizeof(float));
  siib_knn_setup();
  cudaStreamSynchronize(s);
}

int siib_run_wrapvad(const SiibGeom& g, const SiibBuffers& b, int n, bool no_tile, KernelTimer* kt, cudaStream_t s) {
  kt_begin(kt, "siib_wrapvad", s);
  siib_wrapvad_kernel<<<n, kVadThreads, 0, s>>>(g, b, no_tile ? 1 : 0);
  kt_end(kt, s);
  return 1;
}

int siib_run(const SiibGeom& g, const SiibBuffers& b_in, const SiibKnnBuffers* kb, const SiibEigBuffers* eb, int n, int64_t max_F,
             int64_t max_unique, KernelTimer* kt, cudaStream_t s) {
  int launches = 0;
  // FP32 xx lag products for the pairs that skip the Cholesky below (same condition: eb, x features not periodic), unless
  // an A/B run asks for the FP64 ones (NELE_COV_XX64=1) or for the FP64 tridiagonalisation (NELE_TRIDIAG_F64=1)
  static const bool xx64 = [] {
    const char* p = getenv("NELE_COV_XX64");
    const char* q = getenv("NELE_TRIDIAG_F64");
    return (p && p[0] == '1') || (q && q[0] == '1');
  }();
  SiibBuffers b = b_in;
  b.xx32 = (eb && !xx64) ? 1 : 0;
  kt_begin(kt, "siib_vad", s);
  siib_vad_kernel<<<n, kVad2Threads, 0, s>>>(g, b);
  kt_end(kt, s);
  ++launches;
  if (max_F > 0) {
    kt_begin(kt, "siib_spec", s);
    // active index t <= frame f, and frames beyond the first period are copies: t < max_unique is enough
    siib_spec_kernel<<<dim3((unsigned)((std::min(max_F, max_unique) + kSpecWarps * kSpecIter - 1) / (kSpecWarps * kSpecIter)), n), kSpecWarps * 32, sizeof(SpecSmem), s>>>(g, b);
    kt_end(kt, s);
    ++launches;
  }
  kt_begin(kt, "siib_mask", s);
  siib_mask_kernel<<<(2 * n + 3) / 4, 128, 0, s>>>(g, b, 2 * n);
  kt_end(kt, s);
  ++launches;
  kt_begin(kt, "siib_cov", s);
  siib_cov_kernel<<<dim3((kCovTasks64 + kCovWarps - 1) / kCovWarps, n), kCovWarps * 32, 0, s>>>(g, b);
  kt_end(kt, s);
  ++launches;
  kt_begin(kt, "siib_cov32", s);
  siib_cov32_kernel<<<dim3(b.xx32 ? 3 : 2, n), kCov32Warps * 32, 0, s>>>(g, b);
  kt_end(kt, s);
  ++launches;
  kt_begin(kt, "siib_expand", s);
  siib_expand_kernel<<<n, kExpThreads, 0, s>>>(g, b);
  kt_end(kt, s);
  ++launches;
  kt_begin(kt, "siib_chol", s);
  siib_chol_kernel<<<n, kCholThreads, kCholW * kSDim * sizeof(double), s>>>(g, b, eb ? 1 : 0);
  kt_end(kt, s);
  ++launches;
  {
    // r <= 112: one CTA per pair, everything in shared memory; larger ranks: 4-CTA clusters
    const size_t smem = (size_t)2 * kJ2MaxSb * kSLd * sizeof(float);
    if (eb) {
      launches += siib_run_small_eig(g, b, *eb, n, 2 * kJ2MaxSb, kt, s);
    } else {
      kt_begin(kt, "siib_jacobi", s);
      siib_jacobi2_kernel<1, kJ2WarpsSmall><<<n, kJ2WarpsSmall * 32, smem, s>>>(g, b, 2, 2 * kJ2MaxSb);
      kt_end(kt, s);
      ++launches;
    }
    if (eb) {
      launches += siib_run_eig(g, b, *eb, n, 2 * kJ2MaxSb + 1, kt, s);
    } else {
      cudaLaunchConfig_t cfg = {};
      cfg.gridDim = dim3(4 * n);
      cfg.blockDim = dim3(kJ2Warps * 32);
      cfg.dynamicSmemBytes = smem;
      cfg.stream = s;
      cudaLaunchAttribute attr[1];
      attr[0].id = cudaLaunchAttributeClusterDimension;
      attr[0].val.clusterDim.x = 4;
      attr[0].val.clusterDim.y = 1;
      attr[0].val.clusterDim.z = 1;
      cfg.attrs = attr;
      cfg.numAttrs = 1;
      kt_begin(kt, "siib_jacobi_cluster", s);
      cudaLaunchKernelEx(&cfg, siib_jacobi2_kernel<4, kJ2Warps>, g, b, 2 * kJ2MaxSb + 1, kSDim);
      kt_end(kt, s);
      ++launches;
    }
  }
  if (kb) return launches + siib_run_knn(g, b, *kb, n, max_F, kt, s);
  // register-tiled quadratic forms over the folded Sxy / Syy (siib_klt.cu)
  launches += siib_launch_quadform(b, b.info_part, n, kt, s);
  if (!b.no_proj) {
    kt_begin(kt, "siib_projquad", s);
    siib_projquad_kernel<<<n, kPqThreads, kPqSmem, s>>>(g, b);
    kt_end(kt, s);
    ++launches;
  }
  return launches;
}

}  // namespace nele
