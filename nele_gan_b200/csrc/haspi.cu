// HASPI v2 on sm_100a: batched ear model + modulation-correlation back-end.
//
// Data flow for one sub-batch of pairs (reference: pyHASPI/pyhaspi2.py:76-107):
//
//   haspi_prep_kernel      per (pair, signal) CTA: RMS normalise (:81-84), polyphase
//                          resampy kaiser_best to 24 kHz + RMS match (:810-821), middle
//                          ear IIR (:833-841) as a 256-way blocked linear recurrence
//   haspi_control_kernel   per (pair, signal) warp, lane = band: control-path gammatone
//                          envelope power over the whole utterance -> BWx / BWy (:1202-1205)
//   haspi_shift_kernel     group-delay shifts from BWx (:1115-1122)
//   haspi_ear_kernel       per (pair, signal) warp, lane = band: control + signal
//                          gammatone, compression, dB SL, IHC adaptation (:1207-1229), delay
//                          alignment (:1238-1240) and the 52-tap Hann FIR kept at every 9th
//                          sample (:378-414) -> envlp [nsub][32]
//   haspi_cep_kernel       per pair CTA: loudness threshold, compaction, dither, 32->5
//                          cosine projection, column means (:342-375)
//   haspi_modcorr_kernel   per (pair, cepstral coef, time tile) CTA: ten real zero-phase FIR
//                          band-pass filters (= ebm_ModFilt, :275-339) fused with the
//                          correlation sums of ebm_ModCorr (:254-273)
//   haspi_score_kernel     per pair thread: CM -> aveCM -> Intel (:102-104)
//
// The ear-model kernels run one auditory band per lane and walk the time axis
// sequentially: a batch of thousands of pairs supplies 64 independent
// recurrences per pair, which fills the machine without the >2x arithmetic
// overhead of a time-parallel scan of the 4th-order complex gammatone.
#include <stdio.h>

#include <stdlib.h>

#include <type_traits>

#include "host_tables.hpp"
#include "kernels.h"
#include "philox.cuh"

namespace nele {

__constant__ float c_envfir[54];
__constant__ IhcConst c_ihc;
__constant__ float c_cepm[kBands * kNumCep];
__constant__ int c_mod_nhalf[kNumMod];
__constant__ int c_mod_off[kNumMod + 1];
__device__ float g_modtaps[2880];

constexpr int kModTapsTotal = 2850;
constexpr int kModMaxHalf = 307;

// ------------------------------------------------------------------ prep
// Middle-ear filter state update, direct form II transposed exactly as
// scipy.signal.lfilter runs the two sections (pyhaspi2.py:835-840).
struct MidState {
  double z0, z1, z2;
};
__device__ __forceinline__ double mid_step(MidState& s, double x) {
  const double bl = 0.434173751206302, al = -0.131652497587396;
  const double b0 = 0.937260390269893, b1 = -1.874520780539785, b2 = 0.937260390269893;
  const double a1 = -1.870580640735279, a2 = 0.878460920344291;
  const double y1 = bl * x + s.z0;
  s.z0 = bl * x - al * y1;
  const double y2 = b0 * y1 + s.z1;
  s.z1 = b1 * y1 - a1 * y2 + s.z2;
  s.z2 = b2 * y1 - a2 * y2;
  return y2;
}

constexpr int kPrepThreads = 256;
constexpr int kPrepGroups = 85;                       // 3 threads x 4 outputs each = 12 outputs per group
constexpr int kPrepSpan = kPrepGroups * 8 + 128 + 16;  // staged inputs per tile

// F32 (default; NELE_RESAMPLE_F32=0 selects FP64): the 16 -> 24 kHz polyphase products and sums in FP32 instead of
// FP64 (the output is stored as float32 either way; 128 taps of FP32 accumulation leave ~1e-6 relative error,
// measured on a B200: HASPI moves by 8e-7, 12.4 -> 9.6 ms per 4096 x 3 s).  Everything else -- RMS, level match,
// middle ear -- stays FP64.
template <bool F32>
__global__ void __launch_bounds__(kPrepThreads, 4) haspi_prep_kernel(PairGeom g, HaspiBuffers b) {
  using RT = typename std::conditional<F32, float, double>::type;
  const int pair = blockIdx.x, q = blockIdx.y, tid = threadIdx.x;
  const float* __restrict__ src = (q == 0 ? b.ref : b.deg) + g.off16[pair];
  const int L = g.len16[pair];
  const int N = g.n24[pair];
  float* __restrict__ x24 = b.x24 + (int64_t)q * b.tot24 + g.off24[pair];
  float* __restrict__ mid = b.mid + (int64_t)q * b.tot24 + g.off24[pair];
  __shared__ double red[32];
  __shared__ double s_loc[kPrepThreads][3];
  __shared__ double s_M[3][3];

  // 1. RMS of the input (pyhaspi2.py:81-84)
  double ss = 0.0;
  for (int i = tid; i < L; i += kPrepThreads) {
    const double v = (double)src[i];
    ss += v * v;
  }
  ss = block_sum(ss, red);
  // all-zero (or non-finite) input: the reference divides by zero and ends in "Signal below
  // threshold"; scale by 0 instead so that every later stage stays finite and nsel comes out 0
  const double inv_rms = (ss > 0.0 && ss < 1.0e300) ? 1.0 / sqrt(ss / (double)L) : 0.0;

  // 2. resample to 24 kHz (resampy kaiser_best, phase-tabulated) or copy
  double ss24 = 0.0;
  if (b.rs_up == 1 && b.rs_down == 1) {
    for (int t = tid; t < N; t += kPrepThreads) {
      const float v = (float)((double)src[t] * inv_rms);
      x24[t] = v;
      ss24 += (double)v * (double)v;
    }
  } else if (b.rs_up == 3 && b.rs_down == 2) {
    // 16 -> 24 kHz fast path.  y[t] = sum_{m=-63..64} T[r][m] x[n + m], n = floor(2t/3), r = 2t mod 3.
    // A thread owns four outputs of one phase, t = 12 v + phi + 3 j (n advances by 2 per j), so each
    // tap is fetched once for four FP64 FMAs and the inputs slide through an 8-register window.
    // Inputs are staged in shared memory as FP64 with a (a + a/8) skew: lanes 8 samples apart hit
    // distinct banks.
    __shared__ RT s_tap[3][130];
    extern __shared__ double s_xin_raw[];  // skewed tile of inputs
    RT* s_xin = reinterpret_cast<RT*>(s_xin_raw);
    for (int k = tid; k < 3 * 128; k += kPrepThreads) {
      const int r = k / 128, m = k % 128;  // m = 0..127 <-> offset m - 63
      // taps[r][0..63] weigh x[n - i]; taps[r][64..127] weigh x[n + 1 + k]
      s_tap[r][m] = (RT)((m <= 63) ? b.rs_taps[r * 128 + (63 - m)] : b.rs_taps[r * 128 + 64 + (m - 64)]);
    }
    const int n_out = (int)(((int64_t)L * 3) / 2);  // int(L * ratio); the tail up to N is zero (fix_length)
    const int v_loc = tid / 3, phi = tid % 3;
    const bool worker = tid < kPrepGroups * 3;
    // the inputs of the next tile are fetched into registers while the current one is computed
    constexpr int kPre = (kPrepSpan + kPrepThreads - 1) / kPrepThreads;
    float pre[kPre];
    auto fetch = [&](int tile0) {
      const int in0 = (tile0 / 12) * 8 - 63;
#pragma unroll
      for (int c = 0; c < kPre; ++c) {
        const int jx = in0 + tid + c * kPrepThreads;
        pre[c] = (jx >= 0 && jx < L) ? __ldg(src + jx) : 0.f;
      }
    };
    fetch(0);
    for (int tile0 = 0; tile0 < N; tile0 += kPrepGroups * 12) {
      // inputs needed: n in [8 V0 - 63, 8 (V0 + groups) + 1 + 64 + 6], V0 = tile0 / 12
      const int in0 = (tile0 / 12) * 8 - 63;
      __syncthreads();
#pragma unroll
      for (int c = 0; c < kPre; ++c) {
        const int u = tid + c * kPrepThreads;
        if (u < kPrepSpan) s_xin[u + (u >> 3)] = (RT)pre[c];
      }
      __syncthreads();
      if (tile0 + kPrepGroups * 12 < N) fetch(tile0 + kPrepGroups * 12);
      if (worker) {
        const int t0 = tile0 + 12 * v_loc + phi;       // first of the four outputs
        const int n = (2 * t0) / 3, r = (2 * t0) % 3;  // t0 = 12 v + phi -> n = 8 v + {0, 0, 1}
        const int base = n - 63 - in0;                 // staged index of x[n - 63]
        const RT* __restrict__ tp = s_tap[r];
        RT a0 = 0, a1 = 0, a2 = 0, a3 = 0;
        RT w[8];
#pragma unroll
        for (int k = 0; k < 7; ++k) {
          const int u = base + k;
          w[k] = s_xin[u + (u >> 3)];
        }
#pragma unroll 1
        for (int m0 = 0; m0 < 128; m0 += 8) {
#pragma unroll
          for (int sft = 0; sft < 8; ++sft) {
            const int u = base + m0 + sft + 7;
            w[(sft + 7) & 7] = s_xin[u + (u >> 3)];
            const RT tm = tp[m0 + sft];
            a0 = fma(tm, w[sft & 7], a0);
            a1 = fma(tm, w[(sft + 2) & 7], a1);
            a2 = fma(tm, w[(sft + 4) & 7], a2);
            a3 = fma(tm, w[(sft + 6) & 7], a3);
          }
        }
        const double acc[4] = {(double)a0, (double)a1, (double)a2, (double)a3};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int t = t0 + 3 * j;
          if (t < N) {
            const float v = (t < n_out) ? (float)(acc[j] * inv_rms) : 0.f;
            x24[t] = v;
            ss24 += (double)v * (double)v;
          }
        }
      }
    }
  } else {
    const int n_out = (int)(((int64_t)L * b.rs_up) / b.rs_down);  // int(L * ratio)
    for (int t = tid; t < N; t += kPrepThreads) {
      float v = 0.f;
      if (t < n_out) {
        const int64_t pos = (int64_t)t * b.rs_down;
        const int n = (int)(pos / b.rs_up), r = (int)(pos % b.rs_up);
        const double* __restrict__ tp = b.rs_taps + (size_t)r * 128;
        double acc = 0.0;
        const int cl = min(n + 1, 64);
#pragma unroll 4
        for (int i = 0; i < cl; ++i) acc = fma(__ldg(tp + i), (double)__ldg(src + n - i), acc);
        const int cr = min(L - n - 1, 64);
#pragma unroll 4
        for (int k = 0; k < cr; ++k) acc = fma(__ldg(tp + 64 + k), (double)__ldg(src + n + 1 + k), acc);
        v = (float)(acc * inv_rms);
      }
      x24[t] = v;
      ss24 += (double)v * (double)v;
    }
  }
  ss24 = block_sum(ss24, red);
  // (xRMS / yRMS) * y with xRMS = 1 after the normalisation above (pyhaspi2.py:816-818)
  const double scale = (b.rs_up == 1 && b.rs_down == 1 || !(ss24 > 0.0)) ? 1.0 : 1.0 / sqrt(ss24 / (double)N);
  __syncthreads();

  // 3. middle ear as a blocked linear recurrence: each thread filters one
  //    contiguous chunk from a zero state, the chunk-to-chunk carries are
  //    propagated with the 3x3 zero-input transition matrix, then every chunk
  //    is re-run from its exact initial state.
  const int Lc = (N + kPrepThreads - 1) / kPrepThreads;
  const int t0 = min(tid * Lc, N), t1 = min(t0 + Lc, N);
  // A thread's chunk is contiguous, so lanes are Lc samples apart: the samples go through a
  // per-warp shared-memory tile, eight per lane and round, moved with 32-byte row segments (four
  // chunks per warp instruction) instead of one sector per lane and instruction.
  __shared__ float s_tx[kPrepThreads / 32][32][9];
  __shared__ float s_tm[kPrepThreads / 32][32][9];
  const int lane = tid & 31, wib = tid >> 5;
  float (*tx)[9] = s_tx[wib];
  float (*tm)[9] = s_tm[wib];
  const int rounds = (Lc + 7) / 8;
  const int crow = lane >> 3, ccol = lane & 7;   // cooperative moves: rows 4 i + crow, column ccol
  MidState st = {0.0, 0.0, 0.0};
  for (int r = 0; r < rounds; ++r) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int row = 4 * i + crow;
      const int r0 = min((wib * 32 + row) * Lc, N), r1 = min(r0 + Lc, N);
      const int t = r0 + 8 * r + ccol;
      tx[row][ccol] = (t < r1) ? x24[t] : 0.f;
    }
    __syncwarp();
#pragma unroll
    for (int i = 0; i < 8; ++i)
      if (t0 + 8 * r + i < t1) mid_step(st, (double)(float)((double)tx[lane][i] * scale));
    __syncwarp();
  }
  s_loc[tid][0] = st.z0;
  s_loc[tid][1] = st.z1;
  s_loc[tid][2] = st.z2;
  if (tid < 3) {
    MidState u = {tid == 0 ? 1.0 : 0.0, tid == 1 ? 1.0 : 0.0, tid == 2 ? 1.0 : 0.0};
    for (int k = 0; k < Lc; ++k) mid_step(u, 0.0);
    s_M[0][tid] = u.z0;
    s_M[1][tid] = u.z1;
    s_M[2][tid] = u.z2;
  }
  __syncthreads();
  if (tid == 0) {
    double c0 = 0.0, c1 = 0.0, c2 = 0.0;  // state entering chunk k
    for (int k = 0; k < kPrepThreads; ++k) {
      const double l0 = s_loc[k][0], l1 = s_loc[k][1], l2 = s_loc[k][2];
      s_loc[k][0] = c0;
      s_loc[k][1] = c1;
      s_loc[k][2] = c2;
      const double n0 = s_M[0][0] * c0 + s_M[0][1] * c1 + s_M[0][2] * c2 + l0;
      const double n1 = s_M[1][0] * c0 + s_M[1][1] * c1 + s_M[1][2] * c2 + l1;
      const double n2 = s_M[2][0] * c0 + s_M[2][1] * c1 + s_M[2][2] * c2 + l2;
      c0 = n0;
      c1 = n1;
      c2 = n2;
    }
  }
  __syncthreads();
  st.z0 = s_loc[tid][0];
  st.z1 = s_loc[tid][1];
  st.z2 = s_loc[tid][2];
  for (int r = 0; r < rounds; ++r) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int row = 4 * i + crow;
      const int r0 = min((wib * 32 + row) * Lc, N), r1 = min(r0 + Lc, N);
      const int t = r0 + 8 * r + ccol;
      tx[row][ccol] = (t < r1) ? x24[t] : 0.f;
    }
    __syncwarp();
#pragma unroll
    for (int i = 0; i < 8; ++i)
      if (t0 + 8 * r + i < t1) {
        const float v = (float)((double)tx[lane][i] * scale);
        tm[lane][i] = (float)mid_step(st, (double)v);
      }
    __syncwarp();
    // the middle-ear output leaves as float32: the control and main passes run their recurrences in FP32 and
    // converted it on load anyway (round 1 kept it as FP64 in HBM: 1.15 MB per 3 s pair written once, read twice --
    // the largest avoidable stream of the step); the level-matched x24 is not written back (nobody reads it)
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int row = 4 * i + crow;
      const int r0 = min((wib * 32 + row) * Lc, N), r1 = min(r0 + Lc, N);
      const int t = r0 + 8 * r + ccol;
      if (t < r1) mid[t] = tm[row][ccol];
    }
    __syncwarp();
  }
}

// --------------------------------------------------------------- control
constexpr int kEarWarps = 4;          // warps per CTA in the lane = band kernels
constexpr int kEarChunkMax = 576;     // delay-line capacity of the main pass (kEarChunk)
constexpr int kCtlChunk = 256;

// One warp per pair, lane = band; the clean and the processed signal run side by side in a thread
// (two independent recurrence chains, one shared carrier), as in the main pass.
// (Occupancy is not the limit: capping the registers so that 7 instead of 6 CTAs fit per SM -- one wave instead of two at
// 4096 pairs -- measured 13.8 -> 14.5 ms; the FP32 pipe is saturated either way.)
template <typename T>
__global__ void __launch_bounds__(kEarWarps * 32) haspi_control_kernel(PairGeom g, HaspiBuffers b, int n_pairs) {
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const int pair = blockIdx.x * kEarWarps + wib;
  if (pair >= n_pairs) return;
  const float* __restrict__ midx = b.mid + g.off24[pair];
  const float* __restrict__ midy = b.mid + b.tot24 + g.off24[pair];
  const int N = g.n24[pair];
  __shared__ T s_buf[kEarWarps][2][kCtlChunk];
  T* bx = s_buf[wib][0];
  T* by = s_buf[wib][1];
  const BandConst bc = b.bands[lane];
  Carrier<T> car;
  car.init(bc.cf);
  const GtCoef<T> k = make_gt<T>(bc.bw1, bc.erb);
  Gt4<T> fx, fy;
  fx.reset();
  fy.reset();
  double accx = 0.0, accy = 0.0;
  for (int base = 0; base < N; base += kCtlChunk) {
    __syncwarp();
#pragma unroll
    for (int q = 0; q < kCtlChunk / 32; ++q) {
      const int t = base + q * 32 + lane;
      bx[q * 32 + lane] = (t < N) ? (T)midx[t] : (T)0;
      by[q * 32 + lane] = (t < N) ? (T)midy[t] : (T)0;
    }
    __syncwarp();
    car.seed_before(base);
    const int m = min(kCtlChunk, N - base);
    T px = (T)0, py = (T)0;
#pragma unroll 4
    for (int q = 0; q < m; ++q) {
      car.advance();
      const T xs = bx[q], ys = by[q];
      px += fx.step(k, xs * car.c, xs * car.s);
      py += fy.step(k, ys * car.c, ys * car.s);
    }
    accx += (double)px;
    accy += (double)py;
  }
  const int64_t o = (int64_t)pair * 2 * kBands + lane;
  b.bw[o] = bw_from_control(accx, (double)k.gain, N, bc.bwmin[0], bc.bw1);
  b.bw[o + kBands] = bw_from_control(accy, (double)k.gain, N, bc.bwmin[1], bc.bw1);
  if (b.cave) {
    b.cave[o] = (double)k.gain * sqrt(accx / (double)N);
    b.cave[o + kBands] = (double)k.gain * sqrt(accy / (double)N);
  }
}

// group-delay shifts, always from BWx (pyhaspi2.py:1239-1240)
__global__ void haspi_shift_kernel(HaspiBuffers b, int n) {
  const int pair = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (pair >= n) return;
  const double gd = gt_group_delay(b.bw[((int64_t)pair * 2 + 0) * kBands + lane], b.bands[lane].erb);
  const double gmax = warp_max(gd);
  const double sh = gmax - gd;  // 0 .. ~435 samples for BW in [1, 4]
  b.shift[(int64_t)pair * kBands + lane] = (sh >= 0.0 && sh < (double)kEarChunkMax) ? (int)sh : 0;
}

// ------------------------------------------------------------------ main
constexpr int kEarChunk = 576;  // 64 blocks of 9 samples; ring of two chunks per warp and signal

// One warp per pair, lane = band; each lane runs the clean and the processed signal
// side by side (two independent recurrence chains per thread: the sample loop is
// latency bound, and the carrier, the delay-line index and the loop bookkeeping are
// shared because both signals use the shifts of BWx, pyhaspi2.py:1239-1240).
template <typename T>
__global__ void __launch_bounds__(kEarWarps * 32) haspi_ear_kernel(PairGeom g, HaspiBuffers b, int n_pairs) {
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const int pair = blockIdx.x * kEarWarps + wib;
  if (pair >= n_pairs) return;
  const float* __restrict__ midx = b.mid + g.off24[pair];
  const float* __restrict__ midy = b.mid + b.tot24 + g.off24[pair];
  const int N = g.n24[pair], nsub = g.nsub[pair];
  float* __restrict__ outx = b.envlp + (g.offsub[pair]) * kBands;
  float* __restrict__ outy = b.envlp + (b.totsub + g.offsub[pair]) * kBands;
  extern __shared__ __align__(16) unsigned char s_ring_raw[];  // [kEarWarps][2][2 * kEarChunk] of T
  T* ringx = reinterpret_cast<T*>(s_ring_raw) + (size_t)wib * 4 * kEarChunk;
  T* ringy = ringx + 2 * kEarChunk;

  EarLane<T> Lx, Ly;
  Carrier<T> car;
  int shift;
  {
    const BandConst bc = b.bands[lane];
    shift = b.shift[(int64_t)pair * kBands + lane];
    car.init(bc.cf);
    Lx.init(bc, 0, b.bw[((int64_t)pair * 2 + 0) * kBands + lane], c_ihc);
    Ly.init(bc, 1, b.bw[((int64_t)pair * 2 + 1) * kBands + lane], c_ihc);
  }
  int wshift = shift;  // largest delay of the warp (band 31, up to ~435 samples)
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) wshift = max(wshift, __shfl_xor_sync(0xffffffffu, wshift, o));
  int rp = (shift == 0) ? 0 : 2 * kEarChunk - shift;  // ring slot of sample i - shift
  const int nblk = nsub + 2;                           // output j completes after block j + 2
  const int nchunks = (nblk * 9 + kEarChunk - 1) / kEarChunk;
  for (int c = 0; c < nchunks; ++c) {
    const int i0 = c * kEarChunk;
    T* hx = ringx + (c & 1) * kEarChunk;
    T* hy = ringy + (c & 1) * kEarChunk;
    __syncwarp();
#pragma unroll
    for (int k = 0; k < kEarChunk / 32; ++k) {
      const int t = i0 + k * 32 + lane;
      hx[k * 32 + lane] = (t < N) ? (T)midx[t] : (T)0;
      hy[k * 32 + lane] = (t < N) ? (T)midy[t] : (T)0;
    }
    __syncwarp();
    {  // re-seed the carrier from the exact phase (lane-local time axis)
      const int t = i0 - shift;
      if (t >= 0) car.seed_before(t);
      else if (t + kEarChunk > 0) car.seed_before(0);
    }
    const int blk_end = min((c + 1) * (kEarChunk / 9), nblk);
    for (int blk = c * (kEarChunk / 9); blk < blk_end; ++blk) {
      const int ib = blk * 9;
      float vx[9], vy[9];
      if (ib >= wshift && ib + 9 <= N) {
        // steady state (all but the first ~50 and the last block): every lane is past its own
        // delay and inside the signal, so the nine samples need no per-sample predicates
#pragma unroll
        for (int p = 0; p < 9; ++p) {
          const T xs = ringx[rp], ys = ringy[rp];
          rp = (rp + 1 == 2 * kEarChunk) ? 0 : rp + 1;
          car.advance();
          vx[p] = Lx.sample(xs * car.c, xs * car.s);
          vy[p] = Ly.sample(ys * car.c, ys * car.s);
        }
      } else {
#pragma unroll
        for (int p = 0; p < 9; ++p) {
          const int i = ib + p;
          const T xs = ringx[rp], ys = ringy[rp];
          rp = (rp + 1 == 2 * kEarChunk) ? 0 : rp + 1;
          if (i >= shift && i < N) {
            car.advance();
            vx[p] = Lx.sample(xs * car.c, xs * car.s);
            vy[p] = Ly.sample(ys * car.c, ys * car.s);
          } else {
            vx[p] = 0.f;
            vy[p] = 0.f;
          }
        }
      }
      Lx.template accumulate<0>(vx[0], c_envfir); Ly.template accumulate<0>(vy[0], c_envfir);
      Lx.template accumulate<1>(vx[1], c_envfir); Ly.template accumulate<1>(vy[1], c_envfir);
      Lx.template accumulate<2>(vx[2], c_envfir); Ly.template accumulate<2>(vy[2], c_envfir);
      Lx.template accumulate<3>(vx[3], c_envfir); Ly.template accumulate<3>(vy[3], c_envfir);
      Lx.template accumulate<4>(vx[4], c_envfir); Ly.template accumulate<4>(vy[4], c_envfir);
      Lx.template accumulate<5>(vx[5], c_envfir); Ly.template accumulate<5>(vy[5], c_envfir);
      Lx.template accumulate<6>(vx[6], c_envfir); Ly.template accumulate<6>(vy[6], c_envfir);
      Lx.template accumulate<7>(vx[7], c_envfir); Ly.template accumulate<7>(vy[7], c_envfir);
      Lx.template accumulate<8>(vx[8], c_envfir); Ly.template accumulate<8>(vy[8], c_envfir);
      const float ox = Lx.emit(), oy = Ly.emit();
      const int j = blk - 2;
      if (j >= 0) {
        outx[(int64_t)j * kBands + lane] = ox;
        outy[(int64_t)j * kBands + lane] = oy;
      }
    }
  }
}

// ---- the main pass with the clean and the processed chain packed into fma.rn.f32x2 (Blackwell's
// two-wide FP32 FMA: the same FMA-pipe throughput as two FFMA, scripts/micro/ffma2.cu, but one issue
// slot).  Every linear recurrence, the level arithmetic around the MUFU calls and the FIR
// bookkeeping run two-wide; only lg2 / ex2 and the clamps stay scalar: ~130 -> ~75 issue slots per
// sample pair at the same FMA-pipe time.  Measured on a B200 (round 2): 45.1 -> 40.6 ms per
// 4096 x 3 s, HASPI equal to 2.7e-6; default since then (NELE_F32X2=0 selects the scalar kernel).
// The control pass was packed the same way and was *slower* (4.15 -> 4.65 ms per 1024 pairs: it
// has no MUFU / clamp work to overlap, so packing only lengthens the dependency chain): removed.
// (168 registers, three CTAs per SM.  A cap at 128 registers -- four CTAs, two waves instead of three at 4096 pairs, 16 bytes
// of spills -- measured 40.7 -> 41.9 ms: the kernel is bound by the FP32 pipe, not by occupancy or the tail of the grid.)
__global__ void __launch_bounds__(kEarWarps * 32) haspi_ear_x2_kernel(PairGeom g, HaspiBuffers b, int n_pairs) {
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const int pair = blockIdx.x * kEarWarps + wib;
  if (pair >= n_pairs) return;
  const float* __restrict__ midx = b.mid + g.off24[pair];
  const float* __restrict__ midy = b.mid + b.tot24 + g.off24[pair];
  const int N = g.n24[pair], nsub = g.nsub[pair];
  float* __restrict__ outx = b.envlp + (g.offsub[pair]) * kBands;
  float* __restrict__ outy = b.envlp + (b.totsub + g.offsub[pair]) * kBands;
  extern __shared__ __align__(16) unsigned char s_ring_raw[];  // [kEarWarps][2 * kEarChunk] sample pairs (x, y)
  float2* ring = reinterpret_cast<float2*>(s_ring_raw) + (size_t)wib * 2 * kEarChunk;

  EarLane2 L2;
  Carrier<float> car;
  int shift;
  {
    const BandConst bc = b.bands[lane];
    shift = b.shift[(int64_t)pair * kBands + lane];
    car.init(bc.cf);
    L2.init(bc, b.bw[((int64_t)pair * 2 + 0) * kBands + lane], b.bw[((int64_t)pair * 2 + 1) * kBands + lane], c_ihc);
  }
  int wshift = shift;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) wshift = max(wshift, __shfl_xor_sync(0xffffffffu, wshift, o));
  int rp = (shift == 0) ? 0 : 2 * kEarChunk - shift;
  const int nblk = nsub + 2;
  const int nchunks = (nblk * 9 + kEarChunk - 1) / kEarChunk;
  const F2 zero = f2_pack(0.f, 0.f);
  for (int c = 0; c < nchunks; ++c) {
    const int i0 = c * kEarChunk;
    float2* h = ring + (c & 1) * kEarChunk;
    __syncwarp();
#pragma unroll
    for (int k = 0; k < kEarChunk / 32; ++k) {
      const int t = i0 + k * 32 + lane;
      h[k * 32 + lane] = (t < N) ? make_float2((float)midx[t], (float)midy[t]) : make_float2(0.f, 0.f);
    }
    __syncwarp();
    {
      const int t = i0 - shift;
      if (t >= 0) car.seed_before(t);
      else if (t + kEarChunk > 0) car.seed_before(0);
    }
    const int blk_end = min((c + 1) * (kEarChunk / 9), nblk);
    for (int blk = c * (kEarChunk / 9); blk < blk_end; ++blk) {
      const int ib = blk * 9;
      F2 v[9];
      if (ib >= wshift && ib + 9 <= N) {
#pragma unroll
        for (int p = 0; p < 9; ++p) {
          const float2 xy = ring[rp];
          rp = (rp + 1 == 2 * kEarChunk) ? 0 : rp + 1;
          car.advance();
          const F2 XY = f2_pack(xy.x, xy.y);
          v[p] = L2.sample(f2_mul(XY, f2_pack(car.c, car.c)), f2_mul(XY, f2_pack(car.s, car.s)));
        }
      } else {
#pragma unroll
        for (int p = 0; p < 9; ++p) {
          const int i = ib + p;
          const float2 xy = ring[rp];
          rp = (rp + 1 == 2 * kEarChunk) ? 0 : rp + 1;
          if (i >= shift && i < N) {
            car.advance();
            const F2 XY = f2_pack(xy.x, xy.y);
            v[p] = L2.sample(f2_mul(XY, f2_pack(car.c, car.c)), f2_mul(XY, f2_pack(car.s, car.s)));
          } else {
            v[p] = zero;
          }
        }
      }
      L2.accumulate<0>(v[0], c_envfir);
      L2.accumulate<1>(v[1], c_envfir);
      L2.accumulate<2>(v[2], c_envfir);
      L2.accumulate<3>(v[3], c_envfir);
      L2.accumulate<4>(v[4], c_envfir);
      L2.accumulate<5>(v[5], c_envfir);
      L2.accumulate<6>(v[6], c_envfir);
      L2.accumulate<7>(v[7], c_envfir);
      L2.accumulate<8>(v[8], c_envfir);
      float ox, oy;
      f2_unpack(L2.emit(), ox, oy);
      const int j = blk - 2;
      if (j >= 0) {
        outx[(int64_t)j * kBands + lane] = ox;
        outy[(int64_t)j * kBands + lane] = oy;
      }
    }
  }
}

// ------------------------------------------------------------- cepstra
constexpr int kCepThreads = 256;

// One thread per envelope row (the 32 band values of a row are one 128-byte line): the loudness
// test costs 32 EX2 + adds per row instead of a warp reduction, and the projection keeps its
// 2 x 2 x 5 accumulators in registers with the basis as immediate constant operands.  A thread
// projects two consecutive kept rows, so that one Philox block (four normals) dithers one band
// of both rows and both signals.
__global__ void __launch_bounds__(kCepThreads) haspi_cep_kernel(PairGeom g, HaspiBuffers b) {
  const int pair = blockIdx.x, tid = threadIdx.x, lane = tid & 31, wib = tid >> 5;
  constexpr int NW = kCepThreads / 32;
  const int nsub = g.nsub[pair];
  const int64_t rbase = g.offsub[pair];
  const float* __restrict__ ex = b.envlp + rbase * kBands;
  const float* __restrict__ ey = b.envlp + (b.totsub + rbase) * kBands;
  int32_t* __restrict__ keep = b.rowsel + rbase;  // compacted list of the kept rows
  __shared__ int s_cnt[NW];
  __shared__ int s_base;
  __shared__ double s_sum[NW][2 * kNumCep];
  if (tid == 0) s_base = 0;
  __syncthreads();
  // 1. loudness of every reference frame (pyhaspi2.py:352-355) + ordered compaction
  for (int r0 = 0; r0 < nsub; r0 += kCepThreads) {
    const int r = r0 + tid;
    int f = 0;
    if (r < nsub) {
      const float4* row = reinterpret_cast<const float4*>(ex + (int64_t)r * kBands);
      float lin = 0.f;
#pragma unroll
      for (int q4 = 0; q4 < kBands / 4; ++q4) {
        const float4 v = row[q4];
        lin += undb20(v.x) + undb20(v.y) + undb20(v.z) + undb20(v.w);
      }
      f = (db20(lin * (1.0f / kBands)) > 2.5f) ? 1 : 0;
    }
    int inc = f;
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
    if (f) keep[base + woff + inc - 1] = r;
    __syncthreads();
    if (tid == 0) s_base = base + tot;
    __syncthreads();
  }
  const int nsel = s_base;
  if (tid == 0) b.nsel[pair] = nsel;

  // 2. dither + projection of the kept frames (pyhaspi2.py:359-367), column sums
  float sum[2 * kNumCep];
#pragma unroll
  for (int j = 0; j < 2 * kNumCep; ++j) sum[j] = 0.f;
  const uint64_t gp = (uint64_t)(b.pair_base + pair);
  const int mode = b.no_dither ? 0 : (b.dither ? 1 : 2);
  for (int k = tid; 2 * k < nsel; k += kCepThreads) {
    const int ca = 2 * k, cb = 2 * k + 1;
    const bool hb = cb < nsel;
    const int ra = keep[ca], rb = hb ? keep[cb] : ra;
    const float4* xa = reinterpret_cast<const float4*>(ex + (int64_t)ra * kBands);
    const float4* ya = reinterpret_cast<const float4*>(ey + (int64_t)ra * kBands);
    const float4* xb = reinterpret_cast<const float4*>(ex + (int64_t)rb * kBands);
    const float4* yb = reinterpret_cast<const float4*>(ey + (int64_t)rb * kBands);
    float pxa[kNumCep], pya[kNumCep], pxb[kNumCep], pyb[kNumCep];
#pragma unroll
    for (int j = 0; j < kNumCep; ++j) pxa[j] = pya[j] = pxb[j] = pyb[j] = 0.f;
#pragma unroll
    for (int q4 = 0; q4 < kBands / 4; ++q4) {
      const float4 vxa = xa[q4], vya = ya[q4], vxb = xb[q4], vyb = yb[q4];
      float exa[4] = {vxa.x, vxa.y, vxa.z, vxa.w}, eya[4] = {vya.x, vya.y, vya.z, vya.w};
      float exb[4] = {vxb.x, vxb.y, vxb.z, vxb.w}, eyb[4] = {vyb.x, vyb.y, vyb.z, vyb.w};
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int band = 4 * q4 + u;
        if (mode == 1) {
          if (ca < b.dither_rows) {
            exa[u] += 0.1f * b.dither[(int64_t)ca * kBands + band];
            eya[u] += 0.1f * b.dither[(b.dither_rows + ca) * kBands + band];
          }
          if (cb < b.dither_rows) {
            exb[u] += 0.1f * b.dither[(int64_t)cb * kBands + band];
            eyb[u] += 0.1f * b.dither[(b.dither_rows + cb) * kBands + band];
          }
        } else if (mode == 2) {
          float z[4];
          philox_normal4(b.seed, gp, (uint32_t)k, (uint32_t)band, z);
          exa[u] = fmaf(0.1f, z[0], exa[u]);
          eya[u] = fmaf(0.1f, z[1], eya[u]);
          exb[u] = fmaf(0.1f, z[2], exb[u]);
          eyb[u] = fmaf(0.1f, z[3], eyb[u]);
        }
#pragma unroll
        for (int j = 0; j < kNumCep; ++j) {
          const float cm = c_cepm[band * kNumCep + j];
          pxa[j] = fmaf(exa[u], cm, pxa[j]);
          pya[j] = fmaf(eya[u], cm, pya[j]);
          pxb[j] = fmaf(exb[u], cm, pxb[j]);
          pyb[j] = fmaf(eyb[u], cm, pyb[j]);
        }
      }
    }
#pragma unroll
    for (int j = 0; j < kNumCep; ++j) {
      float* cx = b.cep + (int64_t)(0 * kNumCep + j) * b.totsub + rbase;
      float* cy = b.cep + (int64_t)(1 * kNumCep + j) * b.totsub + rbase;
      cx[ca] = pxa[j];
      cy[ca] = pya[j];
      sum[j] += pxa[j];
      sum[kNumCep + j] += pya[j];
      if (hb) {
        cx[cb] = pxb[j];
        cy[cb] = pyb[j];
        sum[j] += pxb[j];
        sum[kNumCep + j] += pyb[j];
      }
    }
  }
#pragma unroll
  for (int j = 0; j < 2 * kNumCep; ++j) {
    const double t = warp_sum((double)sum[j]);
    if (lane == 0) s_sum[wib][j] = t;
  }
  __syncthreads();
  if (tid < 2 * kNumCep) {
    double t = 0.0;
    for (int w = 0; w < NW; ++w) t += s_sum[w][tid];
    b.cepmean[(int64_t)pair * 2 * kNumCep + tid] = (nsel > 0) ? t / (double)nsel : 0.0;
  }
}

// ----------------------------------------------- modulation filter + corr
constexpr int kModThreads = 256;
constexpr int kModPer = 8;                                // consecutive outputs per thread
constexpr int kModTile = kModThreads * kModPer;           // 2048 outputs per CTA
constexpr int kModFront = kModMaxHalf + 1;               // staged samples before the tile
constexpr int kModSpan = kModTile + 2 * kModMaxHalf + 8;  // staged samples per signal
__host__ __device__ constexpr int mod_skew(int u) { return u + (u >> 3); }  // bank-conflict-free for stride-8 lanes

__global__ void __launch_bounds__(kModThreads) haspi_modcorr_kernel(PairGeom g, HaspiBuffers b) {
  const int pair = blockIdx.x, j = blockIdx.y, tile = blockIdx.z, tid = threadIdx.x;
  const int n = b.nsel[pair];
  const int tbase = tile * kModTile;
  if (tbase >= n || n <= 1) return;
  extern __shared__ float smem[];
  float* sx = smem;
  float* sy = smem + mod_skew(kModSpan) + 1;
  float* staps = sy + mod_skew(kModSpan) + 1;
  __shared__ double red[32];
  const int64_t rbase = g.offsub[pair];
  const float* __restrict__ cx = b.cep + (int64_t)(0 * kNumCep + j) * b.totsub + rbase;
  const float* __restrict__ cy = b.cep + (int64_t)(1 * kNumCep + j) * b.totsub + rbase;
  const float mx = (float)b.cepmean[(int64_t)pair * 2 * kNumCep + j];
  const float my = (float)b.cepmean[(int64_t)pair * 2 * kNumCep + kNumCep + j];
  // staged index u <-> sample tbase - kModFront + u, zero outside [0, n)
  for (int u = tid; u < kModSpan; u += kModThreads) {
    const int t = tbase - kModFront + u;
    const bool ok = (t >= 0 && t < n);
    sx[mod_skew(u)] = ok ? cx[t] - mx : 0.f;
    sy[mod_skew(u)] = ok ? cy[t] - my : 0.f;
  }
  for (int k = tid; k < kModTapsTotal; k += kModThreads) staps[k] = g_modtaps[k];
  __syncthreads();

  const int tl = tid * kModPer;  // first local output of this thread
  double* dst = b.modsum + (((int64_t)pair * kNumCep + j) * kNumMod) * 5;
  for (int m = 0; m < kNumMod; ++m) {
    const int nh = c_mod_nhalf[m];
    const float* __restrict__ tp = staps + c_mod_off[m];
    float ax[kModPer], ay[kModPer], wx[kModPer], wy[kModPer];
#pragma unroll
    for (int r = 0; r < kModPer; ++r) ax[r] = ay[r] = 0.f;
    // out[t] = sum_k g[k] x[t + nh - k]; window w[r] = x[tl + r + nh - k]
    const int u0 = tl + kModFront + nh;  // staged index of x[tl + nh] (k = 0, r = 0)
#pragma unroll
    for (int r = 0; r < kModPer; ++r) {
      wx[r] = sx[mod_skew(u0 + r)];
      wy[r] = sy[mod_skew(u0 + r)];
    }
    const int ntap = 2 * nh + 1;
    int k = 0;
    for (; k + kModPer <= ntap; k += kModPer) {
#pragma unroll
      for (int kk = 0; kk < kModPer; ++kk) {
        const float gk = tp[k + kk];
        // logical window after kk shifts: w[r] = x[.. + r - kk]; element (r - kk) mod P holds it
#pragma unroll
        for (int r = 0; r < kModPer; ++r) {
          ax[r] = fmaf(gk, wx[(r - kk + kModPer) % kModPer], ax[r]);
          ay[r] = fmaf(gk, wy[(r - kk + kModPer) % kModPer], ay[r]);
        }
        // slide: logical w[0] for the next tap is x[u0 - (k + kk) - 1]; it replaces the
        // slot that held logical w[P-1], i.e. physical (P - 1 - kk) mod P
        const int un = u0 - (k + kk) - 1;
        wx[(kModPer - 1 - kk + kModPer) % kModPer] = sx[mod_skew(un)];
        wy[(kModPer - 1 - kk + kModPer) % kModPer] = sy[mod_skew(un)];
      }
    }
    // remainder taps (ntap is odd): physical layout is back to identity here
    for (; k < ntap; ++k) {
      const float gk = tp[k];
#pragma unroll
      for (int r = 0; r < kModPer; ++r) {
        ax[r] = fmaf(gk, wx[r], ax[r]);
        ay[r] = fmaf(gk, wy[r], ay[r]);
      }
#pragma unroll
      for (int r = kModPer - 1; r > 0; --r) {
        wx[r] = wx[r - 1];
        wy[r] = wy[r - 1];
      }
      const int un = u0 - k - 1;
      wx[0] = sx[mod_skew(un)];
      wy[0] = sy[mod_skew(un)];
    }
    double s0 = 0, s1 = 0, s2 = 0, s3 = 0, s4 = 0;
#pragma unroll
    for (int r = 0; r < kModPer; ++r) {
      if (tbase + tl + r < n) {
        const double vx = ax[r], vy = ay[r];
        s0 += vx; s1 += vy; s2 += vx * vx; s3 += vy * vy; s4 += vx * vy;
      }
    }
    s0 = block_sum(s0, red);
    s1 = block_sum(s1, red);
    s2 = block_sum(s2, red);
    s3 = block_sum(s3, red);
    s4 = block_sum(s4, red);
    if (tid == 0) {
      atomicAdd(dst + m * 5 + 0, s0);
      atomicAdd(dst + m * 5 + 1, s1);
      atomicAdd(dst + m * 5 + 2, s2);
      atomicAdd(dst + m * 5 + 3, s3);
      atomicAdd(dst + m * 5 + 4, s4);
    }
  }
}

// ------------------------------------- modulation filter + corr, recursive form
// One thread per (cepstral coefficient, modulation band) of a pair, both signals: the three
// sliding complex sums of host::make_mod_recursions per signal, the band output, and the five
// correlation sums, all in FP64 registers; no atomics, no shared memory.  Replaces the direct-form
// kernel above (kept behind NELE_MODCORR_FIR=1 for A/B checks).
__device__ host::ModRecBand g_modrec[kNumMod];  // read once per thread into registers (lane-indexed
                                                // __constant__ reads would serialise on every step)

constexpr int kMod2Threads = 64;  // 50 used: 5 coefficients x 10 bands

__global__ void __launch_bounds__(kMod2Threads) haspi_modcorr2_kernel(PairGeom g, HaspiBuffers b) {
  const int pair = blockIdx.x, tid = threadIdx.x;
  const int n = b.nsel[pair];
  if (tid >= kNumCep * kNumMod || n <= 1) return;
  const int j = tid / kNumMod, m = tid % kNumMod;
  const int64_t rbase = g.offsub[pair];
  const float* __restrict__ cx = b.cep + (int64_t)(0 * kNumCep + j) * b.totsub + rbase;
  const float* __restrict__ cy = b.cep + (int64_t)(1 * kNumCep + j) * b.totsub + rbase;
  const double mx = b.cepmean[(int64_t)pair * 2 * kNumCep + j];
  const double my = b.cepmean[(int64_t)pair * 2 * kNumCep + kNumCep + j];
  double rr[3], ri[3], ar[3], ai[3], br[3], bi[3], wg[3];
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    rr[k] = g_modrec[m].rot[k][0];
    ri[k] = g_modrec[m].rot[k][1];
    ar[k] = g_modrec[m].cin[k][0];
    ai[k] = g_modrec[m].cin[k][1];
    br[k] = -g_modrec[m].cout[k][0];
    bi[k] = -g_modrec[m].cout[k][1];
    wg[k] = g_modrec[m].wgt[k];
  }
  const int nh = g_modrec[m].nh;
  double xr[3] = {0, 0, 0}, xi[3] = {0, 0, 0}, yr[3] = {0, 0, 0}, yi[3] = {0, 0, 0};
  double s0 = 0, s1 = 0, s2 = 0, s3 = 0, s4 = 0;
#pragma unroll 2
  for (int i = -(nh + 1); i < n - 1; ++i) {  // advance the window centre from i to i + 1
    const int a = i + 1 + nh, o = i - nh;
    const double xin = (a < n) ? (double)__ldg(cx + a) - mx : 0.0, yin = (a < n) ? (double)__ldg(cy + a) - my : 0.0;
    const double xout = (o >= 0) ? (double)__ldg(cx + o) - mx : 0.0, yout = (o >= 0) ? (double)__ldg(cy + o) - my : 0.0;
    double vx = 0.0, vy = 0.0;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      const double dxr = fma(ar[k], xin, br[k] * xout), dxi = fma(ai[k], xin, bi[k] * xout);
      const double dyr = fma(ar[k], yin, br[k] * yout), dyi = fma(ai[k], yin, bi[k] * yout);
      const double nxr = fma(rr[k], xr[k], fma(-ri[k], xi[k], dxr)), nxi = fma(rr[k], xi[k], fma(ri[k], xr[k], dxi));
      const double nyr = fma(rr[k], yr[k], fma(-ri[k], yi[k], dyr)), nyi = fma(rr[k], yi[k], fma(ri[k], yr[k], dyi));
      xr[k] = nxr;
      xi[k] = nxi;
      yr[k] = nyr;
      yi[k] = nyi;
      vx = fma(wg[k], nxr, vx);
      vy = fma(wg[k], nyr, vy);
    }
    if (i + 1 >= 0) {
      s0 += vx;
      s1 += vy;
      s2 = fma(vx, vx, s2);
      s3 = fma(vy, vy, s3);
      s4 = fma(vx, vy, s4);
    }
  }
  double* dst = b.modsum + ((((int64_t)pair * kNumCep + j) * kNumMod) + m) * 5;
  dst[0] = s0;
  dst[1] = s1;
  dst[2] = s2;
  dst[3] = s3;
  dst[4] = s4;
}

// ---------------------------------------------------------------- score
__global__ void haspi_score_kernel(HaspiBuffers b, int n, double* intel, double* raw10, int32_t* status) {
  const int pair = blockIdx.x * blockDim.x + threadIdx.x;
  if (pair >= n) return;
  const double w[kNumMod] = {1.361, 1.521, 1.164, 0.492, 0.436, 0.690, 1.142, 0.816, 1.576, 2.269};
  const int ns = b.nsel[pair];
  if (ns <= 1) {  // pyhaspi2.py:357-358 raises; reported per pair instead
    intel[pair] = nan("");
    for (int m = 0; m < kNumMod; ++m) raw10[(int64_t)pair * kNumMod + m] = nan("");
    status[pair] = 1;
    return;
  }
  const double dn = (double)ns;
  double tot = 0.0;
  for (int m = 0; m < kNumMod; ++m) {
    double ave = 0.0;
    for (int j = 0; j < kNumCep; ++j) {
      const double* s = b.modsum + ((((int64_t)pair * kNumCep + j) * kNumMod) + m) * 5;
      const double xs = s[2] - s[0] * s[0] / dn, ys = s[3] - s[1] * s[1] / dn;
      const double xy = s[4] - s[0] * s[1] / dn;
      ave += (xs < 1.0e-30 || ys < 1.0e-30) ? 0.0 : fabs(xy) / sqrt(xs * ys);
    }
    ave /= (double)kNumCep;
    raw10[(int64_t)pair * kNumMod + m] = ave;
    tot += w[m] * ave;
  }
  intel[pair] = tot;
  status[pair] = 0;
}

// ------------------------------------------------------------- launchers
static size_t modcorr_smem_bytes() {
  return (size_t)(2 * (mod_skew(kModSpan) + 1) + kModTapsTotal + 8) * sizeof(float);
}

void haspi_upload_constants(cudaStream_t s) {
  float fir[54];
  make_env_fir(fir);
  cudaMemcpyToSymbolAsync(c_envfir, fir, sizeof(fir), 0, cudaMemcpyHostToDevice, s);
  const IhcConst ih = make_ihc_const();
  cudaMemcpyToSymbolAsync(c_ihc, &ih, sizeof(ih), 0, cudaMemcpyHostToDevice, s);
  cudaStreamSynchronize(s);  // sources are stack temporaries
}

void haspi_upload_tables(const float* cepm, const int* nhalf, const int* off, const float* taps, int ntaps,
                         cudaStream_t s) {
  cudaMemcpyToSymbolAsync(c_cepm, cepm, sizeof(float) * kBands * kNumCep, 0, cudaMemcpyHostToDevice, s);
  cudaMemcpyToSymbolAsync(c_mod_nhalf, nhalf, sizeof(int) * kNumMod, 0, cudaMemcpyHostToDevice, s);
  cudaMemcpyToSymbolAsync(c_mod_off, off, sizeof(int) * (kNumMod + 1), 0, cudaMemcpyHostToDevice, s);
  cudaMemcpyToSymbolAsync(g_modtaps, taps, sizeof(float) * ntaps, 0, cudaMemcpyHostToDevice, s);
  cudaFuncSetAttribute(haspi_modcorr_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)modcorr_smem_bytes());
  {
    host::ModRecBand mr[kNumMod];
    host::make_mod_recursions(mr);
    cudaMemcpyToSymbolAsync(g_modrec, mr, sizeof(mr), 0, cudaMemcpyHostToDevice, s);
    cudaStreamSynchronize(s);  // mr is a stack temporary
  }
  cudaFuncSetAttribute(haspi_ear_kernel<double>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                       (int)(kEarWarps * 4 * kEarChunk * sizeof(double)));
  cudaStreamSynchronize(s);
}

int haspi_run_front(const PairGeom& g, const HaspiBuffers& b, int n, bool f64, KernelTimer* kt, cudaStream_t s) {
  int launches = 0;
  kt_begin(kt, "haspi_prep", s);
  static const bool rs_f32 = [] { const char* p = getenv("NELE_RESAMPLE_F32"); return !(p && p[0] == '0'); }();
  if (rs_f32) haspi_prep_kernel<true><<<dim3(n, 2), kPrepThreads, (kPrepSpan + kPrepSpan / 8 + 8) * sizeof(double), s>>>(g, b);
  else haspi_prep_kernel<false><<<dim3(n, 2), kPrepThreads, (kPrepSpan + kPrepSpan / 8 + 8) * sizeof(double), s>>>(g, b);
  kt_end(kt, s);
  ++launches;
  const int ctas = (n + kEarWarps - 1) / kEarWarps;
  kt_begin(kt, "haspi_control", s);
  if (f64) haspi_control_kernel<double><<<ctas, kEarWarps * 32, 0, s>>>(g, b, n);
  else haspi_control_kernel<float><<<ctas, kEarWarps * 32, 0, s>>>(g, b, n);
  kt_end(kt, s);
  ++launches;
  kt_begin(kt, "haspi_shift", s);
  haspi_shift_kernel<<<(n * 32 + 127) / 128, 128, 0, s>>>(b, n);
  kt_end(kt, s);
  ++launches;
  return launches;
}

int haspi_run(const PairGeom& g, const HaspiBuffers& b, int n, int max_nsub, bool f64, KernelTimer* kt, cudaStream_t s) {
  int launches = haspi_run_front(g, b, n, f64, kt, s);
  kt_begin(kt, "haspi_ear", s);
  const int ear_ctas = (n + kEarWarps - 1) / kEarWarps;
  static const bool x2 = [] {  // packed main pass (default); NELE_F32X2=0 selects the scalar kernel for A/B checks
    const char* p = getenv("NELE_F32X2");
    return !(p && p[0] == '0');
  }();
  if (f64) haspi_ear_kernel<double><<<ear_ctas, kEarWarps * 32, kEarWarps * 4 * kEarChunk * sizeof(double), s>>>(g, b, n);
  else if (x2) haspi_ear_x2_kernel<<<ear_ctas, kEarWarps * 32, kEarWarps * 4 * kEarChunk * sizeof(float), s>>>(g, b, n);
  else haspi_ear_kernel<float><<<ear_ctas, kEarWarps * 32, kEarWarps * 4 * kEarChunk * sizeof(float), s>>>(g, b, n);
  kt_end(kt, s);
  ++launches;
  kt_begin(kt, "haspi_cep", s);
  haspi_cep_kernel<<<n, kCepThreads, 0, s>>>(g, b);
  kt_end(kt, s);
  ++launches;
  static const bool fir = [] {
    const char* p = getenv("NELE_MODCORR_FIR");
    return p && p[0] == '1';
  }();
  kt_begin(kt, "haspi_modcorr", s);
  if (fir) {
    cudaMemsetAsync(b.modsum, 0, sizeof(double) * (size_t)n * kNumCep * kNumMod * 5, s);
    const int tiles = (max_nsub + kModTile - 1) / kModTile;
    haspi_modcorr_kernel<<<dim3(n, kNumCep, tiles), kModThreads, modcorr_smem_bytes(), s>>>(g, b);
  } else {
    haspi_modcorr2_kernel<<<n, kMod2Threads, 0, s>>>(g, b);
  }
  kt_end(kt, s);
  ++launches;
  return launches;
}

int haspi_finish(const HaspiBuffers& b, int n, double* intel, double* raw10, int32_t* status, KernelTimer* kt,
                 cudaStream_t s) {
  kt_begin(kt, "haspi_score", s);
  haspi_score_kernel<<<(n + 127) / 128, 128, 0, s>>>(b, n, intel, raw10, status);
  kt_end(kt, s);
  return 1;
}

}  // namespace nele
