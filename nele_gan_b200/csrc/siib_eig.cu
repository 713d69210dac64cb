// Full-rank KLT of SIIB: eigen-decomposition of the 420 x 420 covariance Sxx by Householder
// tridiagonalisation instead of the one-sided Jacobi of siib.cu (which needs 10-11 sweeps of
// 88 k rotations whatever the preconditioning; scripts/exp_jacobi_sweeps.py).  Used for pairs whose
// numerical rank (pivoted Cholesky, siib_chol_kernel) exceeds 112; rank-deficient pairs keep the
// one-CTA Jacobi on the Cholesky factor.  The numerical recipe was validated on the CPU against
// numpy.linalg.eigh (scripts/exp_tridiag_eig.py: orthogonality <= 3e-8, SIIB equal to 1e-14).
//
//   siib_tridiag_kernel  per pair CTA, thread = row: A = Q T Q^T in FP64.  Work matrix in the
//                        Lc buffer (a copy of Sxx), read and updated row-wise so that every access
//                        of the symmetric matrix is coalesced; the Householder vector of step k
//                        stays in row k (columns > k + 1), as LAPACK keeps it in column k.
//   siib_trieig_kernel   per pair CTA, thread = eigenvalue: bisection with a division-free Sturm
//                        count (three-term recurrence, periodic rescaling), then the eigenvector
//                        of T from the twisted factorisation (dlar1v): the twist index is the
//                        minimum of |gamma|, the entries come from the two recurrences run from
//                        the ends towards the twist (their growing, i.e. stable, direction) with
//                        an explicit exponent, so no O(n) per-thread storage is needed
//   siib_backtf_kernel   per (pair, 64 eigenvectors) CTA, 4 lanes per vector: u = H_0 ... H_{n-3} z
//                        with the vector in registers, reflectors staged through shared memory;
//                        writes G[j] = sqrt(lambda_j) u_j, the format qf::quadform_kernel reads
#include <stdlib.h>

#include <algorithm>

#include "kernels.h"
#include "sturm.cuh"

namespace nele {

constexpr int kEDim = 420, kELd = 448, kEThreads = 448;

// ------------------------------------------------------------ tridiagonalisation
// Blocked as LAPACK's dsytrd / dlatrd: inside a panel of kTriB steps the trailing matrix is only
// *read* (one coalesced pass per step for p = S v); the rank-2 updates of the panel are kept as the
// n x kTriB matrices V, W in shared memory and applied to whatever is touched on the fly
// (S_cur = S_panel - V W^T - W V^T), and the trailing matrix is rewritten once per panel.  Traffic
// per element and step: 8 B + 16 B / kTriB instead of 16 B for the step-by-step form -- the kernel
// is HBM bound (ncu: 4.4 TB/s).
constexpr int kTriB = 8;

__device__ __forceinline__ void householder(double x, double alpha, double sig, int tid, int k, bool own, double& vi,
                                            double& beta, double& tau) {
  if (sig == 0.0) {  // nothing to annihilate
    beta = alpha;
    tau = 0.0;
    vi = 0.0;
    return;
  }
  const double nrm = sqrt(alpha * alpha + sig);
  beta = alpha >= 0.0 ? -nrm : nrm;
  tau = (beta - alpha) / beta;
  vi = (tid == k + 1) ? 1.0 : (tid > k + 1 && own) ? x / (alpha - beta) : 0.0;
}

// (Round 2 measured a variant whose matrix-vector pass read an FP32 copy of the trailing matrix, NELE_TRIDIAG_MV32:
// 55.0 -> 49.5 ms per 1024 pairs, SIIB equal to 3e-7.  The all-FP32 lower-triangle kernel of siib_klt.cu -- 13.9 ms --
// superseded it and it was removed; this FP64 kernel stays as the A/B reference behind NELE_TRIDIAG_F64=1.)
__global__ void __launch_bounds__(kEThreads, 3) siib_tridiag_kernel(SiibGeom g, SiibBuffers b, SiibEigBuffers eb, int rank_lo, int n_pairs) {
  const int tid = threadIdx.x, lane = tid & 31, wib = tid >> 5;
  constexpr int NW = kEThreads / 32;
  // persistent CTAs: the grid is sized so that the work matrices of the resident CTAs fit in the L2
  for (int lp = blockIdx.x; lp < n_pairs; lp += gridDim.x) {
  const int pair = b.pair_lo + lp;
  if (b.rank[pair] < rank_lo) continue;
  __syncthreads();
  const double* __restrict__ A0 = b.Sxx + (int64_t)lp * kEDim * kEDim;
  double* __restrict__ A = b.Lc + (int64_t)lp * kEDim * kEDim;
  double* __restrict__ dd = eb.d + (int64_t)lp * kELd;
  double* __restrict__ ee = eb.e + (int64_t)lp * kELd;
  double* __restrict__ tt = eb.tau + (int64_t)lp * kELd;
  extern __shared__ __align__(16) double s_vw[];  // V[kTriB][448], W[kTriB][448]
  double* sV = s_vw;
  double* sW = s_vw + kTriB * kELd;
  __shared__ double red[32];
  __shared__ double s_part[NW][2 * kTriB];
  __shared__ double s_tot[2 * kTriB];
  __shared__ double s_alpha;
  const bool own = tid < kEDim;
  for (int k0 = 0; k0 < kEDim - 2; k0 += kTriB) {
    const double* Ain = (k0 == 0) ? A0 : A;  // the matrix as of the start of the panel (aliases the work matrix:
                                             // read with ld.global.cg, never through the non-coherent path)
    const int nb = min(kTriB, kEDim - 2 - k0);
    for (int m = 0; m < nb; ++m) {
      const int k = k0 + m;
      // row k of the current matrix: panel-start values minus the updates of the steps before
      double x = (own && tid >= k) ? Ain[(int64_t)k * kEDim + tid] : 0.0;
      for (int mm = 0; mm < m; ++mm) x -= sV[mm * kELd + k] * sW[mm * kELd + tid] + sW[mm * kELd + k] * sV[mm * kELd + tid];
      if (tid == k) dd[k] = x;
      const double xr = (own && tid > k) ? x : 0.0;
      const double sig = block_sum((tid > k + 1) ? xr * xr : 0.0, red);
      if (tid == k + 1) s_alpha = xr;
      __syncthreads();
      double vi, beta, tau;
      householder(xr, s_alpha, sig, tid, k, own, vi, beta, tau);
      sV[m * kELd + tid] = vi;
      if (own && tid > k + 1) A[(int64_t)k * kEDim + tid] = vi;  // the reflector stays in row k of the work matrix
      eb.refl[((int64_t)lp * kEDim + k) * kELd + tid] = (float)vi;  // FP32 copy for the back-transformation (0 in rows <= k, 1 at k + 1)
      if (tid == 0) {
        ee[k] = beta;
        tt[k] = tau;
      }
      __syncthreads();
      // p = S_panel v: one coalesced read-only pass over rows j > k
      double p = 0.0;
      if (own && tid > k) {
        const double* col = Ain + tid;
        const double* __restrict__ v = sV + m * kELd;
        double p0 = 0.0, p1 = 0.0, p2 = 0.0, p3 = 0.0;
        int j = k + 1;
        for (; j + 8 <= kEDim; j += 8) {
          double a[8];
#pragma unroll
          for (int u = 0; u < 8; ++u) a[u] = __ldcg(col + (int64_t)(j + u) * kEDim);
          p0 = fma(a[0], v[j], p0);
          p1 = fma(a[1], v[j + 1], p1);
          p2 = fma(a[2], v[j + 2], p2);
          p3 = fma(a[3], v[j + 3], p3);
          p0 = fma(a[4], v[j + 4], p0);
          p1 = fma(a[5], v[j + 5], p1);
          p2 = fma(a[6], v[j + 6], p2);
          p3 = fma(a[7], v[j + 7], p3);
        }
        for (; j < kEDim; ++j) p0 = fma(__ldcg(col + (int64_t)j * kEDim), v[j], p0);
        p = (p0 + p1) + (p2 + p3);
      }
      // corrections for the steps of this panel: p -= V (W^T v) + W (V^T v)
      if (m > 0) {
        for (int mm = 0; mm < m; ++mm) {
          const double aw = warp_sum(sW[mm * kELd + tid] * vi), av = warp_sum(sV[mm * kELd + tid] * vi);
          if (lane == 0) {
            s_part[wib][2 * mm] = aw;
            s_part[wib][2 * mm + 1] = av;
          }
        }
        __syncthreads();
        if (tid < 2 * m) {
          double t = 0.0;
#pragma unroll
          for (int w = 0; w < NW; ++w) t += s_part[w][tid];
          s_tot[tid] = t;
        }
        __syncthreads();
        for (int mm = 0; mm < m; ++mm) p -= sV[mm * kELd + tid] * s_tot[2 * mm] + sW[mm * kELd + tid] * s_tot[2 * mm + 1];
      }
      p = (own && tid > k) ? tau * p : 0.0;
      const double pv = block_sum(p * vi, red);
      sW[m * kELd + tid] = p - (0.5 * tau * pv) * vi;
      __syncthreads();
    }
    // trailing update of the panel: rows and columns beyond its last step
    const int kl = k0 + nb - 1;
    if (own && tid > kl) {
      int j = kl + 1;
      for (; j + 4 <= kEDim; j += 4) {  // four independent loads in flight (the pass is latency bound otherwise)
        double a[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) a[u] = __ldcg(Ain + (int64_t)(j + u) * kEDim + tid);
        for (int mm = 0; mm < nb; ++mm) {
          const double wt = sW[mm * kELd + tid], vt = sV[mm * kELd + tid];
#pragma unroll
          for (int u = 0; u < 4; ++u) a[u] -= sV[mm * kELd + j + u] * wt + sW[mm * kELd + j + u] * vt;
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) A[(int64_t)(j + u) * kEDim + tid] = a[u];
      }
      for (; j < kEDim; ++j) {
        double a = Ain[(int64_t)j * kEDim + tid];
        for (int mm = 0; mm < nb; ++mm) a -= sV[mm * kELd + j] * sW[mm * kELd + tid] + sW[mm * kELd + j] * sV[mm * kELd + tid];
        A[(int64_t)j * kEDim + tid] = a;
      }
    }
    __syncthreads();
  }
  if (tid == 0) {
    dd[kEDim - 2] = A[(int64_t)(kEDim - 2) * kEDim + (kEDim - 2)];
    dd[kEDim - 1] = A[(int64_t)(kEDim - 1) * kEDim + (kEDim - 1)];
    ee[kEDim - 2] = A[(int64_t)(kEDim - 2) * kEDim + (kEDim - 1)];
    tt[kEDim - 2] = 0.0;
  }
  }
}

// --------------------------------------------------- eigenpairs of the tridiagonal
// Sturm counts: sturm.cuh (sign masks + popc, blocks of 16 steps, exponent-field overflow guard)
constexpr int kSturmLen = sturm_len(kEDim);  // 433
constexpr int kSturmLenSmall = sturm_len(112);  // 113: the Gram path (kSN below)

// value with an explicit binary exponent, kept near 1
struct Scaled {
  double m;
  int ex;
  __device__ __forceinline__ void norm() {
    int k;
    m = frexp(m, &k);
    ex += k;
  }
};

__global__ void __launch_bounds__(kEThreads) siib_trieig_kernel(SiibGeom g, SiibBuffers b, SiibEigBuffers eb, int rank_lo) {
  const int lp = blockIdx.x, pair = b.pair_lo + lp, tid = threadIdx.x;
  if (b.rank[pair] < rank_lo) return;
  __shared__ double s_d[kEDim], s_e[kEDim], s_e2[kEDim];
  __shared__ __align__(16) double2 s_de[kSturmLen];
  __shared__ int s_grid[kEDim + 1];
  __shared__ double s_ie[kEDim];   // 1 / e_i (the divisions by e_i of the vector recurrences are shared by all threads)
  __shared__ double red[32];
  const double* __restrict__ dd = eb.d + (int64_t)lp * kELd;
  const double* __restrict__ ee = eb.e + (int64_t)lp * kELd;
  // scale T by 1 / max|entry| (Sturm recurrence stays in range); eigenvalues scale back at the end
  double mx = 0.0;
  if (tid < kEDim) mx = fmax(fabs(dd[tid]), (tid < kEDim - 1) ? fabs(ee[tid]) : 0.0);
  const double scale = block_max(mx, red);
  const double inv = scale > 0.0 ? 1.0 / scale : 0.0;
  if (tid < kEDim) {
    s_d[tid] = dd[tid] * inv;
    const double ev = (tid < kEDim - 1) ? ee[tid] * inv : 0.0;
    s_e[tid] = ev;
    s_e2[tid] = ev * ev;
    s_ie[tid] = 1.0 / (ev != 0.0 ? ev : 1.0e-300);
    s_de[tid].x = dd[tid] * inv;
    if (tid + 1 < kEDim) s_de[tid + 1].y = ev * ev;
    if (tid == 0) s_de[0].y = 0.0;
  } else if (tid < kSturmLen) {
    s_de[tid] = make_double2(1000.0, 0.0);
  }
  __syncthreads();
  // ---- eigenvalue number tid (ascending) inside the Gershgorin interval [-3, 3].
  // One Sturm count per thread on a uniform grid of 420 cells first: every thread then finds its own cell by a binary
  // search of the counts (9 shared-memory reads instead of 8.7 bisection steps of 420 recurrence steps each) ...
  const double cell = 6.0 / kEDim;
  if (tid < kEDim) s_grid[tid] = tid ? sturm_count<kSturmLen>(s_de, -3.0 + cell * tid) : 0;
  if (tid == 0) s_grid[kEDim] = kEDim;
  __syncthreads();
  if (tid >= kEDim) return;  // no block-wide barrier below
  int klo = 0, khi = kEDim;  // count(grid[klo]) <= tid < count(grid[khi])
  while (khi - klo > 1) {
    const int km = (klo + khi) >> 1;
    if (s_grid[km] > tid) khi = km;
    else klo = km;
  }
  // ... then 38 bisection steps: 6 / 420 * 2^-38 = 5.2e-14 of max|T| (46 steps from the whole interval gave 8.5e-14).
  // T comes from an FP32 tridiagonalisation (entries known to 1e-7), and the score does not react to what fewer steps
  // cost -- orthogonality inside clusters of tiny eigenvalues, whose components carry no information: SIIB deviation
  // unchanged from 58 down to 38 steps from the whole interval (scripts/exp_klt_fp32.py iterations)
  double lo = -3.0 + cell * klo, hi = -3.0 + cell * khi;
  for (int it = 0; it < 38; ++it) {
    const double mid = 0.5 * (lo + hi);
    if (sturm_count<kSturmLen>(s_de, mid) > tid) hi = mid;
    else lo = mid;
  }
  const double lam = 0.5 * (lo + hi);
  eb.lam[(int64_t)lp * kELd + tid] = lam * scale;
  // ---- eigenvector: twisted factorisation of T - lam I
  // backward pass: dm_i (D- of the U D U^T factorisation from the bottom), kept in FP32 scratch
  float* __restrict__ scr = eb.scratch + (int64_t)lp * kEDim * kELd + tid;  // [i][thread]
  const double tiny = 1.0e-300;
  {
    double dm = s_d[kEDim - 1] - lam;
    scr[(int64_t)(kEDim - 1) * kELd] = (float)dm;
    for (int i = kEDim - 2; i >= 0; --i) {
      if (dm == 0.0) dm = tiny;
      dm = (s_d[i] - lam) - s_e2[i] * __drcp_rn(dm);
      scr[(int64_t)i * kELd] = (float)dm;
    }
  }
  // forward pass: dp_i, gamma_i = dp_i + dm_i - (d_i - lam); the scaled forward recurrence
  // z_{i+1} = -z_i dp_i / e_i is carried along and snapshotted at the running minimum of |gamma|
  int kt = 0;
  Scaled zk = {1.0, 0};
  {
    double dp = s_d[0] - lam, best = 1.0e300;
    Scaled z = {1.0, 0};
    for (int i = 0; i < kEDim; ++i) {
      const double gam = dp + (double)scr[(int64_t)i * kELd] - (s_d[i] - lam);
      if (fabs(gam) < best) {
        best = fabs(gam);
        kt = i;
        zk = z;
      }
      if (i == kEDim - 1) break;
      if (dp == 0.0) dp = tiny;
      z.m = -z.m * dp * s_ie[i];
      z.norm();
      dp = (s_d[i + 1] - lam) - s_e2[i] * __drcp_rn(dp);
    }
  }
  // value of the backward recurrence z_i = -z_{i+1} dm_{i+1} / e_i at the twist index
  Scaled zb = {1.0, 0};
  {
    double dm = s_d[kEDim - 1] - lam;
    for (int i = kEDim - 2; i >= kt; --i) {
      if (dm == 0.0) dm = tiny;
      zb.m = -zb.m * dm * s_ie[i];
      zb.norm();
      dm = (s_d[i] - lam) - s_e2[i] * __drcp_rn(dm);
    }
  }
  // write z (z_kt = 1) and its squared norm; entries below 2^-126 flush to zero in FP32
  double nrm2 = 0.0;
  {
    double dp = s_d[0] - lam;
    Scaled z = {1.0, 0};
    const double izk = 1.0 / zk.m;
    for (int i = 0; i <= kt; ++i) {
      const double v = ldexp(z.m * izk, z.ex - zk.ex);
      eb.zt[((int64_t)lp * kEDim + i) * kELd + tid] = (float)v;
      nrm2 += v * v;
      if (i == kt) break;
      if (dp == 0.0) dp = tiny;
      z.m = -z.m * dp * s_ie[i];
      z.norm();
      dp = (s_d[i + 1] - lam) - s_e2[i] * __drcp_rn(dp);
    }
  }
  {
    double dm = s_d[kEDim - 1] - lam;
    Scaled z = {1.0, 0};
    const double izb = 1.0 / zb.m;
    for (int i = kEDim - 1; i > kt; --i) {
      const double v = ldexp(z.m * izb, z.ex - zb.ex);
      eb.zt[((int64_t)lp * kEDim + i) * kELd + tid] = (float)v;
      nrm2 += v * v;
      if (dm == 0.0) dm = tiny;
      z.m = -z.m * dm * s_ie[i - 1];
      z.norm();
      dm = (s_d[i - 1] - lam) - s_e2[i - 1] * __drcp_rn(dm);
    }
  }
  eb.znorm[(int64_t)lp * kELd + tid] = nrm2;
}

// ------------------------------------------------------------ back-transformation
// Lane = eigenvector (32 per CTA), warp w of 8 holds its rows i = 8 m + w (m = 0..52, padded to 56)
// in registers: every element of a reflector is read by all 32 lanes at once (one shared-memory
// wavefront per 128-bit word; a 4-lanes-per-vector layout tried first was bound by shared-memory
// bandwidth), and 16 warps per SM hide the latency of the dependent dot -> update chain (the
// kernel is latency bound: fewer, fatter warps were slower).  The eight partial dot products of a
// vector meet in shared memory, one block barrier per reflector.  A reflector is staged as
// [w][m]; its all-zero head (rows <= k) is skipped in chunks of 64 rows, uniformly over the CTA.
constexpr int kBtVec = 32, kBtParts = 8, kBtThreads = kBtParts * 32;
constexpr int kBtRows = 56;                   // rows per warp (14 x 4), 8 * 56 = 448 >= 420
constexpr int kBtLd = kBtParts * kBtRows;     // staged reflector length
constexpr int kBtPanel = 16;                  // reflectors staged per panel
constexpr int kBtChunk = 2;                   // 128-bit words per skippable chunk

__global__ void __launch_bounds__(kBtThreads, 2) siib_backtf_kernel(SiibGeom g, SiibBuffers b, SiibEigBuffers eb, int rank_lo) {
  const int lp = blockIdx.y, pair = b.pair_lo + lp, tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  if (b.rank[pair] < rank_lo) return;
  const int j = blockIdx.x * kBtVec + lane;
  const bool live = j < kEDim;
  __shared__ __align__(16) float s_v[kBtPanel][kBtLd];
  __shared__ float s_tau[kBtPanel];
  __shared__ float s_dot[2][kBtParts][32];
  const float* __restrict__ R = eb.refl + (int64_t)lp * kEDim * kELd;
  const double* __restrict__ tt = eb.tau + (int64_t)lp * kELd;
  float u[kBtRows];
#pragma unroll
  for (int m = 0; m < kBtRows; ++m) {
    const int i = kBtParts * m + w;
    u[m] = (live && i < kEDim) ? eb.zt[((int64_t)lp * kEDim + i) * kELd + j] : 0.f;
  }
  int par = 0;
  // reflectors k = n-3 .. 0 in panels of 16 (k descending inside a panel)
  for (int k1 = kEDim - 3; k1 >= 0; k1 -= kBtPanel) {
    const int nk = min(kBtPanel, k1 + 1);
    __syncthreads();
    for (int kk = 0; kk < nk; ++kk) {
      const float* __restrict__ row = R + (int64_t)(k1 - kk) * kELd;
#pragma unroll
      for (int i = tid; i < kBtLd; i += kBtThreads) s_v[kk][(i & (kBtParts - 1)) * kBtRows + (i >> 3)] = __ldg(row + i);
    }
    if (tid < nk) s_tau[tid] = (float)tt[k1 - tid];
    __syncthreads();
    for (int kk = 0; kk < nk; ++kk) {
      const float tau = s_tau[kk];
      if (tau == 0.f) continue;
      const float4* v4 = reinterpret_cast<const float4*>(s_v[kk] + w * kBtRows);
      // reflector k is zero in rows <= k, i.e. in the words q < (k + 1) / 32 of every warp
      const int c0 = ((k1 - kk + 1) >> 5) / kBtChunk;
      float d0 = 0.f, d1 = 0.f, d2 = 0.f, d3 = 0.f;
#pragma unroll
      for (int c = 0; c < kBtRows / 4 / kBtChunk; ++c) {
        if (c < c0) continue;
#pragma unroll
        for (int q = c * kBtChunk; q < (c + 1) * kBtChunk; ++q) {
          const float4 vv = v4[q];
          d0 = fmaf(vv.x, u[4 * q], d0);
          d1 = fmaf(vv.y, u[4 * q + 1], d1);
          d2 = fmaf(vv.z, u[4 * q + 2], d2);
          d3 = fmaf(vv.w, u[4 * q + 3], d3);
        }
      }
      s_dot[par][w][lane] = (d0 + d1) + (d2 + d3);
      __syncthreads();
      float dot = 0.f;
#pragma unroll
      for (int t = 0; t < kBtParts; ++t) dot += s_dot[par][t][lane];
      par ^= 1;
      const float cf = -tau * dot;
#pragma unroll
      for (int c = 0; c < kBtRows / 4 / kBtChunk; ++c) {
        if (c < c0) continue;
#pragma unroll
        for (int q = c * kBtChunk; q < (c + 1) * kBtChunk; ++q) {
          const float4 vv = v4[q];
          u[4 * q] = fmaf(cf, vv.x, u[4 * q]);
          u[4 * q + 1] = fmaf(cf, vv.y, u[4 * q + 1]);
          u[4 * q + 2] = fmaf(cf, vv.z, u[4 * q + 2]);
          u[4 * q + 3] = fmaf(cf, vv.w, u[4 * q + 3]);
        }
      }
    }
  }
  if (!live) return;
  // column j of G = sqrt(lambda_j) u_j / |z_j|; eigenvalues at or below 1e-10 lambda_max carry no information
  const double lam = eb.lam[(int64_t)lp * kELd + j], lmax = eb.lam[(int64_t)lp * kELd + kEDim - 1];
  const double nz = eb.znorm[(int64_t)lp * kELd + j];
  // a periodic tiling of rank r < 420 (pivoted Cholesky, FP64): only the r largest eigenvalues are non-zero in exact
  // arithmetic; the FP32 tridiagonalisation leaves the others at +-1e-7 lambda_max, so the rank decides, not the value
  const bool in_range = j >= kEDim - b.rank[pair];
  const float sc = (in_range && lam > 1.0e-10 * lmax && nz > 0.0) ? (float)sqrt(lam / nz) : 0.f;
  float* __restrict__ G = b.G + (int64_t)lp * kEDim * kELd + (int64_t)j * kELd;
#pragma unroll
  for (int m = 0; m < kBtRows; ++m) {
    const int i = kBtParts * m + w;
    G[i] = (i < kEDim) ? sc * u[m] : 0.f;
  }
}

// rank <- 420 for the pairs that took this path (qf::quadform_kernel loops over `rank` columns)
__global__ void siib_eig_finish_kernel(SiibBuffers b, int n, int rank_lo) {
  const int lp = blockIdx.x * blockDim.x + threadIdx.x;
  if (lp >= n) return;
  const int pair = b.pair_lo + lp;
  if (b.rank[pair] >= rank_lo) {
    b.sweeps[pair] = (b.rank[pair] < kEDim) ? -3 : -1;  // marks "tridiagonal path" in the siib.rank stage (-3: rank deficient)
    b.rank[pair] = kEDim;
  }
}

int siib_run_eig(const SiibGeom& g, const SiibBuffers& b, const SiibEigBuffers& eb, int n, int rank_lo, KernelTimer* kt,
                 cudaStream_t s) {
  kt_begin(kt, "siib_tridiag", s);
  // default: FP32 lower-triangle kernel (siib_klt.cu); NELE_TRIDIAG_F64=1 selects the FP64 kernel above (A/B checks)
  static const bool tri_f64 = [] { const char* p = getenv("NELE_TRIDIAG_F64"); return p && p[0] == '1'; }();
  if (!tri_f64) {
    siib_launch_tridiag32(b, eb, n, rank_lo, s);
  } else {
  static const bool tri_attr = [] {
    cudaFuncSetAttribute(siib_tridiag_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 2 * kTriB * kELd * (int)sizeof(double));
    return true;
  }();
  (void)tri_attr;
  static const int tri_grid = [] { const char* p = getenv("NELE_TRIDIAG_GRID"); return p ? atoi(p) : 1 << 30; }();
  siib_tridiag_kernel<<<std::min(n, tri_grid), kEThreads, 2 * kTriB * kELd * sizeof(double), s>>>(g, b, eb, rank_lo, n);
  }
  kt_end(kt, s);
  kt_begin(kt, "siib_trieig", s);
  siib_trieig_kernel<<<n, kEThreads, 0, s>>>(g, b, eb, rank_lo);
  kt_end(kt, s);
  kt_begin(kt, "siib_backtf", s);
  static const bool bt_old = [] { const char* p = getenv("NELE_BACKTF_OLD"); return p && p[0] == '1'; }();  // A/B: one reflector per barrier
  if (bt_old) siib_backtf_kernel<<<dim3((kEDim + kBtVec - 1) / kBtVec, n), kBtThreads, 0, s>>>(g, b, eb, rank_lo);
  else {
    static const bool bt4 = [] { const char* p = getenv("NELE_BACKTF4"); return p && p[0] == '1'; }();  // A/B: one vector per lane
    static const bool bt5 = [] { const char* p = getenv("NELE_BACKTF5"); return p && p[0] == '1'; }();  // A/B: row parts over the CTA
    if (bt4) siib_launch_backtf4(b, eb, n, rank_lo, s);
    else if (bt5) siib_launch_backtf5(b, eb, n, rank_lo, s);
    else siib_launch_backtf6(b, eb, n, rank_lo, s);
  }
  kt_end(kt, s);
  kt_begin(kt, "siib_eig_finish", s);
  siib_eig_finish_kernel<<<(n + 127) / 128, 128, 0, s>>>(b, n, rank_lo);
  kt_end(kt, s);
  return 4;
}

}  // namespace nele

// ====================================================================================
// Low-rank pairs (periodic tilings: numerical rank r <= 112 of 420, siib_chol_kernel).
// With Sxx = L L^T (L: 420 x r, FP32 copy in G), the non-zero eigenpairs of Sxx follow from the
// r x r Gram matrix M = L^T L = V Lambda V^T:  sqrt(lambda_j) u_j = L v_j.  M is zero padded to
// 112 x 112 and goes through the same recipe as the full-rank case, but entirely on one SM
// (the FP64 matrix is 100 KB of shared memory); this replaces the one-CTA Jacobi on the 448-long
// columns of L (8-9 sweeps, 7.2 us per pair at bench size).
//
//   siib_gram_kernel      per pair CTA, 16 x 16 threads x 7 x 7 register tile: M in FP64
//   siib_smalleig_kernel  per pair CTA, thread = row / eigenvalue / eigenvector: unblocked
//                         Householder tridiagonalisation in shared memory, bisection + twisted
//                         factorisation, back-transformation with the vector in registers -> V
//   siib_lv_kernel        per (pair, 32-row tile) CTA: G <- L V in place
namespace nele {

constexpr int kSN = 112, kSNp = kSN + 1;  // padded Gram dimension, shared-memory row stride

// M = L^T L is symmetric: a thread owns one 7 x 7 tile (ty, tx) of the lower triangle (136 tiles for
// the 16 x 16 tile grid, 160 threads) and stores it to both halves; tiles beyond the rank are zero.
constexpr int kGramThreads = 160;

__global__ void __launch_bounds__(kGramThreads) siib_gram_kernel(SiibGeom g, SiibBuffers b, SiibEigBuffers eb, int rank_hi) {
  const int lp = blockIdx.x, pair = b.pair_lo + lp, tid = threadIdx.x;
  const int r = b.rank[pair];
  if (r < 2 || r > rank_hi) return;
  int ty = 0;
  while ((ty + 1) * (ty + 2) / 2 <= tid) ++ty;   // tid = ty (ty + 1) / 2 + tx, tx <= ty
  const int tx = tid - ty * (ty + 1) / 2;
  const bool work = tid < 136 && 7 * tx < r;     // 7 tx <= 7 ty: a tile left of the rank boundary has live columns
  const float* __restrict__ G = b.G + (int64_t)lp * kEDim * kELd;
  __shared__ float sL[kSN][33];
  double acc[7][7];
#pragma unroll
  for (int i = 0; i < 7; ++i)
#pragma unroll
    for (int j = 0; j < 7; ++j) acc[i][j] = 0.0;
  for (int row0 = 0; row0 < kEDim; row0 += 32) {
    __syncthreads();
    for (int idx = tid; idx < kSN * 32; idx += kGramThreads) {
      const int c = idx >> 5, rr = idx & 31;
      sL[c][rr] = (c < r && row0 + rr < kEDim) ? G[(int64_t)c * kELd + row0 + rr] : 0.f;
    }
    __syncthreads();
    if (work && 7 * ty < r) {
#pragma unroll 4
      for (int rr = 0; rr < 32; ++rr) {
        double a[7], c[7];
#pragma unroll
        for (int i = 0; i < 7; ++i) {
          a[i] = (double)sL[7 * ty + i][rr];
          c[i] = (double)sL[7 * tx + i][rr];
        }
#pragma unroll
        for (int i = 0; i < 7; ++i)
#pragma unroll
          for (int j = 0; j < 7; ++j) acc[i][j] = fma(a[i], c[j], acc[i][j]);
      }
    }
  }
  if (tid >= 136) return;
  double* __restrict__ M = eb.gram + (int64_t)lp * kSN * kSN;
#pragma unroll
  for (int i = 0; i < 7; ++i)
#pragma unroll
    for (int j = 0; j < 7; ++j) {
      M[(7 * ty + i) * kSN + 7 * tx + j] = acc[i][j];
      if (tx != ty) M[(7 * tx + j) * kSN + 7 * ty + i] = acc[i][j];
    }
}

// unblocked Householder tridiagonalisation of the padded Gram matrix in shared memory
// (thread = row = column); d, e, tau and the reflectors (in the eliminated rows) go back to global
__global__ void __launch_bounds__(128) siib_smalltri_kernel(SiibGeom g, SiibBuffers b, SiibEigBuffers eb, int rank_hi) {
  const int lp = blockIdx.x, pair = b.pair_lo + lp, tid = threadIdx.x;
  const int r = b.rank[pair];
  if (r < 2 || r > rank_hi) return;
  extern __shared__ __align__(16) double s_A[];  // [112][113]
  __shared__ double s_v[128], s_w[128];
  __shared__ double red[32];
  __shared__ double s_alpha;
  double* __restrict__ M = eb.gram + (int64_t)lp * kSN * kSN;
  double* __restrict__ dd = eb.d + (int64_t)lp * kELd;
  double* __restrict__ ee = eb.e + (int64_t)lp * kELd;
  double* __restrict__ tt = eb.tau + (int64_t)lp * kELd;
  for (int idx = tid; idx < kSN * kSN; idx += 128) s_A[(idx / kSN) * kSNp + idx % kSN] = M[idx];
  __syncthreads();
  // rows / columns >= r are zero padding: nothing to annihilate there (tau = 0), so the inner loops
  // stop at n = r rounded up to a multiple of four
  const int n = min(kSN, (r + 3) & ~3);
  const bool own = tid < n;
  for (int k = 0; k < kSN - 2; ++k) {
    const double x = (own && tid > k) ? s_A[k * kSNp + tid] : 0.0;
    const double sig = block_sum((tid > k + 1) ? x * x : 0.0, red);
    if (tid == k + 1) s_alpha = x;
    if (tid == k) dd[k] = s_A[k * kSNp + k];
    __syncthreads();
    double vi, beta, tau;
    householder(x, s_alpha, sig, tid, k, own, vi, beta, tau);
    s_v[tid] = vi;
    if (own && tid > k + 1) s_A[k * kSNp + tid] = vi;  // the reflector stays in row k
    if (tid == 0) {
      ee[k] = beta;
      tt[k] = tau;
    }
    __syncthreads();
    if (tau == 0.0) continue;
    double p = 0.0;
    if (own && tid > k) {
      double p0 = 0.0, p1 = 0.0, p2 = 0.0, p3 = 0.0;
      int j = k + 1;
      for (; j + 4 <= n; j += 4) {
        p0 = fma(s_A[j * kSNp + tid], s_v[j], p0);
        p1 = fma(s_A[(j + 1) * kSNp + tid], s_v[j + 1], p1);
        p2 = fma(s_A[(j + 2) * kSNp + tid], s_v[j + 2], p2);
        p3 = fma(s_A[(j + 3) * kSNp + tid], s_v[j + 3], p3);
      }
      for (; j < n; ++j) p0 = fma(s_A[j * kSNp + tid], s_v[j], p0);
      p = tau * ((p0 + p1) + (p2 + p3));
    }
    const double pv = block_sum(p * vi, red);
    const double w = p - (0.5 * tau * pv) * vi;
    s_w[tid] = w;
    __syncthreads();
    if (own && tid > k) {
#pragma unroll 4
      for (int j = k + 1; j < n; ++j) s_A[j * kSNp + tid] -= s_v[j] * w + s_w[j] * vi;
    }
    __syncthreads();
  }
  if (tid == 0) {
    dd[kSN - 2] = s_A[(kSN - 2) * kSNp + (kSN - 2)];
    dd[kSN - 1] = s_A[(kSN - 1) * kSNp + (kSN - 1)];
    ee[kSN - 2] = s_A[(kSN - 2) * kSNp + (kSN - 1)];
    ee[kSN - 1] = 0.0;
    tt[kSN - 2] = 0.0;
  }
  for (int idx = tid; idx < kSN * kSN; idx += 128) M[idx] = s_A[(idx / kSN) * kSNp + idx % kSN];
}

// eigenpairs of the 112 x 112 tridiagonal (as siib_trieig_kernel): thread = eigenvalue, descending,
// so that the <= r non-zero ones fill the first columns (qf::quadform_kernel reads `rank` of them)
__global__ void __launch_bounds__(128) siib_smallvec_kernel(SiibGeom g, SiibBuffers b, SiibEigBuffers eb, int rank_hi) {
  const int lp = blockIdx.x, pair = b.pair_lo + lp, tid = threadIdx.x;
  const int r = b.rank[pair];
  if (r < 2 || r > rank_hi) return;
  constexpr int kLen = kSturmLenSmall;
  static_assert(kSturmLenSmall == 1 + ((kSN - 1 + kSturmBlk - 1) / kSturmBlk) * kSturmBlk, "padded length follows kSN");
  __shared__ double s_d[kSN], s_e[kSN], s_e2[kSN];
  __shared__ __align__(16) double2 s_de[kLen];
  __shared__ int s_grid[kSN + 1];
  __shared__ double red[32];
  const double* __restrict__ dd = eb.d + (int64_t)lp * kELd;
  const double* __restrict__ ee = eb.e + (int64_t)lp * kELd;
  const bool own = tid < kSN;
  double mx = own ? fmax(fabs(dd[tid]), (tid < kSN - 1) ? fabs(ee[tid]) : 0.0) : 0.0;
  const double scale = block_max(mx, red);
  const double inv = scale > 0.0 ? 1.0 / scale : 0.0;
  if (own) {
    s_d[tid] = dd[tid] * inv;
    const double ev = (tid < kSN - 1) ? ee[tid] * inv : 0.0;
    s_e[tid] = ev;
    s_e2[tid] = ev * ev;
    s_de[tid].x = dd[tid] * inv;
    if (tid + 1 < kSN) s_de[tid + 1].y = ev * ev;
    if (tid == 0) s_de[0].y = 0.0;
  } else if (tid < kLen) {
    s_de[tid] = make_double2(1000.0, 0.0);
  }
  __syncthreads();
  // grid start and bisection as in siib_trieig_kernel: 52 steps from a cell of 6 / 112 end below the 6 * 2^-58 of the
  // 58 steps from the whole interval that this (FP64) path has always taken
  const double cell = 6.0 / kSN;
  if (own) s_grid[tid] = tid ? sturm_count<kLen>(s_de, -3.0 + cell * tid) : 0;
  if (tid == 0) s_grid[kSN] = kSN;
  __syncthreads();
  if (!own) return;
  float* __restrict__ scr = eb.zt + (int64_t)lp * kEDim * kELd;  // [i][128] D- sequence, then [i][128] z
  const int want = kSN - 1 - tid;  // ascending index of this thread's eigenvalue
  int klo = 0, khi = kSN;
  while (khi - klo > 1) {
    const int km = (klo + khi) >> 1;
    if (s_grid[km] > want) khi = km;
    else klo = km;
  }
  double lo = -3.0 + cell * klo, hi = -3.0 + cell * khi;
  for (int it = 0; it < 52; ++it) {
    const double mid = 0.5 * (lo + hi);
    if (sturm_count<kLen>(s_de, mid) > want) hi = mid;
    else lo = mid;
  }
  const double lam = 0.5 * (lo + hi);
  const double tiny = 1.0e-300;
  float* dmv = scr + tid;             // dmv[i * 128]
  float* zv = scr + kSN * 128 + tid;  // zv[i * 128]
  double nrm2 = 0.0;
  {
    double dm = s_d[kSN - 1] - lam;
    dmv[(kSN - 1) * 128] = (float)dm;
    for (int i = kSN - 2; i >= 0; --i) {
      if (dm == 0.0) dm = tiny;
      dm = (s_d[i] - lam) - s_e2[i] / dm;
      dmv[i * 128] = (float)dm;
    }
  }
  int kt = 0;
  Scaled zk = {1.0, 0};
  {
    double dp = s_d[0] - lam, best = 1.0e300;
    Scaled z = {1.0, 0};
    for (int i = 0; i < kSN; ++i) {
      const double gam = dp + (double)dmv[i * 128] - (s_d[i] - lam);
      if (fabs(gam) < best) {
        best = fabs(gam);
        kt = i;
        zk = z;
      }
      if (i == kSN - 1) break;
      if (dp == 0.0) dp = tiny;
      const double ei = s_e[i] != 0.0 ? s_e[i] : tiny;
      z.m = -z.m * dp / ei;
      z.norm();
      dp = (s_d[i + 1] - lam) - s_e2[i] / dp;
    }
  }
  Scaled zb = {1.0, 0};
  {
    double dm = s_d[kSN - 1] - lam;
    for (int i = kSN - 2; i >= kt; --i) {
      if (dm == 0.0) dm = tiny;
      const double ei = s_e[i] != 0.0 ? s_e[i] : tiny;
      zb.m = -zb.m * dm / ei;
      zb.norm();
      dm = (s_d[i] - lam) - s_e2[i] / dm;
    }
  }
  {
    double dp = s_d[0] - lam;
    Scaled z = {1.0, 0};
    for (int i = 0; i <= kt; ++i) {
      const double v = ldexp(z.m / zk.m, z.ex - zk.ex);
      zv[i * 128] = (float)v;
      nrm2 += v * v;
      if (i == kt) break;
      if (dp == 0.0) dp = tiny;
      const double ei = s_e[i] != 0.0 ? s_e[i] : tiny;
      z.m = -z.m * dp / ei;
      z.norm();
      dp = (s_d[i + 1] - lam) - s_e2[i] / dp;
    }
  }
  {
    double dm = s_d[kSN - 1] - lam;
    Scaled z = {1.0, 0};
    for (int i = kSN - 1; i > kt; --i) {
      const double v = ldexp(z.m / zb.m, z.ex - zb.ex);
      zv[i * 128] = (float)v;
      nrm2 += v * v;
      if (dm == 0.0) dm = tiny;
      const double ei = s_e[i - 1] != 0.0 ? s_e[i - 1] : tiny;
      z.m = -z.m * dm / ei;
      z.norm();
      dm = (s_d[i - 1] - lam) - s_e2[i - 1] / dm;
    }
  }
  eb.lam[(int64_t)lp * kELd + tid] = lam * scale;
  eb.znorm[(int64_t)lp * kELd + tid] = nrm2;
}

// back-transformation u = H_0 ... H_{n-3} z of the 112 eigenvectors: thread = vector (in registers),
// reflectors broadcast from shared memory (FP32 copy), zero heads skipped in chunks of 16 rows
__global__ void __launch_bounds__(128) siib_smallback_kernel(SiibGeom g, SiibBuffers b, SiibEigBuffers eb, int rank_hi) {
  const int lp = blockIdx.x, pair = b.pair_lo + lp, tid = threadIdx.x;
  const int r = b.rank[pair];
  if (r < 2 || r > rank_hi) return;
  extern __shared__ __align__(16) float s_R[];  // [112][112] reflectors, row k = v_k (1 at k + 1, 0 before)
  __shared__ float s_tau[kSN];
  const double* __restrict__ M = eb.gram + (int64_t)lp * kSN * kSN;
  for (int idx = tid; idx < kSN * kSN; idx += 128) {
    const int k = idx / kSN, i = idx % kSN;
    s_R[idx] = (i == k + 1) ? 1.f : (i > k + 1) ? (float)M[idx] : 0.f;
  }
  if (tid < kSN) s_tau[tid] = (float)eb.tau[(int64_t)lp * kELd + tid];
  __syncthreads();
  if (tid >= kSN) return;
  float* __restrict__ scr = eb.zt + (int64_t)lp * kEDim * kELd;
  float* __restrict__ Vout = scr + 2 * kSN * 128;  // [c][112] unit eigenvectors of M
  float u[kSN];
  {
    const float* zv = scr + kSN * 128 + tid;
#pragma unroll
    for (int i = 0; i < kSN; ++i) u[i] = zv[i * 128];
  }
  for (int k = kSN - 3; k >= 0; --k) {
    const float tau = s_tau[k];
    if (tau == 0.f) continue;
    const float4* v4 = reinterpret_cast<const float4*>(s_R + k * kSN);
    const int c0 = (k + 1) >> 4;  // 16-row chunks that are entirely zero (rows <= k)
    float d0 = 0.f, d1 = 0.f, d2 = 0.f, d3 = 0.f;
#pragma unroll
    for (int c = 0; c < kSN / 16; ++c) {
      if (c < c0) continue;
#pragma unroll
      for (int q = 4 * c; q < 4 * c + 4; ++q) {
        const float4 vv = v4[q];
        d0 = fmaf(vv.x, u[4 * q], d0);
        d1 = fmaf(vv.y, u[4 * q + 1], d1);
        d2 = fmaf(vv.z, u[4 * q + 2], d2);
        d3 = fmaf(vv.w, u[4 * q + 3], d3);
      }
    }
    const float cf = -tau * ((d0 + d1) + (d2 + d3));
#pragma unroll
    for (int c = 0; c < kSN / 16; ++c) {
      if (c < c0) continue;
#pragma unroll
      for (int q = 4 * c; q < 4 * c + 4; ++q) {
        const float4 vv = v4[q];
        u[4 * q] = fmaf(cf, vv.x, u[4 * q]);
        u[4 * q + 1] = fmaf(cf, vv.y, u[4 * q + 1]);
        u[4 * q + 2] = fmaf(cf, vv.z, u[4 * q + 2]);
        u[4 * q + 3] = fmaf(cf, vv.w, u[4 * q + 3]);
      }
    }
  }
  const double lam = eb.lam[(int64_t)lp * kELd + tid], lmax = eb.lam[(int64_t)lp * kELd];
  const double nrm2 = eb.znorm[(int64_t)lp * kELd + tid];
  const float sc = (lam > 1.0e-10 * lmax && nrm2 > 0.0) ? (float)(1.0 / sqrt(nrm2)) : 0.f;
#pragma unroll
  for (int c = 0; c < kSN; ++c) Vout[c * kSN + tid] = sc * u[c];
}

// G <- L V in place, one 32-row tile per CTA: out[j][row] = sum_c L[c][row] V[c][j]
__global__ void __launch_bounds__(128) siib_lv_kernel(SiibGeom g, SiibBuffers b, SiibEigBuffers eb, int rank_hi) {
  const int lp = blockIdx.y, pair = b.pair_lo + lp, tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  const int r = b.rank[pair];
  if (r < 2 || r > rank_hi) return;
  const int row0 = blockIdx.x * 32;
  float* __restrict__ G = b.G + (int64_t)lp * kEDim * kELd;
  const float* __restrict__ V = eb.zt + (int64_t)lp * kEDim * kELd + 2 * kSN * 128;
  extern __shared__ __align__(16) float s_lv[];
  float* sL = s_lv;              // [112][33]
  float* sV = s_lv + kSN * 33;   // [112][112]
  for (int idx = tid; idx < kSN * 32; idx += 128) {
    const int c = idx >> 5, rr = idx & 31;
    sL[c * 33 + rr] = (c < r && row0 + rr < kELd) ? G[(int64_t)c * kELd + row0 + rr] : 0.f;
  }
  for (int idx = tid; idx < r * kSN; idx += 128) sV[idx] = V[idx];   // columns of L beyond the rank are zero
  __syncthreads();
  float acc[28];
#pragma unroll
  for (int t = 0; t < 28; ++t) acc[t] = 0.f;
  for (int c = 0; c < r; ++c) {
    const float a = sL[c * 33 + lane];
    const float4* v4 = reinterpret_cast<const float4*>(sV + c * kSN + 28 * w);
#pragma unroll
    for (int q = 0; q < 7; ++q) {
      const float4 vv = v4[q];
      acc[4 * q] = fmaf(a, vv.x, acc[4 * q]);
      acc[4 * q + 1] = fmaf(a, vv.y, acc[4 * q + 1]);
      acc[4 * q + 2] = fmaf(a, vv.z, acc[4 * q + 2]);
      acc[4 * q + 3] = fmaf(a, vv.w, acc[4 * q + 3]);
    }
  }
  if (row0 + lane < kELd) {
#pragma unroll
    for (int t = 0; t < 28; ++t) G[(int64_t)(28 * w + t) * kELd + row0 + lane] = acc[t];
  }
}

__global__ void siib_small_finish_kernel(SiibBuffers b, int n, int rank_hi) {
  const int lp = blockIdx.x * blockDim.x + threadIdx.x;
  if (lp >= n) return;
  const int pair = b.pair_lo + lp;
  const int r = b.rank[pair];
  if (r >= 2 && r <= rank_hi) b.sweeps[pair] = -2;  // marks "Gram path" in the siib.rank stage
}

int siib_run_small_eig(const SiibGeom& g, const SiibBuffers& b, const SiibEigBuffers& eb, int n, int rank_hi, KernelTimer* kt,
                       cudaStream_t s) {
  static const bool attr = [] {
    cudaFuncSetAttribute(siib_smalltri_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSN * kSNp * (int)sizeof(double));
    cudaFuncSetAttribute(siib_smallback_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSN * kSN * (int)sizeof(float));
    cudaFuncSetAttribute(siib_lv_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (kSN * 33 + kSN * kSN) * (int)sizeof(float));
    return true;
  }();
  (void)attr;
  kt_begin(kt, "siib_gram", s);
  siib_gram_kernel<<<n, kGramThreads, 0, s>>>(g, b, eb, rank_hi);
  kt_end(kt, s);
  kt_begin(kt, "siib_smalltri", s);
  siib_smalltri_kernel<<<n, 128, kSN * kSNp * sizeof(double), s>>>(g, b, eb, rank_hi);
  kt_end(kt, s);
  kt_begin(kt, "siib_smallvec", s);
  siib_smallvec_kernel<<<n, 128, 0, s>>>(g, b, eb, rank_hi);
  kt_end(kt, s);
  kt_begin(kt, "siib_smallback", s);
  siib_smallback_kernel<<<n, 128, kSN * kSN * sizeof(float), s>>>(g, b, eb, rank_hi);
  kt_end(kt, s);
  kt_begin(kt, "siib_lv", s);
  siib_lv_kernel<<<dim3(kELd / 32, n), 128, (kSN * 33 + kSN * kSN) * sizeof(float), s>>>(g, b, eb, rank_hi);
  kt_end(kt, s);
  kt_begin(kt, "siib_small_finish", s);
  siib_small_finish_kernel<<<(n + 127) / 128, 128, 0, s>>>(b, n, rank_hi);
  kt_end(kt, s);
  return 6;
}

}  // namespace nele
