// Full-rank KLT of SIIB, FP32 work matrix: Householder tridiagonalisation of the 420 x 420
// covariance Sxx streamed from L2 as a *lower triangle of floats*.
//
// Why FP32 is enough (scripts/exp_klt_fp32.py, tests/test_oracle_estoi_siib.py): SIIB^Gauss sums
// -1/2 log2(1 - (0.75 rho_j)^2) over all 420 KLT components, and that sum is insensitive to rotations
// inside clusters of close eigenvalues -- which is all an FP32 tridiagonalisation does to the basis.
// Measured against numpy.linalg.eigh in FP64 on synthetic and real speech pairs (condition numbers
// 2e5 .. 1e9): relative deviation of SIIB 3e-6 .. 2e-5 with storage *and* arithmetic in FP32 -- LAPACK's
// own ssyev gives the same 4e-7 .. 1.5e-5 -- against a tolerance of 5e-3.  The eigenpairs of the
// tridiagonal T (bisection + twisted factorisation, siib_eig.cu) stay FP64: they cost O(n^2) and keep
// the basis orthogonal to 3e-8.
//
// Why it matters: the FP64 kernel it replaces (siib_tridiag_kernel, kept behind NELE_TRIDIAG_F64=1)
// streams the full 1.41 MB matrix once per step: ~250 MB of L2 / HBM traffic per pair, 56 ms per 1024
// pairs (round 1, 47 % of the SIIB step on utterances whose tiling does not repeat).  Here a step
// reads only the lower triangle of the trailing matrix as floats: n^3 / 6 * 4 B = 49 MB per pair for
// the symmetric matrix-vector products plus 12 MB for the once-per-panel rank-16 update, and the
// 353 KB working set of a resident CTA (2 per SM) stays in the 126 MB L2.
//
//   CTA = pair, 256 threads (8 warps), panels of 8 steps as LAPACK's dsytrd / dlatrd:
//     * the 8 columns of a panel live in shared memory (written by the previous panel's update, so
//       the strided column of the lower triangle is never gathered from global memory);
//     * p = A v: the trailing lower triangle is cut into 16-row x 64-column tiles dealt round-robin
//       to the warps; a lane loads float2 per row (256 B coalesced rows), accumulates the column part
//       p[c] += A[j][c] v[j] in registers and the row part p[j] += A[j][c] v[c] through a 16-value
//       transposed warp reduction (16 shuffles per tile); partial vectors are per warp, combined
//       in a fixed order: no atomics, bit-reproducible;
//     * the rank-2 updates of the panel stay in shared memory (V, W) and are applied on the fly to the
//       panel columns and to p; the trailing matrix is rewritten once per panel.
#include <stdlib.h>

#include <algorithm>

#include "kernels.h"

namespace nele {
namespace klt {

constexpr int N = 420, LD = 448, NT = 256, NW = NT / 32, PB = 8, RG = 16, CB = 64;
constexpr int NE = (LD + NT - 1) / NT;  // vector elements per thread (j = tid, tid + 256)
constexpr int JMAX = (N - 1) / RG;      // last row group

struct Smem {
  float V[PB][LD];     // Householder vectors of the panel
  float W[PB][LD];
  float col[PB][LD];   // the panel's columns as of the start of the panel (rows >= column index)
  float pw[NW][LD];    // per-warp partial products
  float Vt[LD][PB];    // V, W transposed for the trailing update (one 32-byte row per matrix row)
  float Wt[LD][PB];
  float part[NW][2 * PB];
  float tot[2 * PB];
  float red[NW];
  float alpha;
};

__device__ __forceinline__ float bsum(float v, float* red) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  v = warp_sum(v);
  __syncthreads();
  if (lane == 0) red[w] = v;
  __syncthreads();
  float t = 0.f;
#pragma unroll
  for (int i = 0; i < NW; ++i) t += red[i];
  return t;
}

// tiles (J, C) of the lower triangle with rows > k, columns > k: row groups J = (k + 1) / 16 .. 26, column
// blocks C = (k + 1) / 64 .. J / 4; warp w takes tiles w, w + 8, ... of that enumeration
struct TileIter {
  int J, C, cmin, rem;
  __device__ __forceinline__ void init(int k, int wib) {
    J = (k + 1) / RG;
    cmin = (k + 1) / CB;
    C = cmin;
    rem = wib;
  }
  __device__ __forceinline__ bool next() {  // positions (J, C) on the next tile of this warp; false when done
    while (J <= JMAX) {
      const int avail = (J >> 2) - C + 1;
      if (rem < avail) {
        C += rem;
        rem = NW;
        return true;
      }
      rem -= avail;
      ++J;
      C = cmin;
    }
    return false;
  }
};

// one tile of p = A v (A symmetric, lower triangle stored).  v: shared-memory vector (zero for indices <= k and
// >= N), pw: this warp's partial result.
template <bool DIAG>
__device__ __forceinline__ void mv_tile(const float* __restrict__ A, const float* __restrict__ v, float* __restrict__ pw, int J, int C,
                                        int k, int lane) {
  const int j0 = RG * J, c0 = CB * C + 2 * lane;
  const bool colok = c0 + 1 > k;
  float2 a[RG];
#pragma unroll
  for (int r = 0; r < RG; ++r) {
    const int j = j0 + r;
    bool ok = colok && j > k && j < N;
    if (DIAG) ok = ok && c0 <= j;
    a[r] = ok ? __ldcg(reinterpret_cast<const float2*>(A + (size_t)j * LD + c0)) : make_float2(0.f, 0.f);
    if (DIAG && c0 + 1 > j) a[r].y = 0.f;
  }
  const float2 vc = *reinterpret_cast<const float2*>(v + c0);
  float acc0 = 0.f, acc1 = 0.f;
  float t[RG];
#pragma unroll
  for (int q = 0; q < RG / 4; ++q) {
    const float4 vj4 = *reinterpret_cast<const float4*>(v + j0 + 4 * q);
    const float vj[4] = {vj4.x, vj4.y, vj4.z, vj4.w};
#pragma unroll
    for (int rr = 0; rr < 4; ++rr) {
      const int r = 4 * q + rr;
      acc0 = fmaf(a[r].x, vj[rr], acc0);
      acc1 = fmaf(a[r].y, vj[rr], acc1);
      float ax = a[r].x, ay = a[r].y;
      if (DIAG) {  // the diagonal element belongs to the column part only
        const int j = j0 + r;
        if (c0 == j) ax = 0.f;
        if (c0 + 1 == j) ay = 0.f;
      }
      t[r] = fmaf(ax, vc.x, ay * vc.y);
    }
  }
  // row sums over the 32 lanes, 16 rows at once: after the exchanges lane l holds the sum of row l >> 1
  const unsigned full = 0xffffffffu;
  const bool b4 = lane & 16, b3 = lane & 8, b2 = lane & 4, b1 = lane & 2;
  float u8[8], u4[4], u2[2];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const float mine = b4 ? t[i + 8] : t[i], other = b4 ? t[i] : t[i + 8];
    u8[i] = mine + __shfl_xor_sync(full, other, 16);
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float mine = b3 ? u8[i + 4] : u8[i], other = b3 ? u8[i] : u8[i + 4];
    u4[i] = mine + __shfl_xor_sync(full, other, 8);
  }
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    const float mine = b2 ? u4[i + 2] : u4[i], other = b2 ? u4[i] : u4[i + 2];
    u2[i] = mine + __shfl_xor_sync(full, other, 4);
  }
  float u1;
  {
    const float mine = b1 ? u2[1] : u2[0], other = b1 ? u2[0] : u2[1];
    u1 = mine + __shfl_xor_sync(full, other, 2);
  }
  u1 += __shfl_xor_sync(full, u1, 1);
  float2* pc = reinterpret_cast<float2*>(pw + c0);
  float2 o = *pc;
  o.x += acc0;
  o.y += acc1;
  *pc = o;
  __syncwarp();
  if (!(lane & 1)) pw[j0 + (lane >> 1)] += u1;
  __syncwarp();
}

// one tile of the trailing update A -= V W^T + W V^T (rows and columns > kl, lower triangle); columns
// kn .. kn + 7 (the next panel) are copied to s.col on the way
template <bool DIAG>
__device__ __forceinline__ void up_tile(float* __restrict__ A, Smem& s, int J, int C, int kl, int kn, int lane) {
  const int j0 = RG * J, c0 = CB * C + 2 * lane;
  const bool colok = c0 > kl;  // kl is odd, c0 even: both columns of the lane are trailing columns
  float2 a[RG];
#pragma unroll
  for (int r = 0; r < RG; ++r) {
    const int j = j0 + r;
    bool ok = colok && j > kl && j < N;
    if (DIAG) ok = ok && c0 <= j;
    a[r] = ok ? __ldcg(reinterpret_cast<const float2*>(A + (size_t)j * LD + c0)) : make_float2(0.f, 0.f);
  }
  float2 vcx[PB], wcx[PB];
#pragma unroll
  for (int mm = 0; mm < PB; ++mm) {
    vcx[mm] = *reinterpret_cast<const float2*>(&s.V[mm][c0]);
    wcx[mm] = *reinterpret_cast<const float2*>(&s.W[mm][c0]);
  }
  const bool tocol = (c0 & ~7) == kn;
#pragma unroll
  for (int r = 0; r < RG; ++r) {
    const int j = j0 + r;
    bool ok = colok && j > kl && j < N;
    if (DIAG) ok = ok && c0 <= j;
    const float4 v0 = *reinterpret_cast<const float4*>(&s.Vt[j][0]), v1 = *reinterpret_cast<const float4*>(&s.Vt[j][4]);
    const float4 w0 = *reinterpret_cast<const float4*>(&s.Wt[j][0]), w1 = *reinterpret_cast<const float4*>(&s.Wt[j][4]);
    const float vr[PB] = {v0.x, v0.y, v0.z, v0.w, v1.x, v1.y, v1.z, v1.w};
    const float wr[PB] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
    float ax = a[r].x, ay = a[r].y;
#pragma unroll
    for (int mm = 0; mm < PB; ++mm) {
      ax = fmaf(-vr[mm], wcx[mm].x, ax);
      ax = fmaf(-wr[mm], vcx[mm].x, ax);
      ay = fmaf(-vr[mm], wcx[mm].y, ay);
      ay = fmaf(-wr[mm], vcx[mm].y, ay);
    }
    if (ok) {
      const bool both = !DIAG || c0 + 1 <= j;
      if (both) *reinterpret_cast<float2*>(A + (size_t)j * LD + c0) = make_float2(ax, ay);
      else A[(size_t)j * LD + c0] = ax;
      if (tocol) {
        s.col[c0 - kn][j] = ax;
        if (both) s.col[c0 + 1 - kn][j] = ay;
      }
    }
  }
}

// Sxx (FP64, [420][420]) -> d, e, tau (FP64 arrays of the FP32 results, the format siib_trieig_kernel reads) and
// the reflectors refl[k][448] (row k = v_k: 0 up to k, 1 at k + 1), the format siib_backtf_kernel reads.
// Work matrix: Wk [420][448] floats per pair (aliases the FP64 Cholesky buffer Lc, unused on this path).
__global__ void __launch_bounds__(NT, 2) tridiag32_kernel(SiibBuffers b, SiibEigBuffers eb, int rank_lo, int n_pairs) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  Smem& s = *reinterpret_cast<Smem*>(smem_raw);
  const int tid = threadIdx.x, lane = tid & 31, wib = tid >> 5;
  for (int lp = blockIdx.x; lp < n_pairs; lp += gridDim.x) {
    const int pair = b.pair_lo + lp;
    if (b.rank[pair] < rank_lo) continue;
    __syncthreads();
    const double* __restrict__ S = b.Sxx + (int64_t)lp * N * N;
    float* __restrict__ A = reinterpret_cast<float*>(b.Lc + (int64_t)lp * N * N);
    double* __restrict__ dd = eb.d + (int64_t)lp * LD;
    double* __restrict__ ee = eb.e + (int64_t)lp * LD;
    double* __restrict__ tt = eb.tau + (int64_t)lp * LD;
    float* __restrict__ R = eb.refl + (int64_t)lp * N * LD;
    // lower triangle of Sxx as floats (pairs of columns; the element right of the diagonal is never read)
    for (int j = wib; j < N; j += NW) {
      const double2* src = reinterpret_cast<const double2*>(S + (int64_t)j * N);
      float2* dst = reinterpret_cast<float2*>(A + (size_t)j * LD);
      for (int c2 = lane; 2 * c2 <= j; c2 += 32) {
        const double2 v = src[c2];
        dst[c2] = make_float2((float)v.x, (float)v.y);
      }
    }
    // columns of the first panel = rows of the symmetric input
    for (int idx = tid; idx < PB * LD; idx += NT) {
      const int q = idx / LD, j = idx % LD;
      s.col[q][j] = (j < N) ? (float)S[(int64_t)q * N + j] : 0.f;
    }
    __syncthreads();
    for (int k0 = 0; k0 < N - 2; k0 += PB) {
      const int nb = min(PB, N - 2 - k0);
      for (int m = 0; m < nb; ++m) {
        const int k = k0 + m;
        // ---- column k of the current matrix, Householder vector
        float x[NE], vi[NE];
        float sig = 0.f;
#pragma unroll
        for (int e = 0; e < NE; ++e) {
          const int j = tid + e * NT;
          float xv = 0.f;
          if (j < N && j >= k) {
            xv = s.col[m][j];
            for (int mm = 0; mm < m; ++mm) xv -= s.V[mm][k] * s.W[mm][j] + s.W[mm][k] * s.V[mm][j];
          }
          if (j == k) dd[k] = (double)xv;
          if (j == k + 1) s.alpha = xv;
          x[e] = (j > k) ? xv : 0.f;
          if (j > k + 1) sig = fmaf(xv, xv, sig);
        }
        sig = bsum(sig, s.red);
        const float alpha = s.alpha;
        float beta, tau;
        if (sig == 0.f) {  // nothing to annihilate
          beta = alpha;
          tau = 0.f;
#pragma unroll
          for (int e = 0; e < NE; ++e) vi[e] = 0.f;
        } else {
          const float nrm = sqrtf(fmaf(alpha, alpha, sig));
          beta = alpha >= 0.f ? -nrm : nrm;
          tau = (beta - alpha) / beta;
          const float inv = 1.f / (alpha - beta);
#pragma unroll
          for (int e = 0; e < NE; ++e) {
            const int j = tid + e * NT;
            vi[e] = (j == k + 1) ? 1.f : (j > k + 1 && j < N) ? x[e] * inv : 0.f;
          }
        }
#pragma unroll
        for (int e = 0; e < NE; ++e) {
          const int j = tid + e * NT;
          if (j < LD) {
            s.V[m][j] = vi[e];
            R[(size_t)k * LD + j] = vi[e];
          }
        }
        if (tid == 0) {
          ee[k] = (double)beta;
          tt[k] = (double)tau;
        }
        // dot products of v with the panel's earlier vectors (for the corrections of p below)
        for (int mm = 0; mm < m; ++mm) {
          float aw = 0.f, av = 0.f;
#pragma unroll
          for (int e = 0; e < NE; ++e) {
            const int j = tid + e * NT;
            if (j < LD) {
              aw = fmaf(s.W[mm][j], vi[e], aw);
              av = fmaf(s.V[mm][j], vi[e], av);
            }
          }
          aw = warp_sum(aw);
          av = warp_sum(av);
          if (lane == 0) {
            s.part[wib][2 * mm] = aw;
            s.part[wib][2 * mm + 1] = av;
          }
        }
        // this warp's partial product vector starts at zero
#pragma unroll
        for (int i = 0; i < LD / 32; ++i) s.pw[wib][lane + 32 * i] = 0.f;
        __syncthreads();
        if (tau != 0.f) {
          if (tid < 2 * m) {
            float t = 0.f;
#pragma unroll
            for (int w = 0; w < NW; ++w) t += s.part[w][tid];
            s.tot[tid] = t;
          }
          // ---- p = A v over the trailing lower triangle
          TileIter it;
          it.init(k, wib);
          while (it.next()) {
            if (CB * it.C + CB - 1 > RG * it.J) mv_tile<true>(A, s.V[m], s.pw[wib], it.J, it.C, k, lane);
            else mv_tile<false>(A, s.V[m], s.pw[wib], it.J, it.C, k, lane);
          }
        }
        __syncthreads();
        float p[NE];
        float pv = 0.f;
#pragma unroll
        for (int e = 0; e < NE; ++e) {
          const int j = tid + e * NT;
          float pj = 0.f;
          if (tau != 0.f && j > k && j < N) {
#pragma unroll
            for (int w = 0; w < NW; ++w) pj += s.pw[w][j];
            for (int mm = 0; mm < m; ++mm) pj -= s.V[mm][j] * s.tot[2 * mm] + s.W[mm][j] * s.tot[2 * mm + 1];
            pj *= tau;
          }
          p[e] = pj;
          pv = fmaf(pj, vi[e], pv);
        }
        pv = bsum(pv, s.red);
#pragma unroll
        for (int e = 0; e < NE; ++e) {
          const int j = tid + e * NT;
          if (j < LD) s.W[m][j] = p[e] - (0.5f * tau * pv) * vi[e];
        }
        __syncthreads();
      }
      // ---- trailing update of the panel
      const int kl = k0 + nb - 1, kn = k0 + PB;
#pragma unroll
      for (int e = 0; e < NE; ++e) {
        const int j = tid + e * NT;
        if (j < LD) {
          float vv[PB], ww[PB];
#pragma unroll
          for (int mm = 0; mm < PB; ++mm) {
            vv[mm] = (mm < nb) ? s.V[mm][j] : 0.f;
            ww[mm] = (mm < nb) ? s.W[mm][j] : 0.f;
          }
          *reinterpret_cast<float4*>(&s.Vt[j][0]) = make_float4(vv[0], vv[1], vv[2], vv[3]);
          *reinterpret_cast<float4*>(&s.Vt[j][4]) = make_float4(vv[4], vv[5], vv[6], vv[7]);
          *reinterpret_cast<float4*>(&s.Wt[j][0]) = make_float4(ww[0], ww[1], ww[2], ww[3]);
          *reinterpret_cast<float4*>(&s.Wt[j][4]) = make_float4(ww[4], ww[5], ww[6], ww[7]);
          if (nb < PB) {
#pragma unroll
            for (int mm = 0; mm < PB; ++mm)
              if (mm >= nb) {
                s.V[mm][j] = 0.f;
                s.W[mm][j] = 0.f;
              }
          }
        }
      }
      __syncthreads();
      {
        TileIter it;
        it.init(kl, wib);
        while (it.next()) {
          if (CB * it.C + CB - 1 > RG * it.J) up_tile<true>(A, s, it.J, it.C, kl, kn, lane);
          else up_tile<false>(A, s, it.J, it.C, kl, kn, lane);
        }
      }
      __syncthreads();
    }
    if (tid == 0) {
      dd[N - 2] = (double)__ldcg(A + (size_t)(N - 2) * LD + (N - 2));
      dd[N - 1] = (double)__ldcg(A + (size_t)(N - 1) * LD + (N - 1));
      ee[N - 2] = (double)__ldcg(A + (size_t)(N - 1) * LD + (N - 2));
      tt[N - 2] = 0.0;
    }
  }
}

}  // namespace klt

int siib_launch_tridiag32(const SiibBuffers& b, const SiibEigBuffers& eb, int n, int rank_lo, cudaStream_t s) {
  static const bool attr = [] {
    cudaFuncSetAttribute(klt::tridiag32_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(klt::Smem));
    return true;
  }();
  (void)attr;
  klt::tridiag32_kernel<<<n, klt::NT, sizeof(klt::Smem), s>>>(b, eb, rank_lo, n);
  return 1;
}

}  // namespace nele
