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

constexpr int N = 420, LD = 448, NT = 384, NW = NT / 32, PB = 8, RG = 16, CB = 64;
constexpr int NE = (LD + NT - 1) / NT;  // vector elements per thread (j = tid, tid + NT)
constexpr int LC = 424;                 // pitch of the panel-column buffer (rows < N only)
constexpr int JMAX = (N - 1) / RG;      // last row group

struct Smem {
  float V[PB][LD];     // Householder vectors of the panel
  float W[PB][LD];
  float col[PB][LC];   // the panel's columns as of the start of the panel (rows >= column index)
  float pw[NW][LD];    // per-warp partial products
  union {
    struct {
      float Vt[LD][PB];  // V, W transposed for the trailing update (one 32-byte row per matrix row)
      float Wt[LD][PB];
    };
    float tile[NW][RG][CB];  // matvec phase: the tile each warp has in flight (cp.async), 4 KB per warp
  };
  float diag[LD];      // diagonal of the matrix as of the start of the panel
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

// Tiles (J, C) of the lower triangle with rows > k, columns > k: row groups J = jmin .. 26 with jmin = (k + 1) / 16,
// column blocks C = jmin / 4 .. J / 4.  The list depends on jmin only, so the 27 lists live in constant memory
// (one byte per tile: J << 3 | C, at most 112 tiles), written by tile_tables_ready(); warp w takes entries w, w + NW, ...
// (the first version enumerated the tiles with a stateful iterator: 11 % of the kernel's instructions).
constexpr int kMaxTiles = 112;
__constant__ unsigned char c_tiles[JMAX + 1][kMaxTiles];
__constant__ int c_ntiles[JMAX + 1];

// one tile of p = A v (A symmetric, lower triangle stored).  v: shared-memory vector (zero for indices <= k and
// >= N), pw: this warp's partial result.
// Every tile loads through one base pointer with immediate row offsets and no predicates.  What a tile covers
// beyond the trailing lower triangle is harmless by construction: rows <= k and columns <= k meet v = 0 (their
// products vanish, their own results are never read) and hold finite stale entries; the pad rows 420 .. 431 are zero;
// and inside the 64 x 64 diagonal blocks the part right of the diagonal is kept at zero (written once at the start,
// never by the update), so a diagonal tile is an ordinary tile whose only flaw is that A[j][j] v[j] enters p[j] twice
// (column part and row part) -- the caller subtracts it once (Smem::diag).  The first version masked every element of
// the diagonal tiles: 422 instead of 188 instructions for 30 % of the tiles.
// start the copy of tile (J, C) into this warp's shared-memory slot: 256 chunks of 16 bytes, eight per lane
__device__ __forceinline__ void tile_fetch(const float* __restrict__ A, float (*tile)[CB], int J, int C, int lane) {
  const float* __restrict__ src = A + (size_t)(RG * J) * LD + CB * C;
#pragma unroll
  for (int i = 0; i < RG * CB / 4 / 32; ++i) {
    const int id = lane + 32 * i, row = id >> 4, c4 = id & 15;
    const unsigned sa = (unsigned)__cvta_generic_to_shared(&tile[row][4 * c4]);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(sa), "l"(src + (size_t)row * LD + 4 * c4) : "memory");
  }
  asm volatile("cp.async.commit_group;" ::: "memory");
}

__device__ __forceinline__ void mv_tile(const float2 (&a)[RG], const float* __restrict__ v, float* __restrict__ pw, int J, int C,
                                        int lane) {
  const int j0 = RG * J, c0 = CB * C + 2 * lane;
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
      t[r] = fmaf(a[r].x, vc.x, a[r].y * vc.y);
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
// (non-diagonal tiles also update the stale entries and pad rows they cover: nobody reads those)
template <bool DIAG>
__device__ __forceinline__ void up_tile(float* __restrict__ A, Smem& s, int J, int C, int kn, int lane) {
  const int j0 = RG * J, c0 = CB * C + 2 * lane;
  float2 vcx[PB], wcx[PB];
#pragma unroll
  for (int mm = 0; mm < PB; ++mm) {
    vcx[mm] = *reinterpret_cast<const float2*>(&s.V[mm][c0]);
    wcx[mm] = *reinterpret_cast<const float2*>(&s.W[mm][c0]);
  }
  const bool tocol = (c0 & ~7) == kn;
  // two halves of eight rows: the eight loads of a half are in flight together, and the tile's footprint stays
  // inside the 80 registers that two CTAs of 384 threads per SM allow
#pragma unroll 1
  for (int h = 0; h < RG; h += RG / 2) {
    float2 a[RG / 2];
#pragma unroll
    for (int r = 0; r < RG / 2; ++r) {
      const int j = j0 + h + r;
      const bool ok = DIAG ? (j < N && c0 <= j) : true;
      a[r] = ok ? __ldcg(reinterpret_cast<const float2*>(A + (size_t)j * LD + c0)) : make_float2(0.f, 0.f);
    }
#pragma unroll
    for (int r = 0; r < RG / 2; ++r) {
      const int j = j0 + h + r;
      const bool ok = DIAG ? (j < N && c0 <= j) : true;
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
        if (DIAG) {
          if (c0 == j) s.diag[j] = ax;
          if (c0 + 1 == j) s.diag[j] = ay;
        }
        if (tocol && j < N) {
          s.col[c0 - kn][j] = ax;
          if (both) s.col[c0 + 1 - kn][j] = ay;
        }
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
    // lower triangle of Sxx as floats; the rest of each row up to the end of its 64-column diagonal block is zero
    for (int j = wib; j < N; j += NW) {
      const double2* src = reinterpret_cast<const double2*>(S + (int64_t)j * N);
      float2* dst = reinterpret_cast<float2*>(A + (size_t)j * LD);
      const int cend = CB * (j / CB) + CB;
      for (int c2 = lane; 2 * c2 < cend; c2 += 32) {
        float2 o = make_float2(0.f, 0.f);
        if (2 * c2 <= j) {
          const double2 v = src[c2];
          o.x = (float)v.x;
          if (2 * c2 + 1 <= j) o.y = (float)v.y;
        }
        dst[c2] = o;
      }
    }
    for (int j = tid; j < LD; j += NT) s.diag[j] = (j < N) ? (float)S[(int64_t)j * N + j] : 0.f;
    for (int idx = tid; idx < (RG * (JMAX + 1) - N) * LD; idx += NT) A[(size_t)N * LD + idx] = 0.f;   // pad rows 420 .. 431
    // columns of the first panel = rows of the symmetric input
    for (int idx = tid; idx < PB * LC; idx += NT) {
      const int q = idx / LC, j = idx % LC;
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
        // this warp's partial product vector starts at zero
#pragma unroll
        for (int i = 0; i < LD / 32; ++i) s.pw[wib][lane + 32 * i] = 0.f;
        __syncthreads();
        if (tau != 0.f) {
          // dot products of v with the panel's earlier vectors (for the corrections of p below): one warp -- the last,
          // which gets the fewest tiles -- does all 2 m of them while the others start on their tiles
          if (wib == NW - 1 && m > 0) {
            float vm[LD / 32];
#pragma unroll
            for (int i = 0; i < LD / 32; ++i) vm[i] = s.V[m][lane + 32 * i];
            for (int mm = 0; mm < m; ++mm) {
              float aw = 0.f, av = 0.f;
#pragma unroll
              for (int i = 0; i < LD / 32; ++i) {
                aw = fmaf(s.W[mm][lane + 32 * i], vm[i], aw);
                av = fmaf(s.V[mm][lane + 32 * i], vm[i], av);
              }
              aw = warp_sum(aw);
              av = warp_sum(av);
              if (lane == 0) {
                s.tot[2 * mm] = aw;
                s.tot[2 * mm + 1] = av;
              }
            }
          }
          // ---- p = A v over the trailing lower triangle
          // Each warp keeps one tile in flight: the copy of its next tile (cp.async, global -> its 4 KB slot) runs while
          // it multiplies the current one from registers.  Loading a tile straight into registers left the L2 latency
          // exposed once per tile (ncu: 2.8 of 9.8 stall cycles per instruction on the long scoreboard).
          const int jm = (k + 1) / RG, nt = c_ntiles[jm];
          if (wib < nt) {
            const int code = c_tiles[jm][wib];
            tile_fetch(A, s.tile[wib], code >> 3, code & 7, lane);
          }
          for (int t = wib; t < nt; t += NW) {
            const int code = c_tiles[jm][t], J = code >> 3, C = code & 7;
            asm volatile("cp.async.wait_group 0;" ::: "memory");
            __syncwarp();
            float2 a[RG];
#pragma unroll
            for (int r = 0; r < RG; ++r) a[r] = *reinterpret_cast<const float2*>(&s.tile[wib][r][2 * lane]);
            __syncwarp();
            if (t + NW < nt) {
              const int nc = c_tiles[jm][t + NW];
              tile_fetch(A, s.tile[wib], nc >> 3, nc & 7, lane);
            }
            mv_tile(a, s.V[m], s.pw[wib], J, C, lane);
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
            pj -= s.diag[j] * vi[e];   // counted by the column part and by the row part of its diagonal tile
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
        const int jm = (kl + 1) / RG, nt = c_ntiles[jm];
        for (int t = wib; t < nt; t += NW) {
          const int code = c_tiles[jm][t], J = code >> 3, C = code & 7;
          if (CB * C + CB - 1 > RG * J) up_tile<true>(A, s, J, C, kn, lane);
          else up_tile<false>(A, s, J, C, kn, lane);
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

// ------------------------------------------------------------------ quadratic forms
// rho_j = u_j^T Sxy u_j / sqrt(lambda_j u_j^T Syy u_j) for the r eigenvectors of a pair, as one register-tiled
// FP32 product per 64 eigenvectors: W = S^T G (420 x 420 x 64, Sxy and Syy side by side sharing the G
// fragments) with the epilogue a_j = sum_i G[i][j] W[i][j] fused, so W never exists.  G[j] = sqrt(lambda_j) u_j
// is what every KLT route leaves in b.G (column j contiguous); lambda_j = |G_j|^2 comes out of the same epilogue.
// Replaced round 1's siib_quad_kernel (thread = row, 16 eigenvectors per pass: Sxy / Syy re-read 27 times per pair,
// 36 GB of L2 traffic per 1024 pairs, 14.9 ms) -- here they are read 7 times by CTAs that run 64 FMAs per 5 shared
// loads, and only their folded lower triangles (u^T S u = u^T F u, F[c][i] = S[c][i] + S[i][c] for c > i).
//   CTA = (64-eigenvector tile, pair), 256 threads = 16 (rows) x 16 (columns), thread tile 8 x 4 x 2 matrices,
//   k-steps of 16 through double-buffered shared memory, next tile prefetched into registers.
namespace qf {

constexpr int N = 420, LD = 448, BM = 128, BN = 64, BK = 16, NT = 256;
constexpr int KT = (N + BK - 1) / BK;   // 27 k-tiles
constexpr int IB = (N + BM - 1) / BM;   // 4 row blocks
constexpr int JT = (N + BN - 1) / BN;   // 7 eigenvector tiles

struct Smem {
  float Axy[2][BK][BM];
  float Ayy[2][BK][BM];
  float B[2][BK][BN];
  float red[3][BN][17];
  double wsum[NT / 32];
};

__device__ __forceinline__ bool pair_scored_here(const SiibBuffers& b, int pair, int& r) {
  const int Fa = b.Fa[pair];
  const int Nf = Fa - 14;
  if (b.M[pair] <= 0 || (double)Fa / 80.0 < 20.0 || Nf < 2) return false;   // too short: quad_finish reports it
  const int P = b.Pact[pair];
  if (!b.no_proj && b.perflag[2 * pair] && b.perflag[2 * pair + 1] && P > 0 && Nf >= 2 * P) return false;  // siib_projquad_kernel
  r = b.rank[pair];
  return true;
}

__global__ void __launch_bounds__(NT, 2) quadform_kernel(SiibBuffers b, double* __restrict__ info_part) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  Smem& s = *reinterpret_cast<Smem*>(smem_raw);
  const int lp = blockIdx.y, pair = b.pair_lo + lp, jt = blockIdx.x, j0 = jt * BN, tid = threadIdx.x;
  const int ti = tid >> 4, tj = tid & 15, lane = tid & 31, wib = tid >> 5;
  int r = 0;
  if (!pair_scored_here(b, pair, r)) return;
  float* __restrict__ lam_out = b.lambda + (int64_t)pair * N;
  float* __restrict__ rho_out = b.rho + (int64_t)pair * N;
  if (j0 >= r) {  // nothing beyond the rank
    if (tid < BN && j0 + tid < N) {
      lam_out[j0 + tid] = 0.f;
      rho_out[j0 + tid] = 0.f;
    }
    if (tid == 0) info_part[(int64_t)pair * 8 + jt] = 0.0;
    return;
  }
  const float* __restrict__ G = b.G + (int64_t)lp * N * LD;
  const float* __restrict__ Sxy = b.Sxy + (int64_t)lp * N * N;
  const float* __restrict__ Syy = b.Syy + (int64_t)lp * N * N;
  // global -> register staging: A tiles 16 x 128 (two float4 per thread and matrix), B tile 16 x 64 (one float4)
  const int arow = tid >> 5, acol = 4 * (tid & 31);
  const int bj = tid >> 2, bc = 4 * (tid & 3);
  const bool bj_ok = j0 + bj < r;
  float4 rxy[2], ryy[2], rb;
  // Sxy and Syy arrive folded onto the lower triangle (siib_expand_kernel: F[c][i] = S[c][i] + S[i][c], c > i), so a
  // row block only visits the k-tiles from its own first row on, and in the eight tiles that cross the diagonal the
  // elements above it (never written) are replaced by zeros
  auto fetch = [&](int i0, int kt) {
    const int c0 = kt * BK;
    const bool crosses = c0 < i0 + BM;
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int c = c0 + arow + 8 * h, i = i0 + acol;
      const bool ok = c < N && i < N && c >= i;
      rxy[h] = ok ? __ldg(reinterpret_cast<const float4*>(Sxy + (int64_t)c * N + i)) : make_float4(0.f, 0.f, 0.f, 0.f);
      ryy[h] = ok ? __ldg(reinterpret_cast<const float4*>(Syy + (int64_t)c * N + i)) : make_float4(0.f, 0.f, 0.f, 0.f);
      if (crosses) {
        if (c < i + 1) rxy[h].y = ryy[h].y = 0.f;
        if (c < i + 2) rxy[h].z = ryy[h].z = 0.f;
        if (c < i + 3) rxy[h].w = ryy[h].w = 0.f;
      }
    }
    const int c = c0 + bc;
    rb = (bj_ok && c < N) ? __ldg(reinterpret_cast<const float4*>(G + (int64_t)(j0 + bj) * LD + c)) : make_float4(0.f, 0.f, 0.f, 0.f);
  };
  auto stash = [&](int buf) {
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      *reinterpret_cast<float4*>(&s.Axy[buf][arow + 8 * h][acol]) = rxy[h];
      *reinterpret_cast<float4*>(&s.Ayy[buf][arow + 8 * h][acol]) = ryy[h];
    }
    s.B[buf][bc][bj] = rb.x;
    s.B[buf][bc + 1][bj] = rb.y;
    s.B[buf][bc + 2][bj] = rb.z;
    s.B[buf][bc + 3][bj] = rb.w;
  };
  float pa[4] = {0.f, 0.f, 0.f, 0.f}, pc[4] = {0.f, 0.f, 0.f, 0.f}, pl[4] = {0.f, 0.f, 0.f, 0.f};
  for (int ib = 0; ib < IB; ++ib) {
    const int i0 = ib * BM;
    // accumulators as f32x2 pairs of rows (fma.rn.f32x2: two FMAs per issue slot, the same FMA-pipe rate): the
    // scalar version of this loop sat at 73 % issue-slot and 61 % FMA-pipe utilisation, i.e. issue bound
    F2 axy[4][4], ayy[4][4];
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
      for (int c = 0; c < 4; ++c) axy[a][c] = ayy[a][c] = f2_pack(0.f, 0.f);
    __syncthreads();
    const int kt0 = i0 / BK;   // BM is a multiple of BK
    fetch(i0, kt0);
    stash(kt0 & 1);
    __syncthreads();
    for (int kt = kt0; kt < KT; ++kt) {
      const int cur = kt & 1;
      if (kt + 1 < KT) fetch(i0, kt + 1);
#pragma unroll
      for (int k = 0; k < BK; ++k) {
        const float4 x0 = *reinterpret_cast<const float4*>(&s.Axy[cur][k][8 * ti]);
        const float4 x1 = *reinterpret_cast<const float4*>(&s.Axy[cur][k][8 * ti + 4]);
        const float4 y0 = *reinterpret_cast<const float4*>(&s.Ayy[cur][k][8 * ti]);
        const float4 y1 = *reinterpret_cast<const float4*>(&s.Ayy[cur][k][8 * ti + 4]);
        const float4 bq = *reinterpret_cast<const float4*>(&s.B[cur][k][4 * tj]);
        const F2 xa[4] = {f2_pack(x0.x, x0.y), f2_pack(x0.z, x0.w), f2_pack(x1.x, x1.y), f2_pack(x1.z, x1.w)};
        const F2 ya[4] = {f2_pack(y0.x, y0.y), f2_pack(y0.z, y0.w), f2_pack(y1.x, y1.y), f2_pack(y1.z, y1.w)};
        const F2 bb[4] = {f2_pack(bq.x, bq.x), f2_pack(bq.y, bq.y), f2_pack(bq.z, bq.z), f2_pack(bq.w, bq.w)};
#pragma unroll
        for (int a = 0; a < 4; ++a)
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            axy[a][c] = f2_fma(xa[a], bb[c], axy[a][c]);
            ayy[a][c] = f2_fma(ya[a], bb[c], ayy[a][c]);
          }
      }
      if (kt + 1 < KT) stash(cur ^ 1);
      __syncthreads();
    }
    // epilogue of the row block: a_j += sum_i G[i][j] W[i][j], likewise yy and |G_j|^2
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      const int j = j0 + 4 * tj + c;
      if (j < r) {
        const int i = i0 + 8 * ti;
        const float4 g0 = (i < N) ? __ldg(reinterpret_cast<const float4*>(G + (int64_t)j * LD + i)) : make_float4(0.f, 0.f, 0.f, 0.f);
        const float4 g1 = (i + 4 < N) ? __ldg(reinterpret_cast<const float4*>(G + (int64_t)j * LD + i + 4)) : make_float4(0.f, 0.f, 0.f, 0.f);
        const float gg[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
#pragma unroll
        for (int a = 0; a < 4; ++a) {
          float w0, w1, z0, z1;
          f2_unpack(axy[a][c], w0, w1);
          f2_unpack(ayy[a][c], z0, z1);
          pa[c] = fmaf(w0, gg[2 * a], pa[c]);
          pa[c] = fmaf(w1, gg[2 * a + 1], pa[c]);
          pc[c] = fmaf(z0, gg[2 * a], pc[c]);
          pc[c] = fmaf(z1, gg[2 * a + 1], pc[c]);
          pl[c] = fmaf(gg[2 * a], gg[2 * a], pl[c]);
          pl[c] = fmaf(gg[2 * a + 1], gg[2 * a + 1], pl[c]);
        }
      }
    }
  }
  // sums over the 16 row groups, then rho and the information of the tile's components
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    s.red[0][4 * tj + c][ti] = pa[c];
    s.red[1][4 * tj + c][ti] = pc[c];
    s.red[2][4 * tj + c][ti] = pl[c];
  }
  __syncthreads();
  double info = 0.0;
  if (tid < BN) {
    const int j = j0 + tid;
    double a = 0.0, c = 0.0, l = 0.0;
#pragma unroll
    for (int q = 0; q < 16; ++q) {
      a += (double)s.red[0][tid][q];
      c += (double)s.red[1][tid][q];
      l += (double)s.red[2][tid][q];
    }
    double rho = 0.0;
    if (j < r && l > 0.0 && c > 0.0) rho = a / (l * sqrt(c));
    rho = fmin(1.0, fmax(-1.0, rho));
    const double pr = 0.75 * rho;
    info = -0.5 * log2(1.0 - pr * pr);
    if (j < N) {
      lam_out[j] = (j < r) ? (float)l : 0.f;
      rho_out[j] = (float)rho;
    }
  }
  if (tid < 64) {   // two warps, fixed summation order
    info = warp_sum(info);
    if (lane == 0) s.wsum[wib] = info;
  }
  __syncthreads();
  if (tid == 0) info_part[(int64_t)pair * 8 + jt] = s.wsum[0] + s.wsum[1];
}

// score and status of the pairs scored by the quadratic forms (the others: siib_projquad_kernel)
__global__ void quad_finish_kernel(SiibBuffers b, const double* __restrict__ info_part, int n) {
  const int lp = blockIdx.x * blockDim.x + threadIdx.x;
  if (lp >= n) return;
  const int pair = b.pair_lo + lp;
  int r = 0;
  if (!pair_scored_here(b, pair, r)) {
    const int Fa = b.Fa[pair];
    if (b.M[pair] <= 0 || (double)Fa / 80.0 < 20.0 || Fa - 14 < 2) {  // pysiib: "at least 20 seconds of speech"
      b.score[pair] = nan("");
      b.status[pair] = 2;
    }
    return;
  }
  double sum = 0.0;
  for (int jt = 0; jt < JT; ++jt) sum += info_part[(int64_t)pair * 8 + jt];
  const double v = 16000.0 / 200.0 / 15.0 * sum;
  b.score[pair] = v > 0.0 ? v : 0.0;
  // bit 8: the null space of a rank-deficient (exactly periodic) covariance was given zero information
  // (NELE_INFO_SIIB_NULLSPACE).  sweeps: -1 tridiagonal path at full rank, -3 rank deficient, -2 Gram path.
  const int sw = b.sweeps[pair];
  b.status[pair] = (sw == -2 || sw == -3 || (sw >= 0 && r < N)) ? 0x100 : 0;
}

}  // namespace qf

int siib_launch_quadform(const SiibBuffers& b, double* info_part, int n, KernelTimer* kt, cudaStream_t s) {
  static const bool attr = [] {
    cudaFuncSetAttribute(qf::quadform_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(qf::Smem));
    return true;
  }();
  (void)attr;
  kt_begin(kt, "siib_quad", s);
  qf::quadform_kernel<<<dim3(qf::JT, n), qf::NT, sizeof(qf::Smem), s>>>(b, info_part);
  kt_end(kt, s);
  kt_begin(kt, "siib_quad_finish", s);
  qf::quad_finish_kernel<<<(n + 127) / 128, 128, 0, s>>>(b, info_part, n);
  kt_end(kt, s);
  return 2;
}

// ------------------------------------------------------------------ back-transformation
// u_j = H_0 ... H_{n-3} z_j with the vector in registers (lane = eigenvector, 8 warps = 8 interleaved row parts), as
// siib_backtf_kernel, but the reflectors are applied four at a time: one pass forms the four dot products
// d_q = v_q . u, the coefficients follow from the 4 x 4 triangular recurrence
//   y_0 = tau_0 d_0,  y_q = tau_q (d_q - sum_{p<q} y_p (v_q . v_p))
// (the Gram values v_q . v_p are computed once per panel while it is staged), and one pass applies
// u -= sum_q y_q v_q.  Same flops, but one block barrier per four reflectors instead of one per reflector: the
// old kernel ran at a quarter of its issue rate waiting on that barrier (15.1 ms per 1024 pairs).
namespace bt {

constexpr int N = 420, LD = 448, VEC = 32, PARTS = 8, NT = PARTS * 32, ROWS = LD / PARTS, PANEL = 16, GRP = 4;

__device__ __forceinline__ void cp_async4(float* smem, const float* gmem) {
  const unsigned sa = (unsigned)__cvta_generic_to_shared(smem);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(sa), "l"(gmem) : "memory");
}

// The reflectors of the next panel travel global -> shared memory with cp.async (element-wise: the staged layout
// interleaves the rows over the eight warps) while the current panel is applied; the first version staged a panel
// through registers between two barriers and spent most of its time on that exposed L2 latency (ncu: 5.6 cycles of
// long-scoreboard stall per issued instruction, 24.5 ms per 1024 pairs).
__global__ void __launch_bounds__(NT, 2) backtf4_kernel(SiibBuffers b, SiibEigBuffers eb, int rank_lo) {
  const int lp = blockIdx.y, pair = b.pair_lo + lp, tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  const int rk = b.rank[pair];
  if (rk < rank_lo) return;
  const int j = blockIdx.x * VEC + lane;
  const bool live = j < N;
  extern __shared__ __align__(16) float s_vbuf[];   // [2][PANEL][LD]: reflector kk of a panel, staged as [row part w][m]: row 8 m + w
  __shared__ float s_tau[PANEL];
  __shared__ float s_g[PANEL / GRP][8];            // per group: v1.v0, v2.v0, v2.v1, v3.v0, v3.v1, v3.v2
  __shared__ float s_dot[2][PARTS][GRP][32];
  const float* __restrict__ R = eb.refl + (int64_t)lp * N * LD;
  const double* __restrict__ tt = eb.tau + (int64_t)lp * LD;
  auto stage = [&](int k1, int buf) {
    const int nk = min(PANEL, k1 + 1);
    float* dst = s_vbuf + (size_t)buf * PANEL * LD;
    for (int kk = 0; kk < PANEL; ++kk) {
      if (kk < nk) {
        const float* __restrict__ row = R + (int64_t)(k1 - kk) * LD;
#pragma unroll
        for (int i = tid; i < LD; i += NT) cp_async4(dst + kk * LD + (i & (PARTS - 1)) * ROWS + (i >> 3), row + i);
      } else {
        for (int i = tid; i < LD; i += NT) dst[kk * LD + i] = 0.f;
      }
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };
  stage(N - 3, 0);
  float u[ROWS];
#pragma unroll
  for (int m = 0; m < ROWS; ++m) {
    const int i = PARTS * m + w;
    u[m] = (live && i < N) ? eb.zt[((int64_t)lp * N + i) * LD + j] : 0.f;
  }
  int par = 0, buf = 0;
  for (int k1 = N - 3; k1 >= 0; k1 -= PANEL, buf ^= 1) {   // reflectors k1, k1 - 1, ... (descending inside the panel)
    const int nk = min(PANEL, k1 + 1);
    __syncthreads();                                        // everyone is done with the other buffer (previous panel)
    if (k1 - PANEL >= 0) {
      stage(k1 - PANEL, buf ^ 1);
      asm volatile("cp.async.wait_group 1;" ::: "memory");
    } else {
      asm volatile("cp.async.wait_group 0;" ::: "memory");
    }
    if (tid < PANEL) s_tau[tid] = (tid < nk) ? (float)tt[k1 - tid] : 0.f;
    __syncthreads();
    float (*s_v)[LD] = reinterpret_cast<float (*)[LD]>(s_vbuf + (size_t)buf * PANEL * LD);
    // Gram values inside each group of four: 24 dot products of length 448, three per warp
    for (int d = w; d < (PANEL / GRP) * 6; d += PARTS) {
      const int g = d / 6, e = d % 6;
      const int qa = (e == 0) ? 1 : (e < 3) ? 2 : 3;
      const int qb = (e == 0) ? 0 : (e == 1) ? 0 : (e == 2) ? 1 : e - 3;
      const float* va = s_v[GRP * g + qa];
      const float* vb = s_v[GRP * g + qb];
      float acc = 0.f;
#pragma unroll
      for (int i = 0; i < LD / 32; ++i) acc = fmaf(va[lane + 32 * i], vb[lane + 32 * i], acc);
      acc = warp_sum(acc);
      if (lane == 0) s_g[g][e] = acc;
    }
    __syncthreads();
    for (int g = 0; g < PANEL / GRP; ++g) {
      const int kk0 = GRP * g;
      if (kk0 >= nk) break;
      // every reflector of the group is zero in rows <= kmin, i.e. in the words q < kmin / 32 of every warp
      const int kmin = max(k1 - kk0 - (GRP - 1), 0);
      const int q0 = kmin >> 5;
      const float4* v0 = reinterpret_cast<const float4*>(s_v[kk0] + w * ROWS);
      const float4* v1 = reinterpret_cast<const float4*>(s_v[kk0 + 1] + w * ROWS);
      const float4* v2 = reinterpret_cast<const float4*>(s_v[kk0 + 2] + w * ROWS);
      const float4* v3 = reinterpret_cast<const float4*>(s_v[kk0 + 3] + w * ROWS);
      float d0 = 0.f, d1 = 0.f, d2 = 0.f, d3 = 0.f;
#pragma unroll
      for (int q = 0; q < ROWS / 4; ++q) {
        if (q < q0) continue;
        const float4 a = v0[q], bq = v1[q], c = v2[q], e = v3[q];
        d0 = fmaf(a.x, u[4 * q], d0); d0 = fmaf(a.y, u[4 * q + 1], d0); d0 = fmaf(a.z, u[4 * q + 2], d0); d0 = fmaf(a.w, u[4 * q + 3], d0);
        d1 = fmaf(bq.x, u[4 * q], d1); d1 = fmaf(bq.y, u[4 * q + 1], d1); d1 = fmaf(bq.z, u[4 * q + 2], d1); d1 = fmaf(bq.w, u[4 * q + 3], d1);
        d2 = fmaf(c.x, u[4 * q], d2); d2 = fmaf(c.y, u[4 * q + 1], d2); d2 = fmaf(c.z, u[4 * q + 2], d2); d2 = fmaf(c.w, u[4 * q + 3], d2);
        d3 = fmaf(e.x, u[4 * q], d3); d3 = fmaf(e.y, u[4 * q + 1], d3); d3 = fmaf(e.z, u[4 * q + 2], d3); d3 = fmaf(e.w, u[4 * q + 3], d3);
      }
      s_dot[par][w][0][lane] = d0;
      s_dot[par][w][1][lane] = d1;
      s_dot[par][w][2][lane] = d2;
      s_dot[par][w][3][lane] = d3;
      __syncthreads();
      float D0 = 0.f, D1 = 0.f, D2 = 0.f, D3 = 0.f;
#pragma unroll
      for (int t = 0; t < PARTS; ++t) {
        D0 += s_dot[par][t][0][lane];
        D1 += s_dot[par][t][1][lane];
        D2 += s_dot[par][t][2][lane];
        D3 += s_dot[par][t][3][lane];
      }
      par ^= 1;
      const float y0 = s_tau[kk0] * D0;
      const float y1 = s_tau[kk0 + 1] * (D1 - y0 * s_g[g][0]);
      const float y2 = s_tau[kk0 + 2] * (D2 - y0 * s_g[g][1] - y1 * s_g[g][2]);
      const float y3 = s_tau[kk0 + 3] * (D3 - y0 * s_g[g][3] - y1 * s_g[g][4] - y2 * s_g[g][5]);
#pragma unroll
      for (int q = 0; q < ROWS / 4; ++q) {
        if (q < q0) continue;
        const float4 a = v0[q], bq = v1[q], c = v2[q], e = v3[q];
        u[4 * q] -= y0 * a.x + y1 * bq.x + y2 * c.x + y3 * e.x;
        u[4 * q + 1] -= y0 * a.y + y1 * bq.y + y2 * c.y + y3 * e.y;
        u[4 * q + 2] -= y0 * a.z + y1 * bq.z + y2 * c.z + y3 * e.z;
        u[4 * q + 3] -= y0 * a.w + y1 * bq.w + y2 * c.w + y3 * e.w;
      }
    }
  }
  if (!live) return;
  // column j of G = sqrt(lambda_j) u_j / |z_j|.  Eigenvalues at or below 1e-10 lambda_max carry no information; of a
  // periodic tiling of rank r < 420 (pivoted Cholesky, FP64) only the r largest are non-zero in exact arithmetic --
  // the FP32 tridiagonalisation leaves the others at +-1e-7 lambda_max, so there the rank decides, not the value
  const double lam = eb.lam[(int64_t)lp * LD + j], lmax = eb.lam[(int64_t)lp * LD + N - 1];
  const double nz = eb.znorm[(int64_t)lp * LD + j];
  const bool in_range = j >= N - rk;
  const float sc = (in_range && lam > 1.0e-10 * lmax && nz > 0.0) ? (float)sqrt(lam / nz) : 0.f;
  float* __restrict__ G = b.G + (int64_t)lp * N * LD + (int64_t)j * LD;
#pragma unroll
  for (int m = 0; m < ROWS; ++m) {
    const int i = PARTS * m + w;
    G[i] = (i < N) ? sc * u[m] : 0.f;
  }
}

}  // namespace bt

// ---- the same back-transformation with a 14-row x 4-vector register tile per lane
// backtf4 keeps one eigenvector per lane, so every reflector value is a warp-wide broadcast used for a single FMA:
// 16 bytes x 32 lanes returned per LDS.128 for four FMAs, and ncu shows the shared-memory pipe at 75 % with the FMA
// pipe at 28 %.  Here a lane owns rows {32 m + rp : m < 14} of four vectors (rp = one of 32 row parts, four per warp;
// vector group = lane / 4): a loaded reflector value feeds four FMAs, 3.5x fewer shared-memory bytes per flop.  The 16
// partial dot products of a group (4 reflectors x 4 vectors) are summed over the row parts by two shuffle levels inside
// the warp and through shared memory across the eight warps.
namespace bt5 {

constexpr int N = 420, LD = 448, VEC = 32, NW = 8, NT = NW * 32, RP = 32, M = 16, MV = 14, C = 4, PANEL = 16, GRP = 4;
constexpr int MS = 20;        // shared-memory stride of a row part: with 16 the four parts of a warp fall on two bank groups
                              // (ncu: 3.6-way conflicts on the reflector loads); 20 spreads them over banks 0, 20, 8, 28
constexpr int SLD = RP * MS;  // staged length of one reflector

__global__ void __launch_bounds__(NT, 2) backtf5_kernel(SiibBuffers b, SiibEigBuffers eb, int rank_lo) {
  const int lp = blockIdx.y, pair = b.pair_lo + lp, tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  const int rk = b.rank[pair];
  if (rk < rank_lo) return;
  const int rpl = lane & 3, vg = lane >> 2, rp = 4 * w + rpl;
  const int j0 = blockIdx.x * VEC + C * vg;
  extern __shared__ __align__(16) float s_vbuf[];   // [2][PANEL][SLD]: reflector kk staged as [row part][MS]: row 32 m + rp at rp * MS + m
  __shared__ float s_tau[PANEL];
  __shared__ float s_g[PANEL / GRP][8];
  __shared__ __align__(16) float s_dot[2][NW][VEC / C][GRP * C];
  const float* __restrict__ R = eb.refl + (int64_t)lp * N * LD;
  const double* __restrict__ tt = eb.tau + (int64_t)lp * LD;
  for (int i = tid; i < 2 * PANEL * SLD; i += NT) s_vbuf[i] = 0.f;   // the pad rows (m = 14, 15) stay zero for good
  __syncthreads();
  auto stage = [&](int k1, int buf) {
    const int nk = min(PANEL, k1 + 1);
    float* dst = s_vbuf + (size_t)buf * PANEL * SLD;
    for (int kk = 0; kk < PANEL; ++kk) {
      if (kk < nk) {
        const float* __restrict__ row = R + (int64_t)(k1 - kk) * LD;
#pragma unroll
        for (int i = tid; i < LD; i += NT) bt::cp_async4(dst + kk * SLD + (i & (RP - 1)) * MS + (i >> 5), row + i);
      } else {
        for (int i = tid; i < LD; i += NT) dst[kk * SLD + (i & (RP - 1)) * MS + (i >> 5)] = 0.f;
      }
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };
  stage(N - 3, 0);
  // rows in pairs (m, m + 1) packed for fma.rn.f32x2: a float4 of a staged reflector is two such pairs, so every
  // multiply-add of the two passes is two-wide with no packing moves (only the 16 coefficients y are duplicated)
  F2 u[MV / 2][C];
#pragma unroll
  for (int mp = 0; mp < MV / 2; ++mp) {
    const int i0 = RP * (2 * mp) + rp, i1 = i0 + RP;
#pragma unroll
    for (int c = 0; c < C; ++c) {
      const float lo = (i0 < N && j0 + c < N) ? eb.zt[((int64_t)lp * N + i0) * LD + j0 + c] : 0.f;
      const float hi = (i1 < N && j0 + c < N) ? eb.zt[((int64_t)lp * N + i1) * LD + j0 + c] : 0.f;
      u[mp][c] = f2_pack(lo, hi);
    }
  }
  int par = 0, buf = 0;
  for (int k1 = N - 3; k1 >= 0; k1 -= PANEL, buf ^= 1) {
    const int nk = min(PANEL, k1 + 1);
    __syncthreads();
    if (k1 - PANEL >= 0) {
      stage(k1 - PANEL, buf ^ 1);
      asm volatile("cp.async.wait_group 1;" ::: "memory");
    } else {
      asm volatile("cp.async.wait_group 0;" ::: "memory");
    }
    if (tid < PANEL) s_tau[tid] = (tid < nk) ? (float)tt[k1 - tid] : 0.f;
    __syncthreads();
    const float* s_v = s_vbuf + (size_t)buf * PANEL * SLD;
    // Gram values inside each group of four reflectors: 24 dot products, three per warp
    for (int d = w; d < (PANEL / GRP) * 6; d += NW) {
      const int g = d / 6, e = d % 6;
      const int qa = (e == 0) ? 1 : (e < 3) ? 2 : 3;
      const int qb = (e == 0) ? 0 : (e == 1) ? 0 : (e == 2) ? 1 : e - 3;
      const float* va = s_v + (GRP * g + qa) * SLD;
      const float* vb = s_v + (GRP * g + qb) * SLD;
      float acc = 0.f;
#pragma unroll
      for (int i = 0; i < SLD / 32; ++i) acc = fmaf(va[lane + 32 * i], vb[lane + 32 * i], acc);
      acc = warp_sum(acc);
      if (lane == 0) s_g[g][e] = acc;
    }
    __syncthreads();
    for (int g = 0; g < PANEL / GRP; ++g) {
      const int kk0 = GRP * g;
      if (kk0 >= nk) break;
      // rows 32 m + rp <= kmin are zero in every reflector of the group: row pairs mp < (kmin + 1) / 64 are skipped
      const int kmin = max(k1 - kk0 - (GRP - 1), 0);
      const int p0 = (kmin + 1) >> 6;
      const float4* vq[GRP];
#pragma unroll
      for (int q = 0; q < GRP; ++q) vq[q] = reinterpret_cast<const float4*>(s_v + (kk0 + q) * SLD + rp * MS);
      F2 d2[GRP][C];
#pragma unroll
      for (int q = 0; q < GRP; ++q)
#pragma unroll
        for (int c = 0; c < C; ++c) d2[q][c] = f2_pack(0.f, 0.f);
#pragma unroll
      for (int t = 0; t < M / 4; ++t) {
        if (2 * t + 1 < p0) continue;
#pragma unroll
        for (int q = 0; q < GRP; ++q) {
          const float4 a4 = vq[q][t];
          const F2 a2[2] = {f2_pack(a4.x, a4.y), f2_pack(a4.z, a4.w)};
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            const int mp = 2 * t + h;
            if (mp >= MV / 2 || mp < p0) continue;
#pragma unroll
            for (int c = 0; c < C; ++c) d2[q][c] = f2_fma(a2[h], u[mp][c], d2[q][c]);
          }
        }
      }
      float d[GRP][C];
#pragma unroll
      for (int q = 0; q < GRP; ++q)
#pragma unroll
        for (int c = 0; c < C; ++c) {
          float lo, hi;
          f2_unpack(d2[q][c], lo, hi);
          d[q][c] = lo + hi;
        }
      // sum over the four row parts of the warp, then over the warps through shared memory
#pragma unroll
      for (int q = 0; q < GRP; ++q)
#pragma unroll
        for (int c = 0; c < C; ++c) {
          d[q][c] += __shfl_xor_sync(0xffffffffu, d[q][c], 1);
          d[q][c] += __shfl_xor_sync(0xffffffffu, d[q][c], 2);
        }
      if (rpl == 0) {
#pragma unroll
        for (int q = 0; q < GRP; ++q) *reinterpret_cast<float4*>(&s_dot[par][w][vg][q * C]) = make_float4(d[q][0], d[q][1], d[q][2], d[q][3]);
      }
      __syncthreads();
      // lane rpl of a vector group adds up reflector q = rpl over the warps; the four lanes then exchange the totals
      float4 mine = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
      for (int ww = 0; ww < NW; ++ww) {
        const float4 x = *reinterpret_cast<const float4*>(&s_dot[par][ww][vg][rpl * C]);
        mine.x += x.x;
        mine.y += x.y;
        mine.z += x.z;
        mine.w += x.w;
      }
      par ^= 1;
      float D[GRP][C];
#pragma unroll
      for (int q = 0; q < GRP; ++q) {
        const int src = (lane & ~3) | q;
        D[q][0] = __shfl_sync(0xffffffffu, mine.x, src);
        D[q][1] = __shfl_sync(0xffffffffu, mine.y, src);
        D[q][2] = __shfl_sync(0xffffffffu, mine.z, src);
        D[q][3] = __shfl_sync(0xffffffffu, mine.w, src);
      }
      float y[GRP][C];
      const float g10 = s_g[g][0], g20 = s_g[g][1], g21 = s_g[g][2], g30 = s_g[g][3], g31 = s_g[g][4], g32 = s_g[g][5];
      const float tau0 = s_tau[kk0], tau1 = s_tau[kk0 + 1], tau2 = s_tau[kk0 + 2], tau3 = s_tau[kk0 + 3];
#pragma unroll
      for (int c = 0; c < C; ++c) {
        y[0][c] = tau0 * D[0][c];
        y[1][c] = tau1 * (D[1][c] - y[0][c] * g10);
        y[2][c] = tau2 * (D[2][c] - y[0][c] * g20 - y[1][c] * g21);
        y[3][c] = tau3 * (D[3][c] - y[0][c] * g30 - y[1][c] * g31 - y[2][c] * g32);
      }
      F2 ny[GRP][C];
#pragma unroll
      for (int q = 0; q < GRP; ++q)
#pragma unroll
        for (int c = 0; c < C; ++c) ny[q][c] = f2_pack(-y[q][c], -y[q][c]);
#pragma unroll
      for (int t = 0; t < M / 4; ++t) {
        if (2 * t + 1 < p0) continue;
#pragma unroll
        for (int q = 0; q < GRP; ++q) {
          const float4 a4 = vq[q][t];
          const F2 a2[2] = {f2_pack(a4.x, a4.y), f2_pack(a4.z, a4.w)};
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            const int mp = 2 * t + h;
            if (mp >= MV / 2 || mp < p0) continue;
#pragma unroll
            for (int c = 0; c < C; ++c) u[mp][c] = f2_fma(ny[q][c], a2[h], u[mp][c]);
          }
        }
      }
    }
  }
  // column j of G = sqrt(lambda_j) u_j / |z_j| (see backtf4_kernel for the thresholds)
  const double lmax = eb.lam[(int64_t)lp * LD + N - 1];
#pragma unroll
  for (int c = 0; c < C; ++c) {
    const int j = j0 + c;
    if (j >= N) continue;
    const double lam = eb.lam[(int64_t)lp * LD + j], nz = eb.znorm[(int64_t)lp * LD + j];
    const bool in_range = j >= N - rk;
    const float sc = (in_range && lam > 1.0e-10 * lmax && nz > 0.0) ? (float)sqrt(lam / nz) : 0.f;
    float* __restrict__ G = b.G + (int64_t)lp * N * LD + (int64_t)j * LD;
#pragma unroll
    for (int mp = 0; mp < MV / 2; ++mp) {
      float lo, hi;
      f2_unpack(u[mp][c], lo, hi);
      const int i0 = RP * (2 * mp) + rp, i1 = i0 + RP;
      G[i0] = (i0 < N) ? sc * lo : 0.f;
      G[i1] = (i1 < N) ? sc * hi : 0.f;
    }
  }
}

}  // namespace bt5

int siib_launch_backtf5(const SiibBuffers& b, const SiibEigBuffers& eb, int n, int rank_lo, cudaStream_t s) {
  constexpr int smem = 2 * bt5::PANEL * bt5::SLD * (int)sizeof(float);
  static const bool attr = [] {
    cudaFuncSetAttribute(bt5::backtf5_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    return true;
  }();
  (void)attr;
  bt5::backtf5_kernel<<<dim3((bt5::N + bt5::VEC - 1) / bt5::VEC, n), bt5::NT, smem, s>>>(b, eb, rank_lo);
  return 1;
}

// ---- warp-independent form: a warp owns four eigenvectors over ALL rows
// backtf5 still meets at a block barrier once per group of four reflectors (its 32 row parts are spread over the eight
// warps), and with 16 warps per SM that barrier and the shared-memory exchange of the partial dot products leave the
// kernel latency bound (ncu: 54 % issue slots, FMA pipe 33 %, neither the shared-memory nor the FMA pipe saturated).
// Here lane l of a warp holds rows {64 p + 2 l, 64 p + 2 l + 1 : p < 7} of the warp's four vectors (the same 14 x 4
// register tile, as f32x2 row pairs): the dot products of a group are reduced inside the warp (a 16-value transposed
// shuffle reduction and a 16-shuffle all-gather), so warps never wait for each other except when the staged panel
// is swapped.  The staged reflectors keep their natural order: 16-byte cp.async, 64-bit conflict-free loads.
namespace bt6 {

constexpr int N = 420, LD = 448, NW = 8, NT = NW * 32, C = 4, VEC = NW * C, NP = 7, PANEL = 16, GRP = 4;

__global__ void __launch_bounds__(NT, 2) backtf6_kernel(SiibBuffers b, SiibEigBuffers eb, int rank_lo) {
  const int lp = blockIdx.y, pair = b.pair_lo + lp, tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  const int rk = b.rank[pair];
  if (rk < rank_lo) return;
  const int j0 = blockIdx.x * VEC + C * w;
  extern __shared__ __align__(16) float s_vbuf[];   // [2][PANEL][LD]
  __shared__ float s_tau[2][PANEL];
  __shared__ float s_g[2][PANEL / GRP][8];
  const float* __restrict__ R = eb.refl + (int64_t)lp * N * LD;
  const double* __restrict__ tt = eb.tau + (int64_t)lp * LD;
  auto stage = [&](int k1, int buf) {
    const int nk = min(PANEL, k1 + 1);
    float* dst = s_vbuf + (size_t)buf * PANEL * LD;
    for (int idx = tid; idx < PANEL * (LD / 4); idx += NT) {
      const int kk = idx / (LD / 4), i4 = idx % (LD / 4);
      float* d = dst + kk * LD + 4 * i4;
      if (kk < nk) {
        const unsigned sa = (unsigned)__cvta_generic_to_shared(d);
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(sa), "l"(R + (int64_t)(k1 - kk) * LD + 4 * i4) : "memory");
      } else {
        *reinterpret_cast<float4*>(d) = make_float4(0.f, 0.f, 0.f, 0.f);
      }
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };
  stage(N - 3, 0);
  F2 u[NP][C];
#pragma unroll
  for (int p = 0; p < NP; ++p) {
    const int i0 = 64 * p + 2 * lane;
#pragma unroll
    for (int c = 0; c < C; ++c) {
      const bool okj = j0 + c < N;
      const float lo = (okj && i0 < N) ? eb.zt[((int64_t)lp * N + i0) * LD + j0 + c] : 0.f;
      const float hi = (okj && i0 + 1 < N) ? eb.zt[((int64_t)lp * N + i0 + 1) * LD + j0 + c] : 0.f;
      u[p][c] = f2_pack(lo, hi);
    }
  }
  const unsigned full = 0xffffffffu;
  const bool b4 = lane & 16, b3 = lane & 8, b2 = lane & 4, b1 = lane & 2;
  int buf = 0;
  for (int k1 = N - 3; k1 >= 0; k1 -= PANEL, buf ^= 1) {
    const int nk = min(PANEL, k1 + 1);
    __syncthreads();                                        // every warp is done with the other buffer
    if (k1 - PANEL >= 0) {
      stage(k1 - PANEL, buf ^ 1);
      asm volatile("cp.async.wait_group 1;" ::: "memory");
    } else {
      asm volatile("cp.async.wait_group 0;" ::: "memory");
    }
    if (tid < PANEL) s_tau[buf][tid] = (tid < nk) ? (float)tt[k1 - tid] : 0.f;
    __syncthreads();
    const float* s_v = s_vbuf + (size_t)buf * PANEL * LD;
    // Gram values inside each group of four reflectors: 24 dot products, three per warp
    for (int d = w; d < (PANEL / GRP) * 6; d += NW) {
      const int g = d / 6, e = d % 6;
      const int qa = (e == 0) ? 1 : (e < 3) ? 2 : 3;
      const int qb = (e == 0) ? 0 : (e == 1) ? 0 : (e == 2) ? 1 : e - 3;
      const float* va = s_v + (GRP * g + qa) * LD;
      const float* vb = s_v + (GRP * g + qb) * LD;
      float acc = 0.f;
#pragma unroll
      for (int i = 0; i < LD / 32; ++i) acc = fmaf(va[lane + 32 * i], vb[lane + 32 * i], acc);
      acc = warp_sum(acc);
      if (lane == 0) s_g[buf][g][e] = acc;
    }
    __syncthreads();
    for (int g = 0; g < PANEL / GRP; ++g) {
      const int kk0 = GRP * g;
      if (kk0 >= nk) break;
      // rows <= kmin are zero in every reflector of the group: row pairs with 64 p + 63 <= kmin are skipped
      const int kmin = max(k1 - kk0 - (GRP - 1), 0);
      const int p0 = (kmin + 1) >> 6;
      const float2* vq[GRP];
#pragma unroll
      for (int q = 0; q < GRP; ++q) vq[q] = reinterpret_cast<const float2*>(s_v + (kk0 + q) * LD) + lane;
      F2 d2[GRP][C];
#pragma unroll
      for (int q = 0; q < GRP; ++q)
#pragma unroll
        for (int c = 0; c < C; ++c) d2[q][c] = f2_pack(0.f, 0.f);
#pragma unroll
      for (int p = 0; p < NP; ++p) {
        if (p < p0) continue;
#pragma unroll
        for (int q = 0; q < GRP; ++q) {
          const float2 a = vq[q][32 * p];
          const F2 a2 = f2_pack(a.x, a.y);
#pragma unroll
          for (int c = 0; c < C; ++c) d2[q][c] = f2_fma(a2, u[p][c], d2[q][c]);
        }
      }
      // 16 partial dot products per lane -> totals in every lane: transposed reduction (lane l ends with value
      // (l >> 1) & 15 summed over the warp), then an all-gather
      float t[16];
#pragma unroll
      for (int q = 0; q < GRP; ++q)
#pragma unroll
        for (int c = 0; c < C; ++c) {
          float lo, hi;
          f2_unpack(d2[q][c], lo, hi);
          t[q * C + c] = lo + hi;
        }
      float u8[8], u4[4], u2[2], u1;
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
      {
        const float mine = b1 ? u2[1] : u2[0], other = b1 ? u2[0] : u2[1];
        u1 = mine + __shfl_xor_sync(full, other, 2);
      }
      u1 += __shfl_xor_sync(full, u1, 1);
      float D[GRP][C];
#pragma unroll
      for (int q = 0; q < GRP; ++q)
#pragma unroll
        for (int c = 0; c < C; ++c) D[q][c] = __shfl_sync(full, u1, 2 * (q * C + c));
      const float g10 = s_g[buf][g][0], g20 = s_g[buf][g][1], g21 = s_g[buf][g][2], g30 = s_g[buf][g][3], g31 = s_g[buf][g][4],
                  g32 = s_g[buf][g][5];
      const float tau0 = s_tau[buf][kk0], tau1 = s_tau[buf][kk0 + 1], tau2 = s_tau[buf][kk0 + 2], tau3 = s_tau[buf][kk0 + 3];
      F2 ny[GRP][C];
#pragma unroll
      for (int c = 0; c < C; ++c) {
        const float y0 = tau0 * D[0][c];
        const float y1 = tau1 * (D[1][c] - y0 * g10);
        const float y2 = tau2 * (D[2][c] - y0 * g20 - y1 * g21);
        const float y3 = tau3 * (D[3][c] - y0 * g30 - y1 * g31 - y2 * g32);
        ny[0][c] = f2_pack(-y0, -y0);
        ny[1][c] = f2_pack(-y1, -y1);
        ny[2][c] = f2_pack(-y2, -y2);
        ny[3][c] = f2_pack(-y3, -y3);
      }
#pragma unroll
      for (int p = 0; p < NP; ++p) {
        if (p < p0) continue;
#pragma unroll
        for (int q = 0; q < GRP; ++q) {
          const float2 a = vq[q][32 * p];
          const F2 a2 = f2_pack(a.x, a.y);
#pragma unroll
          for (int c = 0; c < C; ++c) u[p][c] = f2_fma(ny[q][c], a2, u[p][c]);
        }
      }
    }
  }
  // column j of G = sqrt(lambda_j) u_j / |z_j| (see backtf4_kernel for the thresholds)
  const double lmax = eb.lam[(int64_t)lp * LD + N - 1];
#pragma unroll
  for (int c = 0; c < C; ++c) {
    const int j = j0 + c;
    if (j >= N) continue;
    const double lam = eb.lam[(int64_t)lp * LD + j], nz = eb.znorm[(int64_t)lp * LD + j];
    const bool in_range = j >= N - rk;
    const float sc = (in_range && lam > 1.0e-10 * lmax && nz > 0.0) ? (float)sqrt(lam / nz) : 0.f;
    float* __restrict__ G = b.G + (int64_t)lp * N * LD + (int64_t)j * LD;
#pragma unroll
    for (int p = 0; p < NP; ++p) {
      float lo, hi;
      f2_unpack(u[p][c], lo, hi);
      const int i0 = 64 * p + 2 * lane;
      *reinterpret_cast<float2*>(G + i0) = make_float2(i0 < N ? sc * lo : 0.f, i0 + 1 < N ? sc * hi : 0.f);
    }
  }
}

}  // namespace bt6

int siib_launch_backtf6(const SiibBuffers& b, const SiibEigBuffers& eb, int n, int rank_lo, cudaStream_t s) {
  constexpr int smem = 2 * bt6::PANEL * bt6::LD * (int)sizeof(float);
  static const bool attr = [] {
    cudaFuncSetAttribute(bt6::backtf6_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    return true;
  }();
  (void)attr;
  bt6::backtf6_kernel<<<dim3((bt6::N + bt6::VEC - 1) / bt6::VEC, n), bt6::NT, smem, s>>>(b, eb, rank_lo);
  return 1;
}

int siib_launch_backtf4(const SiibBuffers& b, const SiibEigBuffers& eb, int n, int rank_lo, cudaStream_t s) {
  constexpr int smem = 2 * bt::PANEL * bt::LD * (int)sizeof(float);
  static const bool attr = [] {
    cudaFuncSetAttribute(bt::backtf4_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    return true;
  }();
  (void)attr;
  bt::backtf4_kernel<<<dim3((bt::N + bt::VEC - 1) / bt::VEC, n), bt::NT, smem, s>>>(b, eb, rank_lo);
  return 1;
}

// tile lists of tridiag32_kernel, one per first row group, in quads of row groups and column-block major inside a quad
// (neighbouring warps then share row groups of v and columns of the partial vectors); per device
static void tile_tables_ready(cudaStream_t s) {
  static bool ready_dev[64] = {};
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev >= 0 && dev < 64 && ready_dev[dev]) return;
  static unsigned char tiles[klt::JMAX + 1][klt::kMaxTiles];
  static int ntiles[klt::JMAX + 1];
  for (int jm = 0; jm <= klt::JMAX; ++jm) {
    int n = 0;
    const int cmin = jm / 4;
    for (int Q = jm / 4; Q <= klt::JMAX / 4; ++Q)
      for (int C = cmin; C <= Q; ++C)
        for (int J = std::max(4 * Q, jm); J <= std::min(4 * Q + 3, klt::JMAX); ++J) tiles[jm][n++] = (unsigned char)(J << 3 | C);
    ntiles[jm] = n;
  }
  cudaMemcpyToSymbolAsync(klt::c_tiles, tiles, sizeof(tiles), 0, cudaMemcpyHostToDevice, s);
  cudaMemcpyToSymbolAsync(klt::c_ntiles, ntiles, sizeof(ntiles), 0, cudaMemcpyHostToDevice, s);
  cudaStreamSynchronize(s);
  if (dev >= 0 && dev < 64) ready_dev[dev] = true;
}

int siib_launch_tridiag32(const SiibBuffers& b, const SiibEigBuffers& eb, int n, int rank_lo, cudaStream_t s) {
  static const bool attr = [] {
    cudaFuncSetAttribute(klt::tridiag32_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(klt::Smem));
    return true;
  }();
  (void)attr;
  tile_tables_ready(s);
  klt::tridiag32_kernel<<<n, klt::NT, sizeof(klt::Smem), s>>>(b, eb, rank_lo, n);
  return 1;
}

}  // namespace nele
