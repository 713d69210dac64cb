// SIIB with the k-nearest-neighbour mutual-information estimator: pysiib.SIIB(x, y, fs,
// gauss=False), pysiib's default (NELE-GAN itself always passes gauss=True, intel.py:77,100).
// Algorithm: oracle/pysiib_np.py (mi_ksg: Kraskov-Stoegbauer-Grassberger estimator 1, max-norm,
// k = max(2, ceil(0.01 n)), capped by the production-noise bound -1/2 log2(1 - 0.75^2)).
//
// Shares everything up to the KLT basis with the Gaussian path (siib.cu: spectra, masking, lag
// products, Cholesky, Jacobi); then, instead of the quadratic forms:
//
//   siib_project_kernel  per (time tile, component tile, pair) CTA: the KLT-domain series
//                        Xk = U^T Xs, Yk = U^T Ys [r][Nf] as an FP32 register-tiled product.
//                        The stacked matrix Xs (420 x Nf) is never formed: row (k, j) of it at
//                        frame t is logspec[t + k][j], read from a staged [28][tile + 14] slab.
//   siib_ksg_kernel      per (pair, component) CTA: points sorted by x (bitonic, shared memory);
//                        one thread per point walks outwards over its x-neighbours, keeps the k
//                        smallest max-norm distances, and stops as soon as the x-distance alone
//                        exceeds the k-th; marginal counts by binary search in the sorted x and
//                        y; digamma terms summed in FP64
//   siib_knn_score_kernel  per pair: sum of the per-component estimates -> bits/s
#include "kernels.h"

namespace nele {

constexpr int kKDim = 420, kKLd = 448, kKLanes = 32, kKBands = 28, kKStack = 15;
constexpr int kKMaxN = 16384;  // frames per component the shared-memory sort holds
constexpr int kKMaxK = 164;    // ceil(0.01 * kKMaxN)

// --------------------------------------------------------------- projection
constexpr int kPjC = 64, kPjT = 128, kPjThreads = 256;
constexpr int kPjRows = kPjT + kKStack - 1;     // 142 frames per slab
constexpr int kPjLd = kPjRows + 2;              // row stride of the transposed slab

__global__ void __launch_bounds__(kPjThreads) siib_project_kernel(SiibGeom g, SiibBuffers b, SiibKnnBuffers kb) {
  const int lp = blockIdx.z, pair = b.pair_lo + lp, tid = threadIdx.x;
  const int Nf = b.Fa[pair] - (kKStack - 1);
  const int r = b.rank[pair];
  const int t0 = blockIdx.x * kPjT, c0 = blockIdx.y * kPjC;
  if (Nf < 2 || t0 >= Nf || c0 >= r) return;
  extern __shared__ __align__(16) float s_pj[];
  float* sU = s_pj;                              // [420][64]  unit eigenvectors, component-minor
  float* sX = sU + kKDim * kPjC;                 // [28][kPjLd]
  float* sY = sX + kKBands * kPjLd;
  __shared__ float s_inv[kPjC];
  const float* __restrict__ G = b.G + (int64_t)lp * kKDim * kKLd;
  // 1 / |column| of the 64 components of this tile
  {
    const int lane = tid & 31, wib = tid >> 5;
    for (int c = wib; c < kPjC; c += kPjThreads / 32) {
      float ss = 0.f;
      if (c0 + c < r) {
        const float* col = G + (int64_t)(c0 + c) * kKLd;
        for (int i = lane; i < kKDim; i += 32) ss = fmaf(col[i], col[i], ss);
      }
      ss = warp_sum(ss);
      if (lane == 0) s_inv[c] = ss > 0.f ? rsqrtf(ss) : 0.f;
    }
  }
  __syncthreads();
  for (int idx = tid; idx < kKDim * kPjC; idx += kPjThreads) {
    const int c = idx / kKDim, i = idx % kKDim;  // coalesced along the column
    sU[i * kPjC + c] = (c0 + c < r) ? G[(int64_t)(c0 + c) * kKLd + i] * s_inv[c] : 0.f;
  }
  const float* __restrict__ X = b.logspec + (g.offF[pair]) * kKLanes;
  const float* __restrict__ Y = b.logspec + (b.totF + g.offF[pair]) * kKLanes;
  const int Fa = b.Fa[pair];
  for (int idx = tid; idx < kPjRows * kKLanes; idx += kPjThreads) {
    const int row = idx / kKLanes, j = idx % kKLanes;
    if (j < kKBands) {
      const bool ok = t0 + row < Fa;
      sX[j * kPjLd + row] = ok ? X[(int64_t)(t0 + row) * kKLanes + j] : 0.f;
      sY[j * kPjLd + row] = ok ? Y[(int64_t)(t0 + row) * kKLanes + j] : 0.f;
    }
  }
  __syncthreads();
  // thread (tc, tt): components 4 tc .. 4 tc + 3, frames tt + 16 u, u = 0..7
  const int tt = tid & 15, tc = tid >> 4;
  float ax[4][8], ay[4][8];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int u = 0; u < 8; ++u) ax[i][u] = ay[i][u] = 0.f;
  for (int k = 0; k < kKStack; ++k) {
#pragma unroll 4
    for (int j = 0; j < kKBands; ++j) {
      const float4 uu = *reinterpret_cast<const float4*>(sU + (k * kKBands + j) * kPjC + 4 * tc);
      const float* xr = sX + j * kPjLd + k + tt;
      const float* yr = sY + j * kPjLd + k + tt;
      float xv[8], yv[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        xv[u] = xr[16 * u];
        yv[u] = yr[16 * u];
      }
      const float uc[4] = {uu.x, uu.y, uu.z, uu.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          ax[i][u] = fmaf(uc[i], xv[u], ax[i][u]);
          ay[i][u] = fmaf(uc[i], yv[u], ay[i][u]);
        }
    }
  }
  float* __restrict__ ox = kb.xk + ((int64_t)lp * 2 * kKDim) * kb.ld;
  float* __restrict__ oy = ox + (int64_t)kKDim * kb.ld;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int c = c0 + 4 * tc + i;
    if (c >= r) continue;
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const int t = t0 + tt + 16 * u;
      if (t < Nf) {
        ox[(int64_t)c * kb.ld + t] = ax[i][u];
        oy[(int64_t)c * kb.ld + t] = ay[i][u];
      }
    }
  }
}

// ------------------------------------------------------------------- KSG
constexpr int kKsgThreads = 256;

// bitonic sort of n_pad (power of two) keys in shared memory, optional payload
template <bool PAYLOAD>
__device__ __forceinline__ void bitonic_sort(float* key, float* val, int n_pad) {
  for (int size = 2; size <= n_pad; size <<= 1) {
    for (int stride = size >> 1; stride > 0; stride >>= 1) {
      __syncthreads();
      for (int i = threadIdx.x; i < (n_pad >> 1); i += blockDim.x) {
        const int lo = 2 * i - (i & (stride - 1));  // index with the `stride` bit clear
        const int hi = lo + stride;
        const bool up = (lo & size) == 0;
        const float a = key[lo], c = key[hi];
        if ((a > c) == up) {
          key[lo] = c;
          key[hi] = a;
          if (PAYLOAD) {
            const float t = val[lo];
            val[lo] = val[hi];
            val[hi] = t;
          }
        }
      }
    }
  }
  __syncthreads();
}

__device__ __forceinline__ double digamma_int(int m, const double* __restrict__ tab, int ntab) {
  if (m < ntab) return tab[m];
  const double x = (double)m, x2 = 1.0 / (x * x);  // asymptotic series, |error| < 1e-15 for m >= 64
  return log(x) - 0.5 / x - x2 * (1.0 / 12.0 - x2 * (1.0 / 120.0 - x2 * (1.0 / 252.0)));
}

__global__ void __launch_bounds__(kKsgThreads) siib_ksg_kernel(SiibGeom g, SiibBuffers b, SiibKnnBuffers kb) {
  const int comp = blockIdx.x, lp = blockIdx.y, pair = b.pair_lo + lp, tid = threadIdx.x;
  const int n = b.Fa[pair] - (kKStack - 1);
  const int r = b.rank[pair];
  double* __restrict__ out = kb.info + (int64_t)pair * kKDim;
  if (comp >= r || n < 4 || n > kKMaxN) {
    if (tid == 0) out[comp] = 0.0;
    return;
  }
  extern __shared__ __align__(16) float s_k[];
  int n_pad = 1;
  while (n_pad < n) n_pad <<= 1;
  float* sx = s_k;              // x sorted ascending
  float* sy = sx + n_pad;       // y in the order of sx
  float* ys = sy + n_pad;       // y sorted ascending
  __shared__ double red[32];
  const float* gx = kb.xk + ((int64_t)lp * 2 * kKDim + comp) * kb.ld;
  const float* gy = gx + (int64_t)kKDim * kb.ld;
  // The neighbour walk runs along the sorted coordinate and stops when that coordinate's distance
  // alone exceeds the k-th max-norm distance: it prunes well only along the coordinate with the
  // larger spread (a low-eigenvalue component has a tiny x range under a noisy y).  The max-norm
  // is symmetric, so the roles of x and y are simply swapped when y spreads more.
  {
    double lo_x = 1.0e300, hi_x = -1.0e300, lo_y = 1.0e300, hi_y = -1.0e300;
    for (int i = tid; i < n; i += kKsgThreads) {
      const double xv = (double)gx[i], yv = (double)gy[i];
      lo_x = fmin(lo_x, xv);
      hi_x = fmax(hi_x, xv);
      lo_y = fmin(lo_y, yv);
      hi_y = fmax(hi_y, yv);
    }
    const double rx = block_max(hi_x, red) + block_max(-lo_x, red);
    const double ry = block_max(hi_y, red) + block_max(-lo_y, red);
    if (!(rx > 0.0)) {  // a component of the numerical null space (zero column of G): no information
      if (tid == 0) out[comp] = 0.0;
      return;
    }
    if (ry > rx) {
      const float* t = gx;
      gx = gy;
      gy = t;
    }
  }
  for (int i = tid; i < n_pad; i += kKsgThreads) {
    const bool ok = i < n;
    const float xv = ok ? gx[i] : 3.0e38f, yv = ok ? gy[i] : 3.0e38f;
    sx[i] = xv;
    sy[i] = yv;
    ys[i] = yv;
  }
  bitonic_sort<true>(sx, sy, n_pad);
  bitonic_sort<false>(ys, nullptr, n_pad);
  const int k = max(2, (n + 99) / 100);  // max(2, ceil(0.01 n))
  float best[kKMaxK];                    // k smallest distances so far, ascending (local memory)
  double acc = 0.0;
  for (int p = tid; p < n; p += kKsgThreads) {
    const float x0 = sx[p], y0 = sy[p];
    int cnt = 0;
    float kth = 3.0e38f;
    int l = p - 1, rr = p + 1;
    float dl = (l >= 0) ? x0 - sx[l] : 3.0e38f, dr = (rr < n) ? sx[rr] - x0 : 3.0e38f;
    while (true) {
      const bool left = dl <= dr;
      const float dx = left ? dl : dr;
      if (dx >= kth) break;  // also ends the scan when both sides are exhausted (dx = 3e38)
      const float yo = left ? sy[l] : sy[rr];
      const float d = fmaxf(dx, fabsf(yo - y0));
      if (left) {
        --l;
        dl = (l >= 0) ? x0 - sx[l] : 3.0e38f;
      } else {
        ++rr;
        dr = (rr < n) ? sx[rr] - x0 : 3.0e38f;
      }
      if (cnt < k) {
        int q = cnt++;
        while (q > 0 && best[q - 1] > d) {
          best[q] = best[q - 1];
          --q;
        }
        best[q] = d;
        if (cnt == k) kth = best[k - 1];
      } else if (d < kth) {
        int q = k - 1;
        while (q > 0 && best[q - 1] > d) {
          best[q] = best[q - 1];
          --q;
        }
        best[q] = d;
        kth = best[k - 1];
      }
    }
    const float eps = (cnt == k) ? kth : best[max(cnt - 1, 0)];
    // marginal counts: points strictly inside (x0 - eps, x0 + eps), self excluded
    int lo = 0, hi = p;  // first index with x0 - sx[i] < eps
    while (lo < hi) {
      const int mid = (lo + hi) >> 1;
      if (x0 - sx[mid] < eps) hi = mid;
      else lo = mid + 1;
    }
    const int xl = lo;
    lo = p;
    hi = n;              // first index > p with sx[i] - x0 >= eps
    while (lo < hi) {
      const int mid = (lo + hi) >> 1;
      if (sx[mid] - x0 < eps) lo = mid + 1;
      else hi = mid;
    }
    const int nx = lo - xl - 1;
    lo = 0;
    hi = n;              // first index with y0 - ys[i] < eps
    while (lo < hi) {
      const int mid = (lo + hi) >> 1;
      if (y0 - ys[mid] < eps) hi = mid;
      else lo = mid + 1;
    }
    const int yl = lo;
    hi = n;              // first index with ys[i] - y0 >= eps
    while (lo < hi) {
      const int mid = (lo + hi) >> 1;
      if (ys[mid] - y0 < eps) lo = mid + 1;
      else hi = mid;
    }
    const int ny = lo - yl - 1;
    acc += digamma_int(max(nx, 0) + 1, kb.digamma, kb.ndigamma) + digamma_int(max(ny, 0) + 1, kb.digamma, kb.ndigamma);
  }
  acc = block_sum(acc, red);
  if (tid == 0) {
    const double nats = digamma_int(k, kb.digamma, kb.ndigamma) + digamma_int(n, kb.digamma, kb.ndigamma) - acc / (double)n;
    const double cap = -0.5 * log2(1.0 - 0.75 * 0.75);
    out[comp] = fmin(nats / 0.6931471805599453, cap);
  }
}

__global__ void siib_knn_score_kernel(SiibGeom g, SiibBuffers b, SiibKnnBuffers kb, int n_pairs) {
  const int pair = b.pair_lo + blockIdx.x, tid = threadIdx.x;
  __shared__ double red[32];
  const int Fa = b.Fa[pair];
  const int Nf = Fa - (kKStack - 1);
  const int M = b.M[pair];
  if (M <= 0 || (double)Fa / 80.0 < 20.0 || Nf < 2) {  // pysiib: "at least 20 seconds of speech"
    if (tid == 0) {
      b.score[pair] = nan("");
      b.status[pair] = 2;
    }
    return;
  }
  if (Nf > kKMaxN) {
    if (tid == 0) {
      b.score[pair] = nan("");
      b.status[pair] = 4;
    }
    return;
  }
  const int r = b.rank[pair];
  double s = 0.0;
  for (int c = tid; c < r; c += blockDim.x) s += kb.info[(int64_t)pair * kKDim + c];
  s = block_sum(s, red);
  if (tid == 0) {
    const double v = (16000.0 / 200.0) / (double)kKStack * s;
    b.score[pair] = v > 0.0 ? v : 0.0;
    b.status[pair] = 0;
  }
}

void siib_knn_setup() {
  cudaFuncSetAttribute(siib_project_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                       (int)((kKDim * kPjC + 2 * kKBands * kPjLd) * sizeof(float)));
  cudaFuncSetAttribute(siib_ksg_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 3 * kKMaxN * (int)sizeof(float));
}

int siib_run_knn(const SiibGeom& g, const SiibBuffers& b, const SiibKnnBuffers& kb, int n, int64_t max_F, KernelTimer* kt,
                 cudaStream_t s) {
  int launches = 0;
  const int64_t max_nf = max_F - (kKStack - 1);
  if (max_nf >= 2) {
    kt_begin(kt, "siib_project", s);
    siib_project_kernel<<<dim3((unsigned)((max_nf + kPjT - 1) / kPjT), (kKDim + kPjC - 1) / kPjC, n), kPjThreads,
                          (kKDim * kPjC + 2 * kKBands * kPjLd) * sizeof(float), s>>>(g, b, kb);
    kt_end(kt, s);
    ++launches;
    int n_pad = 1;
    while (n_pad < std::min<int64_t>(max_nf, kKMaxN)) n_pad <<= 1;
    kt_begin(kt, "siib_ksg", s);
    siib_ksg_kernel<<<dim3(kKDim, n), kKsgThreads, 3 * (size_t)n_pad * sizeof(float), s>>>(g, b, kb);
    kt_end(kt, s);
    ++launches;
  }
  kt_begin(kt, "siib_knn_score", s);
  siib_knn_score_kernel<<<n, 128, 0, s>>>(g, b, kb, n);
  kt_end(kt, s);
  ++launches;
  return launches;
}

}  // namespace nele
