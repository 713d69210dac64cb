// Shared helpers for the nele_score CUDA kernels (sm_100a).
#pragma once
#if defined(__CUDACC__)
#include <cuda_runtime.h>
#endif
#include <stdint.h>

#if defined(__CUDACC__)
#define NELE_HD __host__ __device__ __forceinline__
#else
#define NELE_HD inline
#endif

namespace nele {

constexpr int kBands = 32;       // auditory filterbank channels (pyhaspi2.py:1157)
constexpr int kFs24 = 24000;     // ear-model rate (pyhaspi2.py:811)
constexpr int kDecim = 9;        // int(24000 // 2560)  (pyhaspi2.py:410)
constexpr int kEnvTaps = 52;     // ebm_EnvFilt Hann FIR length (pyhaspi2.py:390-394)
constexpr int kEnvHalf = 26;
constexpr int kNumMod = 10;      // modulation bands (pyhaspi2.py:277)
constexpr int kNumCep = 5;       // cepstral coefficients 2..6 enter the score (pyhaspi2.py:272)

#ifdef __CUDACC__
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double warp_max(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ int warp_sum(int v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// Block-wide sum of doubles; `red` is shared scratch of >= 32 doubles.  Every
// thread gets the result.
__device__ __forceinline__ double block_sum(double v, double* red) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  v = warp_sum(v);
  __syncthreads();
  if (lane == 0) red[w] = v;
  __syncthreads();
  double t = (lane < nw) ? red[lane] : 0.0;
  t = warp_sum(t);
  return t;
}
__device__ __forceinline__ double block_max(double v, double* red) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  v = warp_max(v);
  __syncthreads();
  if (lane == 0) red[w] = v;
  __syncthreads();
  double t = (lane < nw) ? red[lane] : -1.0e300;
  t = warp_max(t);
  return t;
}
#endif

}  // namespace nele
