// Microbenchmark: FP32 FMA issue rate, scalar FFMA vs packed FFMA2 (fma.rn.f32x2), sm_100a.
// Measured on a B200 (1965 MHz): scalar 72.3 TFLOP/s, packed 74.0 TFLOP/s -- the same FMA-pipe throughput
// (peak 148 x 128 x 2 x 1.965 GHz = 74.4), so FFMA2 halves the issue slots of FMA work, not its pipe time.
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ unsigned long long pk(float a, float b) { unsigned long long r; asm("mov.b64 %0, {%1,%2};" : "=l"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ void upk(unsigned long long v, float& a, float& b) { asm("mov.b64 {%0,%1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); }
__device__ __forceinline__ unsigned long long fma2(unsigned long long a, unsigned long long b, unsigned long long c) { unsigned long long d; asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }
constexpr int ITERS = 4096, CH = 8;
__global__ void k_scalar(float* out, float a) {
  float v[2 * CH];
  for (int i = 0; i < 2 * CH; ++i) v[i] = threadIdx.x * 1e-3f + i;
  for (int it = 0; it < ITERS; ++it)
#pragma unroll
    for (int i = 0; i < 2 * CH; ++i) v[i] = fmaf(v[i], a, 0.5f);
  float s = 0; for (int i = 0; i < 2 * CH; ++i) s += v[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void k_packed(float* out, float a) {
  unsigned long long v[CH], aa = pk(a, a), cc = pk(0.5f, 0.5f);
  for (int i = 0; i < CH; ++i) v[i] = pk(threadIdx.x * 1e-3f + i, threadIdx.x * 2e-3f + i);
  for (int it = 0; it < ITERS; ++it)
#pragma unroll
    for (int i = 0; i < CH; ++i) v[i] = fma2(v[i], aa, cc);
  float s = 0; for (int i = 0; i < CH; ++i) { float p, q; upk(v[i], p, q); s += p + q; }
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
int main() {
  float* d; cudaMalloc(&d, 148 * 8 * 1024 * 4);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (int rep = 0; rep < 2; ++rep) {
    float ms;
    cudaEventRecord(e0); k_scalar<<<148 * 8, 1024>>>(d, 0.999f); cudaEventRecord(e1); cudaEventSynchronize(e1); cudaEventElapsedTime(&ms, e0, e1);
    double fl = 148.0 * 8 * 1024 * ITERS * 2 * CH * 2;
    printf("scalar FFMA : %.3f ms  %.1f TFLOP/s\n", ms, fl / ms * 1e-9);
    cudaEventRecord(e0); k_packed<<<148 * 8, 1024>>>(d, 0.999f); cudaEventRecord(e1); cudaEventSynchronize(e1); cudaEventElapsedTime(&ms, e0, e1);
    printf("packed FFMA2: %.3f ms  %.1f TFLOP/s\n", ms, fl / ms * 1e-9);
  }
  return 0;
}
