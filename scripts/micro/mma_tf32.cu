// Throughput of the legacy tensor path (mma.sync.m16n8k8 tf32) on sm_100a, to decide whether a 3xTF32 split
// (a_hi b_hi + a_hi b_lo + a_lo b_hi, ~FP32 accuracy) beats FP32 FFMA for the KLT's dense contractions.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mma_tf32 mma_tf32.cu ; run: ./mma_tf32
#include <cstdio>
#include <cuda_runtime.h>

__global__ void __launch_bounds__(256) mma_loop(float* out, int iters) {
  unsigned a[4] = {0x3f800000u + threadIdx.x, 0x3f900000u, 0x3fa00000u, 0x3fb00000u};
  unsigned b[2] = {0x3f800000u, 0x3f880000u + threadIdx.x};
  float c[8][4];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) c[i][j] = 0.f;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i)
      asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                   : "+f"(c[i][0]), "+f"(c[i][1]), "+f"(c[i][2]), "+f"(c[i][3])
                   : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
  }
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += c[i][0] + c[i][1] + c[i][2] + c[i][3];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void __launch_bounds__(256) ffma_loop(float* out, int iters) {
  float c[32];
  const float a = 1.0001f + threadIdx.x * 1e-6f, b = 0.5f;
#pragma unroll
  for (int i = 0; i < 32; ++i) c[i] = (float)i;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 32; ++i) c[i] = fmaf(c[i], a, b);
  }
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 32; ++i) s += c[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

int main() {
  float* out;
  cudaMalloc(&out, 148 * 8 * 256 * sizeof(float));
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  const int iters = 20000;
  for (int ctas_per_sm = 1; ctas_per_sm <= 8; ctas_per_sm *= 2) {
    const int grid = 148 * ctas_per_sm;
    mma_loop<<<grid, 256>>>(out, 100);
    cudaDeviceSynchronize();
    cudaEventRecord(e0);
    mma_loop<<<grid, 256>>>(out, iters);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    const double flop = (double)grid * 8 /*warps*/ * iters * 8 /*mma*/ * 2.0 * 16 * 8 * 8;
    printf("mma.sync m16n8k8 tf32: %d CTAs/SM x 8 warps: %.3f ms -> %.1f TFLOP/s\n", ctas_per_sm, ms, flop / ms * 1e-9);
  }
  for (int ctas_per_sm = 2; ctas_per_sm <= 8; ctas_per_sm *= 2) {
    const int grid = 148 * ctas_per_sm;
    ffma_loop<<<grid, 256>>>(out, 100);
    cudaDeviceSynchronize();
    cudaEventRecord(e0);
    ffma_loop<<<grid, 256>>>(out, iters);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    const double flop = (double)grid * 256 * iters * 32 * 2.0;
    printf("FFMA: %d CTAs/SM x 256 threads: %.3f ms -> %.1f TFLOP/s\n", ctas_per_sm, ms, flop / ms * 1e-9);
  }
  return 0;
}
