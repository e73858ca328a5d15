// Microbenchmark: issue rate of DMMA.8x8x4 (mma.sync m8n8k4 f64) vs DFMA on one GPU.
// nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o dmma_rate dmma_rate.cu
#include <cstdio>
#include <cuda_runtime.h>

__global__ void dmma_loop(double* out, int iters) {
  double a = threadIdx.x * 1e-3 + 1.0, b = 1.0 + blockIdx.x * 1e-6;
  double c[8][2];
#pragma unroll
  for (int i = 0; i < 8; ++i) c[i][0] = c[i][1] = 0.0;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i)
      asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0, %1}, {%2}, {%3}, {%0, %1};"
                   : "+d"(c[i][0]), "+d"(c[i][1])
                   : "d"(a), "d"(b));
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += c[i][0] + c[i][1];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void dfma_loop(double* out, int iters) {
  double a = threadIdx.x * 1e-9 + 1.0, b = 1e-9;
  double c[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) c[i] = i;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 16; ++i) c[i] = fma(c[i], a, b);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < 16; ++i) s += c[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

int main() {
  double* out;
  cudaMalloc(&out, 148 * 8 * 256 * sizeof(double));
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  const int iters = 20000;
  for (int rep = 0; rep < 2; ++rep) {
    cudaEventRecord(e0);
    dmma_loop<<<148 * 4, 256>>>(out, iters);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    double n = 148.0 * 4 * 8 * iters * 8;  // warp-level DMMA instructions
    printf("DMMA.8x8x4: %.3f ms, %.2f TFLOP/s (512 flop each), %.1f SM-cycles per warp instruction at 1.965 GHz\n", ms,
           n * 512 / ms / 1e9, ms * 1e-3 * 1.965e9 * 148 / n);
    cudaEventRecord(e0);
    dfma_loop<<<148 * 4, 256>>>(out, iters);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    cudaEventElapsedTime(&ms, e0, e1);
    n = 148.0 * 4 * 8 * iters * 16;
    printf("DFMA      : %.3f ms, %.2f TFLOP/s (64 flop each), %.1f SM-cycles per warp instruction\n", ms,
           n * 64 / ms / 1e9, ms * 1e-3 * 1.965e9 * 148 / n);
  }
  return 0;
}
