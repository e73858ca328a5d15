// Microbenchmark: does a non-FP64 instruction cost FP64-pipe time on B200?  One 512-thread CTA per SM (4 warps per
// sub-partition).  (a) warps split by role: NF warps per sub-partition run independent DFMA chains, the others run
// integer chains (or nothing); (b) every warp interleaves R integer instructions per 4 DFMAs.
// If the dispatch port is held for both cycles of a half-rate FP64 instruction, time = 2*N_fp64 + N_other;
// if not, time = max(2*N_fp64, N_fp64 + N_other).
// build: nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o fp64_coissue fp64_coissue.cu
#include <cstdio>
#include <cuda_runtime.h>

#define DFMA(x) asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(x) : "d"(a), "d"(b))
#ifdef IOP_ADD  // ALU pipe instead of the FMA pipe
#define IOP(x) asm volatile("add.s32 %0, %0, %1;" : "+r"(x) : "r"(m + n))
#else
#define IOP(x) asm volatile("mad.lo.s32 %0, %0, %1, %2;" : "+r"(x) : "r"(m), "r"(n))
#endif

template <int NF, bool INT_OTHERS>
__global__ void __launch_bounds__(512, 1) role_kernel(double* out, int iters) {
  const int warp = threadIdx.x >> 5;
  const bool fp = (warp >> 2) < NF;  // warps w, w+4, w+8, w+12 share a sub-partition
  double a = 1.0 + 1e-9 * threadIdx.x, b = 1e-9;
  int m = 3 + threadIdx.x, n = 7;
  double c[16];
  int q[8];
#pragma unroll
  for (int i = 0; i < 16; ++i) c[i] = i;
#pragma unroll
  for (int i = 0; i < 8; ++i) q[i] = i;
  if (fp) {
    for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int i = 0; i < 16; ++i) DFMA(c[i]);
    }
  } else if (INT_OTHERS) {
    for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int r = 0; r < 4; ++r)
#pragma unroll
        for (int i = 0; i < 8; ++i) IOP(q[i]);
    }
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < 16; ++i) s += c[i];
#pragma unroll
  for (int i = 0; i < 8; ++i) s += q[i];
  out[blockIdx.x * 512 + threadIdx.x] = s;
}

template <int R>  // R integer instructions per 4 DFMAs, same warp
__global__ void __launch_bounds__(512, 1) mix_kernel(double* out, int iters) {
  double a = 1.0 + 1e-9 * threadIdx.x, b = 1e-9;
  int m = 3 + threadIdx.x, n = 7;
  double c[16];
  int q[8];
#pragma unroll
  for (int i = 0; i < 16; ++i) c[i] = i;
#pragma unroll
  for (int i = 0; i < 8; ++i) q[i] = i;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int g = 0; g < 4; ++g) {
#pragma unroll
      for (int i = 0; i < 4; ++i) DFMA(c[g * 4 + i]);
#pragma unroll
      for (int i = 0; i < R; ++i) IOP(q[(g * R + i) & 7]);
    }
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < 16; ++i) s += c[i];
#pragma unroll
  for (int i = 0; i < 8; ++i) s += q[i];
  out[blockIdx.x * 512 + threadIdx.x] = s;
}

template <typename F>
static float time_ms(F launch) {
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  launch();
  cudaDeviceSynchronize();
  cudaEventRecord(e0);
  launch();
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms;
  cudaEventElapsedTime(&ms, e0, e1);
  return ms;
}

int main() {
  int sms = 0, khz = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
  double* out;
  cudaMalloc(&out, sizeof(double) * 512 * sms);
  const int iters = 200000;
  // cycles per DFMA warp-instruction and sub-partition = ms * clock / (DFMA warp-instructions per sub-partition)
  auto cyc = [&](float ms, double dfma_per_smsp) { return ms * 1e-3 * khz * 1e3 / dfma_per_smsp; };
  float t;
  t = time_ms([&] { role_kernel<4, false><<<sms, 512>>>(out, iters); });
  printf("roles: 4 DFMA warps / SMSP                 %.3f ms  %.3f cycles per DFMA\n", t, cyc(t, 4.0 * 16 * iters));
  t = time_ms([&] { role_kernel<2, false><<<sms, 512>>>(out, iters); });
  printf("roles: 2 DFMA warps, 2 idle                %.3f ms  %.3f cycles per DFMA\n", t, cyc(t, 2.0 * 16 * iters));
  t = time_ms([&] { role_kernel<2, true><<<sms, 512>>>(out, iters); });
  printf("roles: 2 DFMA warps, 2 integer warps (32 IMAD per 16 DFMA)  %.3f ms  %.3f cycles per DFMA\n", t, cyc(t, 2.0 * 16 * iters));
  t = time_ms([&] { role_kernel<1, false><<<sms, 512>>>(out, iters); });
  printf("roles: 1 DFMA warp, 3 idle                 %.3f ms  %.3f cycles per DFMA\n", t, cyc(t, 1.0 * 16 * iters));
  t = time_ms([&] { role_kernel<0, true><<<sms, 512>>>(out, iters); });
  printf("roles: 4 integer warps                     %.3f ms  %.3f cycles per IMAD\n", t, cyc(t, 4.0 * 32 * iters));
  t = time_ms([&] { mix_kernel<0><<<sms, 512>>>(out, iters); });
  printf("mix: 0 IMAD per 4 DFMA                     %.3f ms  %.3f cycles per DFMA\n", t, cyc(t, 4.0 * 16 * iters));
  t = time_ms([&] { mix_kernel<1><<<sms, 512>>>(out, iters); });
  printf("mix: 1 IMAD per 4 DFMA                     %.3f ms  %.3f cycles per DFMA\n", t, cyc(t, 4.0 * 16 * iters));
  t = time_ms([&] { mix_kernel<2><<<sms, 512>>>(out, iters); });
  printf("mix: 2 IMAD per 4 DFMA                     %.3f ms  %.3f cycles per DFMA\n", t, cyc(t, 4.0 * 16 * iters));
  t = time_ms([&] { mix_kernel<4><<<sms, 512>>>(out, iters); });
  printf("mix: 4 IMAD per 4 DFMA                     %.3f ms  %.3f cycles per DFMA\n", t, cyc(t, 4.0 * 16 * iters));
  t = time_ms([&] { mix_kernel<8><<<sms, 512>>>(out, iters); });
  printf("mix: 8 IMAD per 4 DFMA                     %.3f ms  %.3f cycles per DFMA\n", t, cyc(t, 4.0 * 16 * iters));
  printf("clock %d kHz (attribute; the measured cycles assume it)\n", khz);
  return 0;
}
