// Microbenchmark: the affine inner loop of demux_score_kernel in isolation (no ring, no epilogue) against a
// pure DFMA chain, to separate instruction-level limits from pipeline/epilogue effects.
// build: nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o affine_loop affine_loop.cu
#include <cstdio>
#include <cuda_runtime.h>

template <int AC, int WARPS, bool PAIR>
__global__ void __launch_bounds__(WARPS * 32, 1) loop_kernel(const double* __restrict__ fin, double* out, int nsnp, int reps) {
  __shared__ double sf[64 * 64];  // [snp 64][sample 64]: j in 0..31, k in 32..63
  for (int i = threadIdx.x; i < 64 * 64; i += blockDim.x) sf[i] = fin[i];
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int kidx = (warp & 3) * 8 + (lane & 7), jg = ((warp >> 2) * 4 + (lane >> 3)) & 15;
  double acc[2][AC];
  double al[AC];
#pragma unroll
  for (int a = 0; a < AC; ++a) {
    acc[0][a] = acc[1][a] = 1.0;
    al[a] = 0.05 * (a + 1);
  }
  for (int r = 0; r < reps; ++r) {
    if (!PAIR) {
#pragma unroll 2
      for (int s = 0; s < nsnp; ++s) {
        const double2 fj = *reinterpret_cast<const double2*>(&sf[(s & 63) * 64 + jg * 2]);
        const double fk = sf[(s & 63) * 64 + 32 + kidx];
        const double d0 = fk - fj.x, d1 = fk - fj.y;
#pragma unroll
        for (int a = 0; a < AC; ++a) {
          acc[0][a] *= fma(al[a], d0, fj.x);
          acc[1][a] *= fma(al[a], d1, fj.y);
        }
      }
    } else {
      for (int s = 0; s < nsnp; s += 2) {
        const double2 fj1 = *reinterpret_cast<const double2*>(&sf[(s & 63) * 64 + jg * 2]);
        const double2 fj2 = *reinterpret_cast<const double2*>(&sf[((s + 1) & 63) * 64 + jg * 2]);
        const double fk1 = sf[(s & 63) * 64 + 32 + kidx], fk2 = sf[((s + 1) & 63) * 64 + 32 + kidx];
        const double d10 = fk1 - fj1.x, d11 = fk1 - fj1.y, d20 = fk2 - fj2.x, d21 = fk2 - fj2.y;
        const double c00 = fj1.x * fj2.x, c01 = fj1.y * fj2.y;
        const double c20 = d10 * d20, c21 = d11 * d21;
        const double c10 = fma(fj1.x, d20, fj2.x * d10), c11 = fma(fj1.y, d21, fj2.y * d11);
#pragma unroll
        for (int a = 0; a < AC; ++a) {
          acc[0][a] *= fma(al[a], fma(al[a], c20, c10), c00);
          acc[1][a] *= fma(al[a], fma(al[a], c21, c11), c01);
        }
      }
    }
#pragma unroll
    for (int a = 0; a < AC; ++a) {  // keep the products in range (stand-in for the renormalisation)
      acc[0][a] = acc[0][a] * 0x1p+60 + 1.0;
      acc[1][a] = acc[1][a] * 0x1p+60 + 1.0;
    }
  }
  double s = 0;
#pragma unroll
  for (int a = 0; a < AC; ++a) s += acc[0][a] + acc[1][a];
  if (s == 123.456) out[0] = s;
}

__global__ void __launch_bounds__(256) peak_kernel(double* sink, int iters) {
  double a0 = 1.0 + threadIdx.x * 1e-9, a1 = a0 + 1e-3, a2 = a0 + 2e-3, a3 = a0 + 3e-3, a4 = a0 + 4e-3, a5 = a0 + 5e-3,
         a6 = a0 + 6e-3, a7 = a0 + 7e-3;
  const double x = 0.999999, y = 1e-7;
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      a0 = fma(a0, x, y); a1 = fma(a1, x, y); a2 = fma(a2, x, y); a3 = fma(a3, x, y);
      a4 = fma(a4, x, y); a5 = fma(a5, x, y); a6 = fma(a6, x, y); a7 = fma(a7, x, y);
    }
  }
  const double s = ((a0 + a1) + (a2 + a3)) + ((a4 + a5) + (a6 + a7));
  if (s == 123.456) sink[0] = s;
}

template <typename F>
float time_ms(F f) {
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  f();
  cudaDeviceSynchronize();
  cudaEventRecord(e0);
  f();
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms;
  cudaEventElapsedTime(&ms, e0, e1);
  return ms;
}

int main() {
  double *fin, *out;
  cudaMalloc(&fin, 64 * 64 * 8);
  cudaMalloc(&out, 64);
  double h[64 * 64];
  for (int i = 0; i < 64 * 64; ++i) h[i] = 0.3 + 0.6 * ((i * 2654435761u) % 1000) / 1000.0;
  cudaMemcpy(fin, h, sizeof(h), cudaMemcpyHostToDevice);
  const int nsnp = 192, reps = 400;
  {
    float ms = time_ms([&] { peak_kernel<<<148 * 4, 256>>>(out, 20000); });
    printf("peak DFMA chain        : %.2f TFLOP/s\n", 148.0 * 4 * 256 * 20000.0 * 64 * 2 / (ms * 1e-3) / 1e12);
  }
  auto report = [&](const char* name, float ms, double instr_per_snp_thread, int threads) {
    const double inst = 148.0 * threads * (double)nsnp * reps * instr_per_snp_thread;
    printf("%-22s : %.3f ms, FP64 instr rate %.2f TFLOP/s-equivalent (x2 flops)\n", name, ms, inst * 2 / (ms * 1e-3) / 1e12);
  };
  report("affine 16 warps AC=10", time_ms([&] { loop_kernel<10, 16, false><<<148, 512>>>(fin, out, nsnp, reps); }), 42, 512);
  report("affine  8 warps AC=10", time_ms([&] { loop_kernel<10, 8, false><<<148, 256>>>(fin, out, nsnp, reps); }), 42, 256);
  report("affine 12 warps AC=10", time_ms([&] { loop_kernel<10, 12, false><<<148, 384>>>(fin, out, nsnp, reps); }), 42, 384);
  report("pair   16 warps AC=10", time_ms([&] { loop_kernel<10, 16, true><<<148, 512>>>(fin, out, nsnp, reps); }), 36, 512);
  printf("(pair: 72 FP64 instr per 2 SNPs per thread for the same 40 likelihood factors; affine: 84)\n");
  cudaError_t e = cudaDeviceSynchronize();
  printf("status: %s\n", cudaGetErrorString(e));
  return e != cudaSuccess;
}
