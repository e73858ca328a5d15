// demux_kernels.cu -- sm_100a kernels of the demuxlet droplet-likelihood engine.
//
//   gp_avg / gp_mix   genotype-error mixing of the float GP matrix   (sc_drop_seq.cpp:287-315)
//   K1 tile           per-(cell,SNP) read-product tile, 9 x nAlpha   (cmd_cram_demuxlet.cpp:655-725)
//   compact           per-cell list of entries whose SNP has GPs     (cmd_cram_demuxlet.cpp:733)
//   K2s singlet       llksAB[j][0][0] for every sample j              (:733-747 at k=0,n=0; :806,:828)
//   K2 score          llksAB[j][k][n>=1] for every ordered pair x alpha, reduced in-kernel to
//                     log-sum-exp partials + top-2 candidates        (:733-747, :809-821, :883-906)
//   K3 finalize       per-droplet merge -> ccu_demux_result           (:788-906)
//
// Arithmetic notes.  K1 and gp_* reproduce the reference's operation order with explicit
// round-to-nearest intrinsics (no FMA contraction), so their outputs are bit-identical to the CPU
// loop.  K2 uses the exact factorisation  S[j,k,n] = sum_l g_j[l] * T[k,n,l],
// T[k,n,l] = sum_m g_k[m] * pG[n,l,m]  and accumulates  prod_snp S  as (mantissa, exponent)
// instead of  sum_snp log S : every factor lies in [~1e-10, ~1], the mantissa is renormalised
// into [1,2) by integer exponent extraction whenever an underflow budget of 960 binary orders of
// magnitude is used up (about every 29 general or 96 single-read SNPs), and a single log per
// reported value is taken at the very end.  This is the same real number as the reference's sum of logs to
// ~1e-13 relative (FP64 throughout).
#include <cuda_runtime.h>

#include <cstdlib>
#include <math_constants.h>
#include <stdint.h>

#include <cub/device/device_scan.cuh>

#include "demux_internal.h"
#include "div_exact.cuh"

namespace ccu {

// ------------------------------------------------------------------------------------------
// small device helpers
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes)
               : "memory");
}
#ifndef CCU_PRODUCER_SLEEP_NS
#define CCU_PRODUCER_SLEEP_NS 200
#endif
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  do {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
  } while (!ok);
}
// The producer's wait for a free ring stage: it has a chunk or more of slack, and a bare try_wait loop came back every
// ~24 cycles -- 7 % of the kernel's instructions, all issued on the one sub-partition the producer warp shares with four
// consumer warps.  Sleeping between the tries takes them out of the instruction count; K2's time did not move
// (18.69 ms at C2 with 50, 200 or 1000 ns), so the spins were not what the consumers waited for.
__device__ __forceinline__ void mbar_wait_relaxed(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  while (true) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
    if (ok) break;
    __nanosleep(CCU_PRODUCER_SLEEP_NS);
  }
}
// TMA 1-D bulk copy global -> shared, completion counted in bytes on an mbarrier (SASS: UBLKCP)
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
      "l"(src), "r"(bytes), "r"(bar)
      : "memory");
}
__device__ __forceinline__ void fence_proxy_async() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}

// 2^d for d <= 0 (0 below the normal range)
__device__ __forceinline__ double pow2_nonpos(int d) {
  return (d < -1022) ? 0.0 : __hiloint2double((1023 + d) << 20, 0);
}

// Renormalise a positive product into [1,2), moving its binary exponent into e.
// Zero / subnormal products (possible only when a genotype row sums to ~0) become exactly 0.
__device__ __forceinline__ void renorm(double& m, int& e) {
  int hi = __double2hiint(m);
  int ex = (hi >> 20) & 0x7ff;
  if (ex == 0) {
    m = 0.0;
  } else {
    e += ex - 1023;
    m = __hiloint2double((hi & 0x800fffff) | 0x3ff00000, __double2loint(m));
  }
}

// The same inside the hot loop, where only the exponent has to come down: a zero stays zero and a subnormal stays
// what it is (every later factor is <= 1, so it ends as exactly 0 in the epilogue's full renormalisation).
__device__ __forceinline__ void renorm_live(double& m, int& e) {
  const int hi = __double2hiint(m);
  const int ex = hi >> 20;  // products are non-negative: no sign bit
  if (ex != 0) {
    e += ex - 1023;
    m = __hiloint2double((hi & 0x000fffff) | 0x3ff00000, __double2loint(m));
  }
}

__device__ __forceinline__ bool cand_better(const Cand& a, const Cand& b) {
  if (a.e != b.e) return a.e > b.e;
  if (a.m != b.m) return a.m > b.m;
  return (uint32_t)a.pos < (uint32_t)b.pos;  // empty (-1) sorts last
}
__device__ __forceinline__ void cand_insert(Cand& b1, Cand& b2, const Cand& x) {
  if (cand_better(x, b1)) {
    b2 = b1;
    b1 = x;
  } else if (cand_better(x, b2)) {
    b2 = x;
  }
}
__device__ __forceinline__ Cand cand_empty() {
  Cand c;
  c.m = 0.0;
  c.e = INT_MIN;
  c.pos = -1;
  return c;
}
__device__ __forceinline__ Cand cand_shfl_xor(const Cand& c, int d) {
  Cand r;
  r.m = __shfl_xor_sync(0xffffffffu, c.m, d);
  r.e = __shfl_xor_sync(0xffffffffu, c.e, d);
  r.pos = __shfl_xor_sync(0xffffffffu, c.pos, d);
  return r;
}
// running sum  A * 2^E  of non-negative terms
__device__ __forceinline__ void lse_add(double& A, int& E, double m, int e) {
  if (m <= 0.0) return;
  if (e > E) {
    A = A * pow2_nonpos(E - e) + m;
    E = e;
  } else {
    A += m * pow2_nonpos(e - E);
  }
}
constexpr int kLseEmpty = INT_MIN / 2;

// ------------------------------------------------------------------------------------------
// genotype preparation (sc_drop_seq.cpp:287-315)
// ------------------------------------------------------------------------------------------
// avg_out[s] = {avg0, avg1, avg2, err_clamped}; one thread per SNP because the reference's
// running sums over samples are sequential (bit-exactness).
__global__ void gp_avg_kernel(const float* __restrict__ gp, const double* __restrict__ snp_err,
                              const uint8_t* __restrict__ has_gp, int64_t nsnps, int nv,
                              double* __restrict__ avg_out) {
  int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= nsnps) return;
  double a0 = 1e-10, a1 = 1e-10, a2 = 1e-10;
  double err = 0;
  if (has_gp == nullptr || has_gp[s]) {
    const float* f = gp + s * (int64_t)nv * 3;
    for (int i = 0; i < nv; ++i) {
      a0 = __dadd_rn(a0, (double)f[3 * i]);
      a1 = __dadd_rn(a1, (double)f[3 * i + 1]);
      a2 = __dadd_rn(a2, (double)f[3 * i + 2]);
    }
    double sum = __dadd_rn(__dadd_rn(a0, a1), a2);
    a0 = a0 / sum;
    a1 = a1 / sum;
    a2 = a2 / sum;
    err = snp_err[s];
    if (err > 0.999) err = 0.999;
    if (err < 0) err = 0;
  }
  avg_out[s * 4 + 0] = a0;
  avg_out[s * 4 + 1] = a1;
  avg_out[s * 4 + 2] = a2;
  avg_out[s * 4 + 3] = err;
}

// G[s][i][g] = (1-err)*gp + err*avg[g]   (padded samples and SNPs without GP: 0)
__global__ void gp_mix_kernel(const float* __restrict__ gp, const double* __restrict__ avg,
                              const uint8_t* __restrict__ has_gp, int64_t nsnps, int nv, int nvp,
                              double* __restrict__ G) {
  const int64_t row = (int64_t)nvp * 3;
  const int64_t total = nsnps * row;
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (int64_t)gridDim.x * blockDim.x) {
    int64_t s = idx / row;
    int r = (int)(idx - s * row);
    int i = r / 3, g = r - 3 * i;
    double v = 0.0;
    if (i < nv && (has_gp == nullptr || has_gp[s])) {
      v = (double)gp[s * (int64_t)nv * 3 + r];
      double err = avg[s * 4 + 3];
      if (err > 0) v = __dadd_rn(__dmul_rn(1 - err, v), __dmul_rn(err, avg[s * 4 + g]));
    }
    G[idx] = v;
  }
}

cudaError_t launch_gp_prepare(const float* gp_f32, const double* snp_err, const uint8_t* has_gp,
                              int64_t nsnps, int nv, int nvp, double* avg_tmp, double* G,
                              cudaStream_t st) {
  if (nsnps == 0) return cudaSuccess;
  gp_avg_kernel<<<(unsigned)((nsnps + 127) / 128), 128, 0, st>>>(gp_f32, snp_err, has_gp, nsnps, nv,
                                                                avg_tmp);
  int64_t total = nsnps * (int64_t)nvp * 3;
  int64_t blocks = (total + 255) / 256;
  if (blocks > 148 * 64) blocks = 148 * 64;
  gp_mix_kernel<<<(unsigned)blocks, 256, 0, st>>>(gp_f32, avg_tmp, has_gp, nsnps, nv, nvp, G);
  return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------
// K1: per-(cell,SNP) tile (cmd_cram_demuxlet.cpp:655-725)
// ------------------------------------------------------------------------------------------
// K1a (thread per entry) scans the entry's reads and classifies it.  With at most one informative
// read (allele code != 2) the tile is an affine function of the allele fraction
// p = (1-alpha)*l/2 + alpha*m/2, i.e. pG[n][l][m] = (1-alpha)*v(l) + alpha*v(m) with
// v(l) = pG[0][l][.], so only v is needed downstream: K1a computes it with the reference's own
// operation order (the running maximum of :686-687 over all 9*nA values is attained at p = 0 or
// p = 1, both of which are in the alpha[0] block) and stores it as vaff[e] = {v0, v1, v2, 0}.
// Entries with two or more informative reads are left to K1b.
__global__ void __launch_bounds__(256)
    demux_entry_kernel(int64_t nnz, const int64_t* __restrict__ read_ptr, const uint8_t* __restrict__ al,
                       const uint8_t* __restrict__ bq, const double* __restrict__ phred,
                       double* __restrict__ vaff, uint8_t* __restrict__ eclass, int32_t* __restrict__ maxbq) {
  const int lane = threadIdx.x & 31;
  int qmax = 0;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < nnz; e += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r0 = read_ptr[e], r1 = read_ptr[e + 1];
    int nuse = 0, a1 = 2, q1 = 0;
    for (int64_t r = r0; r < r1; ++r) {
      const int a = al[r];
      if (a != 2) {  // :664
        const int q = bq[r];
        if (nuse == 0) {
          a1 = a;
          q1 = q;
        }
        ++nuse;
        qmax = max(qmax, q);
      }
    }
    eclass[e] = (nuse > 1) ? 1 : 0;
    if (nuse <= 1) {
      double v0 = 1.0, v1 = 1.0, v2 = 1.0;  // :657
      if (nuse == 1) {
        const double er = phred[q1], mt = phred[256 + q1];
        const double pR = (a1 == 0) ? mt : er / 3.0;  // :666
        const double pA = (a1 == 1) ? mt : er / 3.0;  // :667
        // :673 at alpha[0] = 0: p = 0.5*l, 1-p
        v0 = __dmul_rn(v0, __dadd_rn(__dmul_rn(pR, 1.0), __dmul_rn(pA, 0.0)));  // :685
        v1 = __dmul_rn(v1, __dadd_rn(__dmul_rn(pR, 0.5), __dmul_rn(pA, 0.5)));
        v2 = __dmul_rn(v2, __dadd_rn(__dmul_rn(pR, 0.0), __dmul_rn(pA, 1.0)));
        const double mx = fmax(v0, fmax(v1, v2));
        v0 = v0 / mx;  // :696
        v1 = v1 / mx;
        v2 = v2 / mx;
      }
      v0 = __dadd_rn(v0, 1e-10);  // :704-725
      v1 = __dadd_rn(v1, 1e-10);
      v2 = __dadd_rn(v2, 1e-10);
      const double mx = fmax(v0, fmax(v1, v2));
      double4 o;
      o.x = v0 / mx;
      o.y = v1 / mx;
      o.z = v2 / mx;
      o.w = 0.0;
      reinterpret_cast<double4*>(vaff)[e] = o;
    }
  }
#pragma unroll
  for (int d = 16; d >= 1; d >>= 1) qmax = max(qmax, __shfl_xor_sync(0xffffffffu, qmax, d));
  if (lane == 0 && qmax > 0) atomicMax(maxbq, qmax);  // bounds the smallest single-read factor for K2
}

// K1b: the full 9 x nAlpha tile of the entries with two or more informative reads (every entry when `all_entries`
// is set: ccu_demux_get_tiles).  One entry per warp at a time: its 9 * nA values are spread over the 32 lanes (NS
// per lane), so the per-read chain -- 9 * nA multiply-adds, the running maximum over all of them (:686-687), 9 * nA
// divisions by it (:692-699) -- costs NS multiply-adds, one warp-wide maximum and NS divisions per lane, and a
// warp's trip count is the read count of ITS entry (a lane-per-alpha layout with two entries per warp ran
// max(reads) trips on 11 of 16 lanes).  The maximum of non-negative doubles is the maximum of their bit patterns:
// two integer warp reductions (REDUX) on the high and low words.  The reads of an entry are fetched 32 at a time,
// one per lane, and handed round by shuffles; the Phred tables sit in shared memory (err / 3 precomputed: the same
// IEEE division the reference does per read).  Rows are staged through shared memory so that the warp writes each
// one as a contiguous run.
//
// The cost of an entry is its read count (hundreds on a hot SNP), so the warps do not own fixed slices: each takes
// the next batch of `batch` 32-entry groups from a global counter when it is done (a static split left 12 % of the
// warp slots busy on skewed coverage: every CTA waited for its deepest entry).
constexpr int kTileWarps = 4;
constexpr int kTileOcc = 8;

template <int NS>
__global__ void __launch_bounds__(kTileWarps * 32, kTileOcc)
    demux_tile_kernel(int64_t nnz, const int64_t* __restrict__ read_ptr, const uint8_t* __restrict__ al,
                      const uint8_t* __restrict__ bq, int nA, const double* __restrict__ ptab,
                      const double* __restrict__ phred, const uint8_t* __restrict__ eclass, int all_entries,
                      const int32_t* __restrict__ tslot, double* __restrict__ tiles, double* __restrict__ vaff,
                      unsigned long long* __restrict__ next_batch, int batch) {
  __shared__ double s_err3[256];
  __shared__ double s_mat[256];
  __shared__ double s_out[kTileWarps][CCU_MAX_ALPHA * kTileStride];
  for (int i = threadIdx.x; i < 256; i += blockDim.x) {
    s_err3[i] = phred[i] / 3.0;  // :666-667
    s_mat[i] = phred[256 + i];
  }
  __syncthreads();

  constexpr uint32_t kFull = 0xffffffffu;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int nval = 9 * nA;
  double p[NS], omp[NS];
  int opos[NS];  // where the value goes in the staged row: alpha * kTileStride + (l, m)
#pragma unroll
  for (int s = 0; s < NS; ++s) {
    const int vi = s * 32 + lane;
    const bool ok = vi < nval;
    p[s] = ok ? ptab[vi] : 0.0;
    omp[s] = ok ? ptab[nval + vi] : 0.0;
    opos[s] = ok ? (vi / 9) * kTileStride + vi % 9 : -1;
  }
  const int rowlen = nA * kTileStride;
  double* out = s_out[warp];

  for (;;) {
    unsigned long long b0 = 0;
    if (lane == 0) b0 = atomicAdd(next_batch, 1ull);
    b0 = __shfl_sync(kFull, b0, 0);
    const int64_t first = (int64_t)b0 * batch * 32;
    if (first >= nnz) break;
    const int64_t last = min(nnz, first + (int64_t)batch * 32);
    for (int64_t base = first; base < last; base += 32) {
      const int64_t el = base + lane;
      unsigned todo = __ballot_sync(kFull, (el < nnz) && (all_entries || eclass[el] != 0));
      while (todo != 0u) {
        const int64_t e = base + (__ffs(todo) - 1);
        todo &= todo - 1;
        const int64_t r0 = read_ptr[e];
        const int cnt = (int)(read_ptr[e + 1] - r0);
        double v[NS];
#pragma unroll
        for (int s = 0; s < NS; ++s) v[s] = 1.0;  // :657
        for (int i0 = 0; i0 < cnt; i0 += 32) {
          const int mine = (i0 + lane < cnt) ? ((int)al[r0 + i0 + lane] | ((int)bq[r0 + i0 + lane] << 8)) : 2;
          const int nhere = min(32, cnt - i0);
          for (int i = 0; i < nhere; ++i) {
            const int aq = __shfl_sync(kFull, mine, i);
            const int a = aq & 0xff, q = aq >> 8;
            if (a == 2) continue;  // :664 (warp-uniform)
            const double pR = (a == 0) ? s_mat[q] : s_err3[q];  // :666
            const double pA = (a == 1) ? s_mat[q] : s_err3[q];  // :667
            double mx = 0.0;
#pragma unroll
            for (int s = 0; s < NS; ++s) {
              if (opos[s] >= 0) {
                v[s] = __dmul_rn(v[s], __dadd_rn(__dmul_rn(pR, omp[s]), __dmul_rn(pA, p[s])));  // :685
                mx = fmax(mx, v[s]);
              }
            }
            // the largest of all 9 * nA values (:686-687)
            const unsigned hi = (unsigned)__double2hiint(mx);
            const unsigned ghi = __reduce_max_sync(kFull, hi);
            const unsigned glo = __reduce_max_sync(kFull, (hi == ghi) ? (unsigned)__double2loint(mx) : 0u);
            mx = __hiloint2double((int)ghi, (int)glo);
            // :696 -- NS IEEE divisions by one divisor: the reciprocal is refined once (div_exact.cuh; bit-identical
            // to v / mx, which a lane takes instead whenever a value has left [2^-500, 1])
            unsigned lo = ghi;
#pragma unroll
            for (int s = 0; s < NS; ++s)
              if (opos[s] >= 0) lo = min(lo, (unsigned)__double2hiint(v[s]));
            if (lo >= kDivLoHi) {
              const double y = rcp_refined(mx);
#pragma unroll
              for (int s = 0; s < NS; ++s) v[s] = div_with_rcp(v[s], mx, y);
            } else {
#pragma unroll
              for (int s = 0; s < NS; ++s) v[s] = v[s] / mx;
            }
          }
        }
        // :704-725
        double mx = 0.0;
#pragma unroll
        for (int s = 0; s < NS; ++s) {
          if (opos[s] >= 0) {
            v[s] = __dadd_rn(v[s], 1e-10);
            mx = fmax(mx, v[s]);
          }
        }
        {
          const unsigned hi = (unsigned)__double2hiint(mx);
          const unsigned ghi = __reduce_max_sync(kFull, hi);
          const unsigned glo = __reduce_max_sync(kFull, (hi == ghi) ? (unsigned)__double2loint(mx) : 0u);
          mx = __hiloint2double((int)ghi, (int)glo);
        }
#pragma unroll
        for (int s = 0; s < NS; ++s)
          if (opos[s] >= 0) out[opos[s]] = v[s] / mx;
        for (int n = lane; n < nA; n += 32) out[n * kTileStride + 9] = 0.0;
        __syncwarp();
        if (lane == 0) {
          double4 w;
          w.x = out[0];
          w.y = out[3];
          w.z = out[6];
          w.w = 0.0;
          reinterpret_cast<double4*>(vaff)[e] = w;
        }
        double* dst = tiles + (tslot ? (int64_t)tslot[e] : e) * rowlen;
        for (int i = lane; i < rowlen; i += 32) dst[i] = out[i];
        __syncwarp();
      }
    }
  }
}

// ---- input check: the droplet store is caller data and every later kernel indexes with it.  bit 0: cell_ptr is
// not a non-decreasing walk from 0 to nnz; bit 1: a SNP id outside the genotype matrix; bit 2: read_ptr is not a
// non-decreasing walk from 0 to nreads.  Pure streaming over the three index arrays.
__global__ void __launch_bounds__(256)
    demux_validate_kernel(int64_t nbcs, int64_t nnz, int64_t nreads, int64_t nsnps, const int64_t* __restrict__ cell_ptr,
                          const int32_t* __restrict__ snp_idx, const int64_t* __restrict__ read_ptr,
                          uint32_t* __restrict__ flag) {
  uint32_t bad = 0;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nbcs; i += stride) {
    const int64_t a = cell_ptr[i], b = cell_ptr[i + 1];
    if (a > b || a < 0 || b > nnz || (i == 0 && a != 0) || (i == nbcs - 1 && b != nnz)) bad |= 1u;
  }
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nnz; i += stride) {
    const int32_t s = snp_idx[i];
    if (s < 0 || s >= nsnps) bad |= 2u;
    const int64_t a = read_ptr[i], b = read_ptr[i + 1];
    if (a > b || a < 0 || b > nreads || (i == 0 && a != 0) || (i == nnz - 1 && b != nreads)) bad |= 4u;
  }
  bad = __reduce_or_sync(0xffffffffu, bad);
  if (bad && (threadIdx.x & 31) == 0) atomicOr(flag, bad);
}

cudaError_t launch_validate(const DemuxDev& d, int64_t nreads, uint32_t* flag, cudaStream_t st) {
  const int64_t n = d.nbcs > d.nnz ? d.nbcs : d.nnz;
  if (n == 0) return cudaSuccess;
  int64_t blocks = (n + 255) / 256;
  if (blocks > 148 * 8) blocks = 148 * 8;
  demux_validate_kernel<<<(unsigned)blocks, 256, 0, st>>>(d.nbcs, d.nnz, nreads, d.nsnps, d.cell_ptr, d.snp_idx, d.read_ptr, flag);
  return cudaGetLastError();
}

// p[i] -= base: the pointer arrays of a store's second part arrive with the offsets of the whole store
__global__ void __launch_bounds__(256) demux_rebase_kernel(int64_t* __restrict__ p, int64_t n, int64_t base) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) p[i] -= base;
}

cudaError_t launch_rebase(int64_t* p, int64_t n, int64_t base, cudaStream_t st) {
  if (n <= 0 || base == 0) return cudaSuccess;
  int64_t blocks = (n + 255) / 256;
  if (blocks > 148 * 8) blocks = 148 * 8;
  demux_rebase_kernel<<<(unsigned)blocks, 256, 0, st>>>(p, n, base);
  return cudaGetLastError();
}

// entries with >= 2 informative reads (the same test as demux_entry_kernel), counted once per upload
__global__ void __launch_bounds__(256)
    demux_count_general_kernel(int64_t nnz, const int64_t* __restrict__ read_ptr, const uint8_t* __restrict__ al,
                               unsigned long long* __restrict__ count) {
  unsigned n = 0;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < nnz; e += (int64_t)gridDim.x * blockDim.x) {
    int nuse = 0;
    for (int64_t r = read_ptr[e]; r < read_ptr[e + 1] && nuse < 2; ++r) nuse += (al[r] != 2);
    n += (nuse > 1);
  }
  n = __reduce_add_sync(0xffffffffu, n);
  if ((threadIdx.x & 31) == 0 && n) atomicAdd(count, (unsigned long long)n);
}

cudaError_t launch_count_general(const DemuxDev& d, unsigned long long* count_dev, cudaStream_t st) {
  cudaError_t e0 = cudaMemsetAsync(count_dev, 0, sizeof(unsigned long long), st);
  if (e0 != cudaSuccess || d.nnz == 0) return e0;
  int64_t blocks = (d.nnz + 255) / 256;
  if (blocks > 148 * 32) blocks = 148 * 32;
  demux_count_general_kernel<<<(unsigned)blocks, 256, 0, st>>>(d.nnz, d.read_ptr, d.al, count_dev);
  return cudaGetLastError();
}

cudaError_t launch_tile(const DemuxDev& d, bool all_entries, void* scan_tmp, size_t scan_bytes, size_t* scan_need,
                        cudaStream_t st) {
  if (scan_bytes == 0) {  // size query of the prefix sum
    size_t need = 0;
    cub::DeviceScan::ExclusiveSum(nullptr, need, (const uint8_t*)nullptr, (int32_t*)nullptr, (int)(d.nnz > 0 ? d.nnz : 1), st);
    if (scan_need) *scan_need = need;
    return cudaSuccess;
  }
  // maxbq[0] and, in the same 16-byte buffer, the tile kernel's batch counter
  cudaError_t e0 = cudaMemsetAsync(d.maxbq, 0, 16, st);
  if (e0 != cudaSuccess) return e0;
  if (d.nnz == 0) return cudaSuccess;
  {
    int64_t blocks = (d.nnz + 255) / 256;
    if (blocks > 148 * 32) blocks = 148 * 32;
    demux_entry_kernel<<<(unsigned)blocks, 256, 0, st>>>(d.nnz, d.read_ptr, d.al, d.bq, d.phred, d.vaff, d.eclass,
                                                         d.maxbq);
  }
  if (!all_entries) {  // tile row of every general entry = number of general entries before it
    e0 = cub::DeviceScan::ExclusiveSum(scan_tmp, scan_bytes, d.eclass, d.tslot, (int)d.nnz, st);
    if (e0 != cudaSuccess) return e0;
  }
  // persistent warps (kTileOcc CTAs of kTileWarps per SM) fed from a counter; batches of 32-entry groups: one group
  // where a sizeable share of the entries needs a tile, 32 groups where hardly any does (fewer trips to the counter)
  const int batch = (all_entries || d.ngen * 64 > d.nnz) ? 1 : 32;
  int64_t per_block = (int64_t)kTileWarps * 32 * batch;
  int64_t blocks = (d.nnz + per_block - 1) / per_block;
  if (blocks > 148 * kTileOcc) blocks = 148 * kTileOcc;
  unsigned long long* next_batch = reinterpret_cast<unsigned long long*>(d.maxbq) + 1;
  const int ns = (9 * d.nA + 31) / 32;  // values per lane
#define CCU_TILE(NS)                                                                                 \
  demux_tile_kernel<NS><<<(unsigned)blocks, kTileWarps * 32, 0, st>>>(d.nnz, d.read_ptr, d.al, d.bq, d.nA, d.ptab, \
                                                                       d.phred, d.eclass, all_entries ? 1 : 0,      \
                                                                       all_entries ? nullptr : d.tslot, d.tiles, d.vaff, \
                                                                       next_batch, batch)
  if (ns <= 1) CCU_TILE(1);
  else if (ns <= 2) CCU_TILE(2);
  else if (ns <= 3) CCU_TILE(3);
  else if (ns <= 4) CCU_TILE(4);
  else if (ns <= 6) CCU_TILE(6);
  else CCU_TILE(9);
#undef CCU_TILE
  return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------
// compaction: per cell, the entries whose SNP has genotypes (:733), general entries first and
// the affine (<= 1 informative read) entries after them:
//   vent[cell_ptr[c] + r], r in [0, vcnt_g[c])                    general
//   vent[cell_ptr[c] + vcnt_g[c] + r], r in [0, vcnt_a[c])        affine
// ------------------------------------------------------------------------------------------
__global__ void compact_kernel(int64_t nbcs, const int64_t* __restrict__ cell_ptr,
                               const int32_t* __restrict__ snp_idx, const uint8_t* __restrict__ has_gp,
                               const uint8_t* __restrict__ eclass, int64_t* __restrict__ vent,
                               int32_t* __restrict__ vsnp, int32_t* __restrict__ vcnt_g,
                               int32_t* __restrict__ vcnt_a) {
  const int lane = threadIdx.x & 31;
  const int64_t c = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (c >= nbcs) return;
  const int64_t b = cell_ptr[c], e = cell_ptr[c + 1];
  int run = 0, ngen = 0;
  for (int pass = 0; pass < 2; ++pass) {  // pass 0: general, pass 1: affine
    for (int64_t i = b; i < e; i += 32) {
      const int64_t idx = i + lane;
      bool ok = false;
      int32_t snp = 0;
      if (idx < e) {
        snp = snp_idx[idx];
        ok = ((has_gp == nullptr) || has_gp[snp]) && ((eclass[idx] != 0) == (pass == 0));
      }
      const uint32_t mask = __ballot_sync(0xffffffffu, ok);
      if (ok) {
        const int64_t o = b + run + __popc(mask & ((1u << lane) - 1u));
        vent[o] = idx;
        vsnp[o] = snp;
      }
      run += __popc(mask);
    }
    if (pass == 0) ngen = run;
  }
  if (lane == 0) {
    vcnt_g[c] = ngen;
    vcnt_a[c] = run - ngen;
  }
}

cudaError_t launch_compact(const DemuxDev& d, cudaStream_t st) {
  if (d.nbcs == 0) return cudaSuccess;
  int64_t blocks = (d.nbcs * 32 + 255) / 256;
  compact_kernel<<<(unsigned)blocks, 256, 0, st>>>(d.nbcs, d.cell_ptr, d.snp_idx, d.has_gp, d.eclass, d.vent,
                                                   d.vsnp, d.vcnt_g, d.vcnt_a);
  return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------
// K2s: per-(droplet, sample) prepass
//   * singlet likelihoods llksAB[j][0][0] (cmd_cram_demuxlet.cpp:733-747 at k=0, n=0).  alpha[0]==0
//     makes the tile independent of m, but the reference still multiplies by sample 0's genotype
//     row, so T0 uses g_0;
//   * for the affine entries: f_j = F_j / s_j with F_j = sum_l g_j[l]*pG[0][l][0] and
//     s_j = sum_l g_j[l], written as [32-sample tile][entry][32] so that the scoring kernel streams
//     a (tile x entry chunk) block with one bulk copy, and the per-droplet product of the s_j.
//     For such an entry sum_{l,m} g_j[l] g_k[m] pG[n][l][m] = s_j s_k ((1-alpha_n) f_j + alpha_n f_k).
// One warp per (droplet, 32-sample tile), lanes over samples.
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
    demux_sample_kernel(DemuxDev d) {
  const int lane = threadIdx.x & 31;
  const int ntile = d.nvp >> 5;
  const int64_t w = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t c = w / ntile;
  if (c >= d.nbcs) return;
  const int tile = (int)(w - c * ntile);
  const int j = tile * 32 + lane;
  const int64_t base = d.cell_ptr[c];
  const int ng = d.vcnt_g[c], na = d.vcnt_a[c];
  const int64_t grow = (int64_t)d.nvp * 3;
  double* fdst = d.fmat + base * d.nvp + (int64_t)tile * 32 * na + lane;
  double m = 1.0, pm = 1.0;
  int ex = 0, pe = 0;
  const int cnt = ng + na;
#pragma unroll 2
  for (int i = 0; i < cnt; ++i) {
    const int64_t e = d.vent[base + i];
    // alpha[0] == 0 makes pG[0][l][m] independent of m (bitwise: the same p = 0.5*l), so the three
    // values v(l) = pG[0][l][.] stand for the whole alpha[0] block
    const double4 v = reinterpret_cast<const double4*>(d.vaff)[e];
    const double* g = d.G + (int64_t)d.vsnp[base + i] * grow;
    const double g00 = g[0], g01 = g[1], g02 = g[2];
    const double t0 = fma(g02, v.x, fma(g01, v.x, g00 * v.x));
    const double t1 = fma(g02, v.y, fma(g01, v.y, g00 * v.y));
    const double t2 = fma(g02, v.z, fma(g01, v.z, g00 * v.z));
    const double* gj = g + 3 * j;
    const double gj0 = gj[0], gj1 = gj[1], gj2 = gj[2];
    m *= fma(gj2, t2, fma(gj1, t1, gj0 * t0));
    if (i >= ng) {
      const double F = fma(gj2, v.z, fma(gj1, v.y, gj0 * v.x));
      const double sj = (gj0 + gj1) + gj2;
      const bool live = (j < d.nv) && (sj > 0.0);
      fdst[(int64_t)(i - ng) * 32] = live ? F / sj : ((j < d.nv) ? 0.0 : 1.0);
      pm *= (j < d.nv) ? sj : 1.0;
    }
    if ((i & 15) == 15) {
      renorm(m, ex);
      renorm(pm, pe);
    }
  }
  renorm(m, ex);
  renorm(pm, pe);
  if (j < d.nv)
    d.sng_llk[c * d.nv + j] = (m > 0.0) ? (log(m) + (double)ex * 0.6931471805599453094) : -CUDART_INF;
  d.pm[c * d.nvp + j] = pm;
  d.pe[c * d.nvp + j] = pe;
}

cudaError_t launch_sample(const DemuxDev& d, cudaStream_t st) {
  if (d.nbcs == 0) return cudaSuccess;
  int64_t blocks = (d.nbcs * (d.nvp >> 5) * 32 + 255) / 256;
  demux_sample_kernel<<<(unsigned)blocks, 256, 0, st>>>(d);
  return cudaGetLastError();
}

__global__ void singlet_to_dbg_kernel(DemuxDev d, double* __restrict__ llk, int64_t c0, int64_t c1) {
  int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  int64_t n = (c1 - c0) * d.nv;
  if (idx >= n) return;
  int64_t c = idx / d.nv;
  int j = (int)(idx - c * d.nv);
  llk[((c * d.nv + j) * d.nv + 0) * d.nA + 0] = d.sng_llk[(c0 + c) * d.nv + j];
}

cudaError_t launch_singlet_to_dbg(const DemuxDev& d, double* llk_dbg, int64_t c0, int64_t c1,
                                  cudaStream_t st) {
  int64_t n = (c1 - c0) * d.nv;
  if (n <= 0) return cudaSuccess;
  singlet_to_dbg_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(d, llk_dbg, c0, c1);
  return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------
// K2: pair x alpha scoring (cmd_cram_demuxlet.cpp:733-747, reduced per :809-821 and :883-906)
// ------------------------------------------------------------------------------------------
// Work item = (droplet, j-tile of 32 samples, k-tile of 32*TK samples, chunk of AC doublet
// alphas).  A persistent CTA (one per SM) walks its items as ONE stream of entry chunks:
//   * the producer warp streams the chunks into a shared-memory ring with TMA bulk copies
//     (cp.async.bulk, completion on "full" mbarriers) and publishes each chunk's metadata next to
//     it; consumers hand stages back through "empty" mbarriers, so the ring never drains across
//     item boundaries and consumer warps never touch global memory for control flow;
//   * 16 consumer warps; each thread owns 2 samples j x (TK x AC) columns = 20 running products
//     in registers.  Two kinds of chunk feed them:
//       general  (entries with >= 2 informative reads): per SNP the j-rows and k-rows of the FP64
//                genotype matrix and the alpha block of the droplet's tile; the k-rows x tile
//                block becomes T[k][n][l] (double-buffered smem), then 2 DFMA + 2 DMUL per
//                (SNP, j, k, alpha);
//       affine   (<= 1 informative read, the bulk of sparse single-cell data): the tile is affine
//                in the allele fraction, so the term is s_j s_k ((1-alpha) f_j + alpha f_k).  The
//                chunk is two (sample tile x entries) blocks of the prepass' f matrix, and the
//                inner loop is 1 DFMA + 1 DMUL per (SNP, j, k, alpha) plus one DADD per (j, k):
//                f_j + alpha (f_k - f_j).  The s_j s_k products are per-droplet constants applied
//                in the item epilogue;
//   * at the end of an item the products become log-sum-exp / top-2 partials (no log needed:
//     ordering by (exponent, mantissa) is ordering by LLK; the epilogue rescales them to the warp's
//     largest exponent and works on plain doubles, see `epilogue` below);
//   * items are handed out by a global counter (the first one per CTA by its index), read one
//     item ahead by the producer warp.
#ifndef CCU_NS
#define CCU_NS 3     // ring stages of the in-step kernel
#endif
#ifndef CCU_SCA1
#define CCU_SCA1 64  // entries per affine chunk at TK == 1 (32: 19.07 ms at C2, 64: 18.83, 96 with two stages: 18.94)
#endif
// Consumer groups out of step (the LAGP template parameter, TK == 1 only).  The 16 consumer warps form two groups of
// eight by k half -- two warps of each group on every SM sub-partition -- and group 1 stays at least LAGP chunks
// behind the warps of group 0 that share its sub-partition: a group builds the T rows of its own k half and meets at
// its own barrier, so one group's T build, barrier and ring hand-over issue under the other group's DFMAs.
// Measured (B200): general chunks gain (skewed coverage, a third of the entries general: 6.54 -> 5.81 ms with LAGP = 1
// and a five-stage ring of 32-entry chunks), affine chunks do not (C2 within 1 %, C5 1.3 % slower, and the deeper ring
// costs the 64-entry chunks their room) -- whatever lag and ring depth, the item epilogue was NOT hidden behind the
// other group's products.  So the launch picks LAGP = 1 for stores with more than a tenth of
// their entries general and 0 otherwise; CCU_LAG_MODE forces it (0 / 1) for A/B builds.
#ifndef CCU_LAG_MODE
#define CCU_LAG_MODE -1
#endif
#ifndef CCU_SC1
#define CCU_SC1 8   // SNPs per general chunk at TK == 1 (4 halves the T buffers: room for a deeper ring)
#endif
constexpr int kNSWanted = CCU_NS;                   // raw ring stages (fewer where shared memory is short)
constexpr int kTJ = 2;                              // samples j per thread
constexpr int kConsumers = kScoreThreads;           // 16 consumer warps
constexpr int kScoreBlock = kScoreThreads + 128;    // + the producer warpgroup (1 working warp)
// Register budget: the kernel launches with 96 registers/thread (65536 / 640 rounded down to the
// allocation unit); the producer warpgroup shrinks to 32 and the four consumer warpgroups grow
// to 112 out of the CTA's own pool (16*32*112 + 4*32*32 = 61440 = 640*96).
constexpr int kRegsProducer = 32;
constexpr int kRegsConsumer = 112;
static_assert(kScoreThreads == 512, "thread mapping below assumes 16 consumer warps");

__device__ __forceinline__ void consumer_bar() {
  asm volatile("bar.sync 1, %0;" ::"n"(kConsumers) : "memory");
}
__device__ __forceinline__ void group_bar(int grp) {  // one of the two consumer groups (LAGP > 0)
  if (grp == 0) asm volatile("bar.sync 1, %0;" ::"n"(kConsumers / 2) : "memory");
  else asm volatile("bar.sync 2, %0;" ::"n"(kConsumers / 2) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}

struct ChunkMeta {   // published by the producer with each ring stage
  int32_t cnt;       // entries in this chunk
  int32_t flags;     // bit0 valid (0 = end of stream), bit1 last chunk of its item, bit2 affine
  int32_t cell;
  int32_t j0, k0, n0;
  int64_t item;
};

template <int AC, int TK, int LAGP = 0>
struct ScoreCfg {
  static constexpr int SC = (TK >= 8) ? 4 : (TK == 1) ? CCU_SC1 : 8;  // SNPs per general chunk (bounded by shared memory)
  static constexpr int LAG = (TK == 1) ? LAGP : 0;  // chunks group 1 stays behind group 0 (0: one group)
  static constexpr int SCA = (TK == 1) ? (LAG > 0 ? 32 : CCU_SCA1) : (TK <= 4) ? 16 : 8;  // entries per affine chunk
  static constexpr int NSW = (LAG > 0) ? 5 : kNSWanted;  // ring stages wanted
  static constexpr int NCOL = AC * TK;
  static constexpr int TROW0 = NCOL * 3;
  // row stride (doubles) of a thread's T block: even, and (stride/2) odd => the 8 distinct
  // 16-byte segments a warp reads per LDS.128 fall into distinct bank groups
  static constexpr int TROW = ((TROW0 / 2) % 2 == 1) ? TROW0 : TROW0 + 2;
  static constexpr int KT = 32 * TK;
  static constexpr int GJ_BYTES = kJT * 24;
  static constexpr int GK_BYTES = KT * 24;
  static constexpr int PG_BYTES = AC * kTileStride * 8;
  static constexpr int SLOT_BYTES = GJ_BYTES + GK_BYTES + PG_BYTES;
  static constexpr int PART_BYTES = SCA * 32 * 8;          // one (32-sample tile x SCA entries) block
  static constexpr int STAGE_G = SC * SLOT_BYTES;
  static constexpr int STAGE_A = (1 + TK) * PART_BYTES;
  static constexpr int STAGE_BYTES = (STAGE_G > STAGE_A) ? STAGE_G : STAGE_A;
  static constexpr int T_BYTES = SC * 32 * TROW * 8;
  static constexpr int RED_BYTES = (kConsumers / 32) * 64;
  static constexpr int FIXED_BYTES = 2 * T_BYTES + RED_BYTES + 48 * 8 + 64;
  static constexpr int NS = ((NSW * (STAGE_BYTES + 32 + 16) + FIXED_BYTES) <= 227 * 1024) ? NSW : 3;
  static constexpr int META_BYTES = NS * (int)sizeof(ChunkMeta);
  static constexpr int ALPHA_BYTES = 48 * 8;
  static constexpr int OFF_T = NS * STAGE_BYTES;
  static constexpr int OFF_RED = OFF_T + 2 * T_BYTES;
  static constexpr int OFF_META = OFF_RED + RED_BYTES;
  static constexpr int OFF_ALPHA = OFF_META + META_BYTES;
  static constexpr int OFF_BAR = OFF_ALPHA + ALPHA_BYTES;
  static constexpr int SMEM_BYTES = OFF_BAR + 2 * NS * 8 + 16;
  // T-compute task split: NG alpha groups per (SNP slot, k) pair, APT alphas per task.  TTHREADS threads share
  // TPAIRS pairs (a consumer group builds the rows of its own k half when the groups run out of step); NG is the
  // divisor of AC that needs the fewest alpha evaluations per thread, rounds counted whole.
  static constexpr int PAIRS = SC * KT;
  static constexpr int TTHREADS = (LAG > 0) ? kConsumers / 2 : kConsumers;
  static constexpr int TPAIRS = (LAG > 0) ? PAIRS / 2 : PAIRS;
  static constexpr int t_cost(int ng) { return ((TPAIRS * ng + TTHREADS - 1) / TTHREADS) * (AC / ng); }
  static constexpr int pick_ng() {
    int best = 1;
    for (int ng = 2; ng <= AC; ++ng)
      if (AC % ng == 0 && (ng & (ng - 1)) == 0 && t_cost(ng) < t_cost(best)) best = ng;  // power-of-two decode first
    for (int ng = 2; ng <= AC; ++ng)
      if (AC % ng == 0 && t_cost(ng) < t_cost(best)) best = ng;
    return best;
  }
  static constexpr int NG = pick_ng();
  static constexpr int APT = AC / NG;
  static_assert(NCOL % 2 == 0, "columns are processed in pairs");
  static_assert(TROW % 2 == 0, "16-byte aligned T rows");
  static_assert(AC % NG == 0, "alpha groups must divide the chunk");
  static_assert((KT & (KT - 1)) == 0, "power-of-two task decode");
  static_assert(STAGE_BYTES % 16 == 0 && OFF_T % 16 == 0, "bulk copy alignment");
  static_assert(sizeof(ChunkMeta) == 32, "ChunkMeta layout");
  static_assert(SMEM_BYTES <= 227 * 1024, "shared memory budget");
};

struct WarpRed {
  double lse_a;
  int32_t lse_e, _pad;
  Cand b1, b2;
};
static_assert(sizeof(WarpRed) <= 64, "WarpRed slot");

template <int AC, int TK, int LAGP>
__global__ void __launch_bounds__(kScoreBlock, 1)
    demux_score_kernel(DemuxDev d, int njt, int nkt, int nac, int64_t nitems, uint32_t half_mask,
                       double2* __restrict__ dbg, int64_t dbg_c0, int64_t dbg_c1,
                       unsigned long long* __restrict__ next_item, int force_exact) {
  using Cfg = ScoreCfg<AC, TK, LAGP>;
  constexpr int kSC = Cfg::SC;
  constexpr int kSCA = Cfg::SCA;
  constexpr int kNS = Cfg::NS;
  extern __shared__ __align__(128) unsigned char smem[];
  unsigned char* raw = smem;
  double* Tbuf = reinterpret_cast<double*>(smem + Cfg::OFF_T);
  ChunkMeta* meta = reinterpret_cast<ChunkMeta*>(smem + Cfg::OFF_META);
  double* s_alpha = reinterpret_cast<double*>(smem + Cfg::OFF_ALPHA);
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + Cfg::OFF_BAR);
  uint64_t* empty = full + kNS;

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

  // zero the ring once (partial k-tiles / alpha chunks leave their tail untouched)
  for (int i = tid; i < kNS * Cfg::STAGE_BYTES / 8; i += kScoreBlock) reinterpret_cast<double*>(raw)[i] = 0.0;
  if (tid < 48) s_alpha[tid] = (tid < d.nA) ? d.alpha[tid] : 0.0;
  if (tid < kConsumers / 32) reinterpret_cast<int*>(smem + Cfg::OFF_RED)[tid] = 0;
  if (tid == 0) {
    for (int s = 0; s < kNS; ++s) {
      mbar_init(smem_u32(&full[s]), 1);
      mbar_init(smem_u32(&empty[s]), kConsumers / 32);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  fence_proxy_async();
  __syncthreads();

  if (warp >= kConsumers / 32) {
    // ======================= producer warp: TMA bulk copies + chunk metadata =======================
    asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(kRegsProducer));
    if (warp != kConsumers / 32) return;
    const int ipc = njt * nkt * nac;
    // Items are handed out by a global counter (the first one per CTA by its index): droplets differ in depth, and
    // with fixed strides the whole grid waited for the CTA that drew the deepest ones.  The counter is read one item
    // ahead, so its round trip hides behind the item being streamed.
    int64_t item = blockIdx.x;
    uint32_t g = 0;
    while (true) {
      unsigned long long nxt = 0;
      if (lane == 0) nxt = atomicAdd(next_item, 1ull);
      const bool valid = item < nitems;
      int32_t cell = 0, ng = 0, na = 0, nchg = 0, nch = 1, j0 = 0, k0 = 0, n0 = 1;
      int64_t ent0 = 0;
      if (valid) {
        const int64_t c64 = item / ipc;
        int t = (int)(item - c64 * ipc);
        const int ac = t % nac;
        t /= nac;
        cell = (int32_t)c64;
        ent0 = d.cell_ptr[c64];
        ng = d.vcnt_g[c64];
        na = d.vcnt_a[c64];
        nchg = (ng + kSC - 1) / kSC;
        nch = max(1, nchg + (na + kSCA - 1) / kSCA);
        j0 = (t / nkt) * kJT;
        k0 = (t % nkt) * Cfg::KT;
        n0 = 1 + ac * AC;
      }
      const int krows = min(Cfg::KT, d.nvp - k0);
      const int nal = min(AC, d.nA - n0);
      const uint32_t kbytes = (uint32_t)krows * 24u, pbytes = (uint32_t)nal * (kTileStride * 8);
      for (int chunk = 0; chunk < nch; ++chunk, ++g) {
        const int stage = g % kNS;
        if (g >= (uint32_t)kNS) mbar_wait_relaxed(smem_u32(&empty[stage]), ((g / kNS) - 1) & 1);
        const uint32_t bar = smem_u32(&full[stage]);
        const bool affine = chunk >= nchg;
        int cnt;
        uint32_t tx;
        if (!affine) {
          cnt = valid ? max(0, min(kSC, ng - chunk * kSC)) : 0;
          tx = (uint32_t)cnt * (Cfg::GJ_BYTES + kbytes + pbytes);
        } else {
          cnt = valid ? max(0, min(kSCA, na - (chunk - nchg) * kSCA)) : 0;
          tx = (uint32_t)cnt * 256u * (uint32_t)(1 + krows / 32);
        }
        int64_t e = 0, snp = 0;
        if (!affine && lane < cnt) {
          const int64_t idx = ent0 + (int64_t)chunk * kSC + lane;
          e = d.vent[idx];
          snp = d.vsnp[idx];
        }
        if (lane == 0) {
          ChunkMeta m;
          m.cnt = cnt;
          m.flags = (valid ? 1 : 0) | ((chunk + 1 >= nch) ? 2 : 0) | (affine ? 4 : 0);
          m.cell = cell;
          m.j0 = j0;
          m.k0 = k0;
          m.n0 = n0;
          m.item = item;
          meta[stage] = m;
          mbar_expect_tx(bar, tx);  // release: meta visible
        }
        __syncwarp();
        if (!affine) {
          if (lane < cnt) {
            const uint32_t slot = smem_u32(raw + stage * Cfg::STAGE_BYTES + lane * Cfg::SLOT_BYTES);
            const double* grow = d.G + snp * (int64_t)d.nvp * 3;
            bulk_g2s(slot, grow + (int64_t)j0 * 3, Cfg::GJ_BYTES, bar);
            bulk_g2s(slot + Cfg::GJ_BYTES, grow + (int64_t)k0 * 3, kbytes, bar);
            bulk_g2s(slot + Cfg::GJ_BYTES + Cfg::GK_BYTES, d.tiles + ((int64_t)d.tslot[e] * d.nA + n0) * kTileStride, pbytes, bar);
          }
        } else if (cnt > 0 && lane <= krows / 32) {
          // part 0: the j tile; parts 1..: the 32-sample tiles of the k tile
          const int tile = (lane == 0) ? (j0 >> 5) : ((k0 >> 5) + lane - 1);
          const double* src = d.fmat + ent0 * d.nvp + ((int64_t)tile * na + (int64_t)(chunk - nchg) * kSCA) * 32;
          bulk_g2s(smem_u32(raw + stage * Cfg::STAGE_BYTES + lane * Cfg::PART_BYTES), src, (uint32_t)cnt * 256u, bar);
        }
      }
      if (!valid) break;  // the chunk just published was the end-of-stream sentinel
      item = (int64_t)gridDim.x + (int64_t)__shfl_sync(0xffffffffu, nxt, 0);
    }
    return;
  }

  // ======================= consumer warps =======================
  asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(kRegsConsumer));
  // Warps w, w+4, w+8, w+12 share an SM sub-partition.  In step (LAG == 0) they share the k-lanes and differ in j;
  // out of step they share the j-groups and differ in k, so that a group (k half = warp >> 3) owns its T rows.
  constexpr int kLag = Cfg::LAG;
  const int kidx = (kLag > 0) ? (warp >> 2) * 8 + (lane & 7) : (warp & 3) * 8 + (lane & 7);   // k-lane 0..31
  const int jg = (kLag > 0) ? (warp & 3) * 4 + (lane >> 3) : (warp >> 2) * 4 + (lane >> 3);   // j-group 0..15
  // the group is read through a lane-0 broadcast so that the compiler knows it is warp-uniform: a branch on a
  // value derived from threadIdx counts as divergent, and behind one the alphas of the pair loop no longer travel
  // as uniform-register operands (25.5 ms instead of 20.1 on C2)
  const int warp_u = (kLag > 0) ? __shfl_sync(0xffffffffu, warp, 0) : 0;
  const int cgrp = warp_u >> 3;
  volatile int* prog = reinterpret_cast<volatile int*>(smem + Cfg::OFF_RED);  // chunks finished, per warp of group 0

  // T[s][klane][(kk*AC + a)*3 + l] = sum_m g_k[m] * pG[a][l][m]
  auto compute_T = [&](int cnt, int stage, int tb) {
    double* T = Tbuf + tb * (Cfg::T_BYTES / 8);
    const unsigned char* st = raw + stage * Cfg::STAGE_BYTES;
    // out of step: a group builds the rows of its own k half with its own 256 threads
    constexpr int kThreadsT = Cfg::TTHREADS;
    constexpr int kKTg = (kLag > 0) ? Cfg::KT / 2 : Cfg::KT;
    constexpr int kTasks = Cfg::TPAIRS * Cfg::NG;
    const int tidg = (kLag > 0) ? (tid & (kConsumers / 2 - 1)) : tid;
#pragma unroll
    for (int u0 = 0; u0 < kTasks; u0 += kThreadsT) {
      const int u = u0 + tidg;
      const int grp = u % Cfg::NG;
      const int pr = u / Cfg::NG;
      const int kabs = (pr & (kKTg - 1)) + cgrp * kKTg;
      const int s = (u < kTasks) ? pr / kKTg : kSC;  // past the last task of a ragged round: no SNP slot
      if (s < cnt) {
        const unsigned char* slot = st + s * Cfg::SLOT_BYTES;
        const double* gk = reinterpret_cast<const double*>(slot + Cfg::GJ_BYTES) + kabs * 3;
        const double g0 = gk[0], g1 = gk[1], g2 = gk[2];
        const double* pgb = reinterpret_cast<const double*>(slot + Cfg::GJ_BYTES + Cfg::GK_BYTES);
        double* ob = T + (s * 32 + kabs / TK) * Cfg::TROW + ((kabs % TK) * AC) * 3;
#pragma unroll
        for (int i = 0; i < Cfg::APT; ++i) {
          const int a = grp * Cfg::APT + i;
          const double2* p2 = reinterpret_cast<const double2*>(pgb + a * kTileStride);
          const double2 q0 = p2[0], q1 = p2[1], q2 = p2[2], q3 = p2[3];
          const double q8 = pgb[a * kTileStride + 8];
          double* o = ob + a * 3;
          o[0] = fma(g2, q1.x, fma(g1, q0.y, g0 * q0.x));
          o[1] = fma(g2, q2.y, fma(g1, q2.x, g0 * q1.y));
          o[2] = fma(g2, q8, fma(g1, q3.y, g0 * q3.x));
        }
      }
    }
  };

  double acc[kTJ][Cfg::NCOL];
  int aex[kTJ][Cfg::NCOL];
  auto reset_acc = [&]() {
#pragma unroll
    for (int jj = 0; jj < kTJ; ++jj)
#pragma unroll
      for (int cc = 0; cc < Cfg::NCOL; ++cc) {
        acc[jj][cc] = 1.0;
        aex[jj][cc] = 0;
      }
  };
  auto renorm_all = [&]() {
#pragma unroll
    for (int jj = 0; jj < kTJ; ++jj)
#pragma unroll
      for (int cc = 0; cc < Cfg::NCOL; ++cc) renorm_live(acc[jj][cc], aex[jj][cc]);
  };
  // Underflow budget (binary orders of magnitude a running product may still lose before it must
  // be renormalised).  A general factor is >= ~1e-10 (the +1e-10 floor of :704-725) = 34 bits; an
  // affine factor is >= err(maxbq)/3, i.e. abits = ceil(0.3322*maxbq + 1.585) + 1, at most 34.
  constexpr int kBudget = 960, kGenBits = 34;
  const int abits = min(kGenBits, (int)(0.33220 * (double)d.maxbq[0] + 3.585));
  const int inv_abits = 65536 / abits;  // floor: (budget * inv_abits) >> 16 never exceeds budget / abits
  int budget = kBudget;
  bool pair_ok = true;  // every alpha within [0, 0.5]
  for (int n = 0; n < d.nA; ++n) pair_ok = pair_ok && (s_alpha[n] >= 0.0) && (s_alpha[n] <= 0.5);

  // ---- general chunk: kTJ x NCOL running products over the chunk's SNPs through T
  auto main_general = [&](int cnt, int stage, int tb) {
    const unsigned char* st = raw + stage * Cfg::STAGE_BYTES;
    const double* T = Tbuf + tb * (Cfg::T_BYTES / 8) + kidx * Cfg::TROW;
    for (int s = 0; s < cnt; ++s) {
      const double2* gp2 = reinterpret_cast<const double2*>(st + s * Cfg::SLOT_BYTES) + jg * (kTJ * 3 / 2);
      double gv[kTJ * 3];
#pragma unroll
      for (int i = 0; i < kTJ * 3 / 2; ++i) {
        const double2 v = gp2[i];
        gv[2 * i] = v.x;
        gv[2 * i + 1] = v.y;
      }
      const double2* t2 = reinterpret_cast<const double2*>(T + s * 32 * Cfg::TROW);
#pragma unroll
      for (int c2 = 0; c2 < Cfg::NCOL / 2; ++c2) {
        const double2 u0 = t2[3 * c2], u1 = t2[3 * c2 + 1], u2 = t2[3 * c2 + 2];
#pragma unroll
        for (int jj = 0; jj < kTJ; ++jj) {
          const double da = fma(gv[jj * 3 + 2], u1.x, fma(gv[jj * 3 + 1], u0.y, gv[jj * 3] * u0.x));
          const double db = fma(gv[jj * 3 + 2], u2.y, fma(gv[jj * 3 + 1], u2.x, gv[jj * 3] * u1.y));
          acc[jj][2 * c2] *= da;
          acc[jj][2 * c2 + 1] *= db;
        }
      }
    }
  };

  // ---- affine chunk: acc[j][k][n] *= f_j + alpha_n (f_k - f_j).
  // Two SNPs are taken at a time: (a1 + alpha d1)(a2 + alpha d2) = c0 + alpha (c1 + alpha c2) costs 6 FP64
  // instructions per (j,k) for the coefficients plus 2 DFMA + 1 DMUL per alpha, i.e. 1.8 instead of 2.1 per
  // likelihood and SNP.  With every alpha in [0, 0.5] each factor is >= half its f_j, so the expanded form has
  // no cancellation to speak of (terms / value <= ~4); grids reaching beyond 0.5 use the one-SNP form.
  auto main_affine = [&](int cnt, int stage, int n0) {
    const double* fj_base = reinterpret_cast<const double*>(raw + stage * Cfg::STAGE_BYTES) + jg * kTJ;
    // the thread's TK consecutive samples k sit inside one 32-sample part
    const double* fk_base = reinterpret_cast<const double*>(raw + stage * Cfg::STAGE_BYTES + Cfg::PART_BYTES) +
                            ((kidx * TK) >> 5) * (Cfg::PART_BYTES / 8) + ((kidx * TK) & 31);
    // The alphas must not sit in ordinary registers: hoisted out of the loop they cost 2*AC of the 112 and ptxas
    // then funnels every product through one temporary.  The grid travels by value in the kernel parameters
    // (d.alpha_c, constant bank) and n0 is warp-uniform, so the compiler keeps the chunk's alphas in UNIFORM
    // registers and feeds them to the FP64 instructions as uniform operands: no vector registers, no loads in
    // the loop, ten independent products in flight.
    auto load_fk = [&](int s, double* fk) {
      if constexpr (TK == 1) {
        fk[0] = fk_base[s * 32];
      } else {
#pragma unroll
        for (int kk = 0; kk < TK; kk += 2) {
          const double2 v = *reinterpret_cast<const double2*>(fk_base + s * 32 + kk);
          fk[kk] = v.x;
          fk[kk + 1] = v.y;
        }
      }
    };
    int s = 0;
    while (s < cnt) {
      if (budget < 2 * abits) {
        renorm_all();
        budget = kBudget;
      }
      const int run = min(cnt - s, (budget * inv_abits) >> 16);  // SNPs that fit into the budget (>= 1 here)
      budget -= run * abits;
      const int s_end = s + run;
      if (pair_ok) {
        for (; s + 1 < s_end; s += 2) {
          const double2 fa = *reinterpret_cast<const double2*>(fj_base + s * 32);
          const double2 fb = *reinterpret_cast<const double2*>(fj_base + (s + 1) * 32);
          double ka[TK], kb[TK];
          load_fk(s, ka);
          load_fk(s + 1, kb);
          const double c00 = fa.x * fb.x, c01 = fa.y * fb.y;
#pragma unroll
          for (int kk = 0; kk < TK; ++kk) {
            const double da0 = ka[kk] - fa.x, da1 = ka[kk] - fa.y, db0 = kb[kk] - fb.x, db1 = kb[kk] - fb.y;
            const double c20 = da0 * db0, c21 = da1 * db1;
            const double c10 = fma(fa.x, db0, fb.x * da0), c11 = fma(fa.y, db1, fb.y * da1);
#pragma unroll
            for (int a = 0; a < AC; ++a) {
              const double al = d.alpha_c[n0 + a];
              acc[0][kk * AC + a] *= fma(al, fma(al, c20, c10), c00);
              acc[1][kk * AC + a] *= fma(al, fma(al, c21, c11), c01);
            }
          }
        }
      }
#pragma unroll 1
      for (; s < s_end; ++s) {
        const double2 fj = *reinterpret_cast<const double2*>(fj_base + s * 32);
        double fk[TK];
        load_fk(s, fk);
#pragma unroll
        for (int kk = 0; kk < TK; ++kk) {
          const double d0 = fk[kk] - fj.x, d1 = fk[kk] - fj.y;
#pragma unroll
          for (int a = 0; a < AC; ++a) {
            const double al = d.alpha_c[n0 + a];
            acc[0][kk * AC + a] *= fma(al, d0, fj.x);
            acc[1][kk * AC + a] *= fma(al, d1, fj.y);
          }
        }
      }
    }
  };

  // ---- item epilogue: this warp's products -> log-sum-exp partial + top-2 candidates.  Each
  // consumer warp writes its own record (K3 merges them), so no cross-warp synchronisation.
  // Cost matters (it is paid per likelihood, not per SNP): validity is folded into the per-(j,k)
  // multiplier, and the warp-wide largest exponent is found first so that the log-sum-exp and the
  // top-2 scans touch only the handful of products that can matter; whole-warp votes skip the rest.
  double pjk[kTJ][TK];  // prod s_j * prod s_k of the affine entries, or 0 where (j,k) is not a doublet pair
  int ejk[kTJ][TK];
  auto load_item_consts = [&](int cell32, int j0, int k0) {  // issued before the item's last chunk
    const int64_t cell = cell32;
    double pmj[kTJ], pmk[TK];
    int pej[kTJ], pek[TK];
#pragma unroll
    for (int jj = 0; jj < kTJ; ++jj) {
      const int j = j0 + jg * kTJ + jj;
      pmj[jj] = d.pm[cell * d.nvp + j];
      pej[jj] = d.pe[cell * d.nvp + j];
    }
#pragma unroll
    for (int kk = 0; kk < TK; ++kk) {
      const int k = min(k0 + kidx * TK + kk, d.nvp - 1);
      pmk[kk] = d.pm[cell * d.nvp + k];
      pek[kk] = d.pe[cell * d.nvp + k];
    }
#pragma unroll
    for (int jj = 0; jj < kTJ; ++jj)
#pragma unroll
      for (int kk = 0; kk < TK; ++kk) {
        pjk[jj][kk] = pmj[jj] * pmk[kk];  // in [1,4) or 0
        ejk[jj][kk] = pej[jj] + pek[kk];
      }
  };

  auto epilogue = [&](int cell32, int j0, int k0, int n0, int64_t item) {
    const int64_t cell = cell32;
    constexpr uint32_t kFull = 0xffffffffu;
    if constexpr (TK > 2) load_item_consts(cell32, j0, k0);
    if ((dbg != nullptr) && cell >= dbg_c0 && cell < dbg_c1) {  // inspection path (ccu_demux_get_llk)
#pragma unroll
      for (int jj = 0; jj < kTJ; ++jj) {
        const int j = j0 + jg * kTJ + jj;
#pragma unroll
        for (int cc = 0; cc < Cfg::NCOL; ++cc) {
          const int k = k0 + kidx * TK + cc / AC, n = n0 + cc % AC;
          double m = acc[jj][cc] * pjk[jj][cc / AC];
          int e = aex[jj][cc] + ejk[jj][cc / AC];
          renorm(m, e);
          if (j < d.nv && k < d.nv && n < d.nA)
            dbg[(((cell - dbg_c0) * d.nv + j) * d.nv + k) * d.nA + n] = make_double2(m, (double)e);
        }
      }
    }
    // (j,k) that are not doublet pairs contribute nothing: zero multiplier; alpha == 0.5 weights (:812-817:
    // counted once, k < j, with twice the prior)
    double whalf[kTJ][TK];
#pragma unroll
    for (int jj = 0; jj < kTJ; ++jj) {
      const int j = j0 + jg * kTJ + jj;
#pragma unroll
      for (int kk = 0; kk < TK; ++kk) {
        const int k = k0 + kidx * TK + kk;
        if (j >= d.nv || k >= d.nv || j == k) pjk[jj][kk] = 0.0;
        whalf[jj][kk] = (k < j) ? 2.0 : 0.0;
      }
    }
    const int nal = d.nA - n0;  // valid alphas of this chunk (>= AC unless the grid was padded)
    // pass 1: final (mantissa, exponent) of every product and the largest exponent (branch-free)
    int emax = kLseEmpty, nlive = 0;
#pragma unroll
    for (int jj = 0; jj < kTJ; ++jj) {
#pragma unroll
      for (int cc = 0; cc < Cfg::NCOL; ++cc) {
        double m = acc[jj][cc] * pjk[jj][cc / AC];
        int e = aex[jj][cc] + ejk[jj][cc / AC];
        renorm(m, e);
        if ((cc % AC) >= nal) m = 0.0;
        const bool live = m > 0.0;
        e = live ? e : kLseEmpty;
        nlive += live ? 1 : 0;
        acc[jj][cc] = m;
        aex[jj][cc] = e;
        emax = max(emax, e);
      }
    }
    auto pos_of = [&](int t) {
      const int jj = t / Cfg::NCOL, cc = t - jj * Cfg::NCOL;
      return ((j0 + jg * kTJ + jj) * d.nv + (k0 + kidx * TK + cc / AC)) * d.nA + n0 + cc % AC;
    };
    double la;
    int le;
    Cand b1, b2;
    // pass 2, fast form.  With the WARP's largest exponent E known first (one REDUX), every product becomes the plain
    // double x = m * 2^(e - E) by an integer add on its high word (0 below 2^-1022 of the largest: nothing in the
    // sum), and order by x is order by (e, m).  So the sum is a chain of DFMAs reduced by plain adds, the thread's
    // top-2 a chain of compares on doubles, and the warp's top-2 two arg-max rounds of three REDUX each (high word,
    // low word, smallest position among equals) instead of five shuffle rounds over two (m, e, pos) candidates:
    // 2.15 -> ~1.2 ms of C2's 20.1 ms went to this epilogue, most of it latency of the old chains.
    // Exactness: x is exact for every product it does not flush, ties are broken as before (scan order inside a
    // thread, position across threads); when the warp shows fewer than two visible
    // candidates although it holds two live products or more, the exact (m, e) form below runs instead (a droplet
    // whose best hypothesis leads every other one of the warp by more than 708 nats).
    const int Emax = __reduce_max_sync(kFull, emax);
    bool need_exact;
    {
      double s0 = 0.0, s1 = 0.0, v1 = 0.0, v2 = 0.0;
      int t1 = 0, t2 = 0;
#pragma unroll
      for (int jj = 0; jj < kTJ; ++jj) {
#pragma unroll
        for (int cc = 0; cc < Cfg::NCOL; ++cc) {
          const double m = acc[jj][cc];
          const int e = aex[jj][cc];
          const int dd = e - Emax;  // <= 0; empty products carry e = kLseEmpty and m = 0
          const double x = (dd >= -1022) ? __hiloint2double(__double2hiint(m) + dd * (1 << 20), __double2loint(m)) : 0.0;
          if ((half_mask >> (n0 + cc % AC)) & 1u) {  // warp-uniform: a predicated DFMA beside a predicated DADD
            if (cc & 1) s1 = fma(whalf[jj][cc / AC], x, s1);
            else s0 = fma(whalf[jj][cc / AC], x, s0);
          } else {
            if (cc & 1) s1 += x;
            else s0 += x;
          }
          const int t = jj * Cfg::NCOL + cc;
          const bool g1 = x > v1;
          const double mn = g1 ? v1 : x;
          const int tn = g1 ? t1 : t;
          v1 = g1 ? x : v1;
          t1 = g1 ? t : t1;
          const bool g2 = mn > v2;
          v2 = g2 ? mn : v2;
          t2 = g2 ? tn : t2;
        }
      }
      la = s0 + s1;
#pragma unroll
      for (int dd = 16; dd >= 1; dd >>= 1) la += __shfl_xor_sync(kFull, la, dd);
      le = (la > 0.0) ? Emax : kLseEmpty;
      // warp arg-max, twice: the winner hands its second value up
      auto warp_pick = [&](Cand& out) {
        const uint32_t hi = (uint32_t)__double2hiint(v1), lo = (uint32_t)__double2loint(v1);
        const uint32_t H = __reduce_max_sync(kFull, hi);
        const uint32_t L = __reduce_max_sync(kFull, (hi == H) ? lo : 0u);
        const bool own = (hi == H) && (lo == L) && ((H | L) != 0u);
        const uint32_t mypos = own ? (uint32_t)pos_of(t1) : 0xffffffffu;
        const uint32_t P = __reduce_min_sync(kFull, mypos);
        if (own && mypos == P) {  // at most one lane: positions are unique
          v1 = v2;
          t1 = t2;
          v2 = 0.0;
        }
        if ((H | L) != 0u) {
          out.m = __hiloint2double((int)((H & 0x000fffffu) | 0x3ff00000u), (int)L);
          out.e = Emax + (int)(H >> 20) - 1023;
          out.pos = (int)P;
        } else {
          out = cand_empty();
        }
      };
      warp_pick(b1);
      warp_pick(b2);
      // fewer than two visible candidates although the warp holds two live products or more: some were flushed
      need_exact = false;
      if (b2.pos < 0) need_exact = __reduce_add_sync(kFull, nlive) >= 2;
      need_exact = need_exact || (force_exact != 0);  // CCU_DEMUX_EXACT_EPILOGUE=1: tests drive the exact form
#ifdef CCU_EPI_EXACT  // A/B builds: always the exact form
      need_exact = true;
#endif
    }
    if (need_exact) {
      // pass 2, exact form: sum of w * m * 2^(e - emax) and the thread's top-2 (scan order = increasing pos); branch-free,
      // two interleaved partial sums
      double la0 = 0.0, la1 = 0.0;
      double bm1 = 0.0, bm2 = 0.0;
      int be1 = INT_MIN, be2 = INT_MIN, bt1 = -1, bt2 = -1;  // t = jj*NCOL + cc
#pragma unroll
      for (int jj = 0; jj < kTJ; ++jj) {
#pragma unroll
        for (int cc = 0; cc < Cfg::NCOL; ++cc) {
          const double m = acc[jj][cc];
          const int e = aex[jj][cc];
          const double w = ((half_mask >> (n0 + cc % AC)) & 1u) ? whalf[jj][cc / AC] : 1.0;
          if (cc & 1) la1 = fma(w * m, pow2_nonpos(e - emax), la1);
          else la0 = fma(w * m, pow2_nonpos(e - emax), la0);
          const bool gt1 = (e > be1) || (e == be1 && m > bm1);  // empty products carry e = kLseEmpty > INT_MIN
          const bool gt2 = (e > be2) || (e == be2 && m > bm2);
          bm2 = gt1 ? bm1 : (gt2 ? m : bm2);
          be2 = gt1 ? be1 : (gt2 ? e : be2);
          bt2 = gt1 ? bt1 : (gt2 ? (jj * Cfg::NCOL + cc) : bt2);
          bm1 = gt1 ? m : bm1;
          be1 = gt1 ? e : be1;
          bt1 = gt1 ? (jj * Cfg::NCOL + cc) : bt1;
        }
      }
      la = la0 + la1;
      le = (la > 0.0) ? emax : kLseEmpty;
      b1.m = bm1; b1.e = (bm1 > 0.0) ? be1 : INT_MIN; b1.pos = (bm1 > 0.0) ? pos_of(bt1) : -1;
      b2.m = bm2; b2.e = (bm2 > 0.0) ? be2 : INT_MIN; b2.pos = (bm2 > 0.0) ? pos_of(bt2) : -1;
#pragma unroll
      for (int dd = 16; dd >= 1; dd >>= 1) {
        const double oa = __shfl_xor_sync(kFull, la, dd);
        const int oe = __shfl_xor_sync(kFull, le, dd);
        lse_add(la, le, oa, oe);
        const Cand o1 = cand_shfl_xor(b1, dd), o2 = cand_shfl_xor(b2, dd);
        cand_insert(b1, b2, o1);
        cand_insert(b1, b2, o2);
      }
    }
    if (lane == 0) {
      ItemPartial p;
      p.lse_a = la;
      p.lse_e = le;
      p._pad = 0;
      p.best = b1;
      p.next = b2;
      d.partials[item * kScoreWarps + warp] = p;
    }
    reset_acc();
    budget = kBudget;
  };

  // ---------------- chunk stream: every warp walks the ring on its own; the only block-level
  // synchronisation is the one barrier between writing and reading T in a general chunk ----------------
  reset_acc();
  uint32_t gt = 0;  // general chunks seen so far (T double buffer)
  for (uint32_t g = 0;; ++g) {
    const int stage = g % kNS;
    if constexpr (kLag > 0) {
      // group 1 starts chunk g once the two warps of group 0 on its sub-partition have finished chunk g + kLag - 1
      // (or the stream).  Pacing only -- nothing read below depends on it.
      if (cgrp == 1) {
        const int wa = warp_u & 3;
        while (min(prog[wa], prog[wa + 4]) < (int)g + kLag) __nanosleep(40);
      }
    }
    mbar_wait(smem_u32(&full[stage]), (g / kNS) & 1);
    const ChunkMeta* mt = &meta[stage];
    const int cnt = mt->cnt, flags = mt->flags;
    if (!(flags & 1)) {
      if constexpr (kLag > 0) {
        if (cgrp == 0 && lane == 0) prog[warp] = INT_MAX;
      }
      break;
    }
    const int m_cell = mt->cell, m_j0 = mt->j0, m_k0 = mt->k0, m_n0 = mt->n0;
    const int64_t m_item = mt->item;
    // the item's global constants are fetched before its last chunk so the loads fly while it is multiplied in
    // (only where the 2*TK extra registers are affordable)
    if constexpr (TK <= 2) {
      if (flags & 2) load_item_consts(m_cell, m_j0, m_k0);
    }
    if (flags & 4) {
      main_affine(cnt, stage, m_n0);
    } else {
      if (budget < kGenBits * cnt) {
        renorm_all();
        budget = kBudget;
      }
      budget -= kGenBits * cnt;
      compute_T(cnt, stage, gt & 1);
      if constexpr (kLag > 0) group_bar(cgrp);
      else consumer_bar();
      main_general(cnt, stage, gt & 1);
      ++gt;
    }
    if (flags & 2) epilogue(m_cell, m_j0, m_k0, m_n0, m_item);
    __syncwarp();
    if (lane == 0) {
      mbar_arrive(smem_u32(&empty[stage]));  // this warp is done with ring stage `stage`
      if constexpr (kLag > 0) {
        if (cgrp == 0) prog[warp] = (int)g + 1;
      }
    }
  }
}

// dbg[i] = (mantissa, exponent) -> llk = log(m) + e*ln2, in place (first double of each pair)
__global__ void dbg_convert_kernel(double2* __restrict__ dbg, double* __restrict__ llk, int64_t n) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const double2 v = dbg[i];
  double r;
  if (v.x != v.x) r = v.x;  // NaN marker: entry not computed by the engine
  else r = (v.x > 0.0) ? (log(v.x) + v.y * 0.6931471805599453094) : -CUDART_INF;
  llk[i] = r;
}

cudaError_t launch_dbg_convert(double2* dbg, double* llk, int64_t n, cudaStream_t st) {
  if (n <= 0) return cudaSuccess;
  dbg_convert_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(dbg, llk, n);
  return cudaGetLastError();
}

ScorePlan make_score_plan(int nv, int nA, const double* alpha) {
  ScorePlan p;
  for (int n = 0; n < CCU_MAX_ALPHA; ++n) p.is_half[n] = (n < nA && alpha[n] == 0.5) ? 1 : 0;
  const int nd = nA - 1;
  // alphas per chunk: minimise padded work, prefer larger chunks on ties
  const int cand[4] = {10, 5, 2, 1};
  int best = 1, best_cost = 1 << 30;
  for (int i = 0; i < 4; ++i) {
    int ac = cand[i];
    int cost = ((nd + ac - 1) / ac) * ac;
    if (cost < best_cost) {
      best_cost = cost;
      best = ac;
    }
  }
  p.ac = best;
  p.tk = (best == 10) ? 1 : (best == 5) ? 2 : (best == 2) ? 4 : 8;
  p.kt = 32 * p.tk;
  const int nvp = (nv + 31) / 32 * 32;
  p.njt = nvp / kJT;
  p.nkt = (nvp + p.kt - 1) / p.kt;
  p.nac = (nd + p.ac - 1) / p.ac;
  p.items_per_cell = p.njt * p.nkt * p.nac;
  switch (p.ac) {
    case 10: p.smem_bytes = ScoreCfg<10, 1>::SMEM_BYTES; break;
    case 5: p.smem_bytes = ScoreCfg<5, 2>::SMEM_BYTES; break;
    case 2: p.smem_bytes = ScoreCfg<2, 4>::SMEM_BYTES; break;
    default: p.smem_bytes = ScoreCfg<1, 8>::SMEM_BYTES; break;
  }
  return p;
}

template <int AC, int TK, int LAGP>
static cudaError_t launch_score_t(const DemuxDev& d, const ScorePlan& plan, int num_sms, double2* llk_dbg,
                                  int64_t c0, int64_t c1, int* grid_out, cudaStream_t st) {
  using Cfg = ScoreCfg<AC, TK, LAGP>;
  cudaError_t err = cudaFuncSetAttribute(demux_score_kernel<AC, TK, LAGP>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES);
  if (err != cudaSuccess) return err;
  const int64_t nitems = d.nbcs * plan.items_per_cell;
  int grid = (int)((nitems < num_sms) ? nitems : num_sms);
  if (grid_out) *grid_out = grid;
  uint32_t half_mask = 0;
  for (int n = 0; n < d.nA && n < 32; ++n)
    if (plan.is_half[n]) half_mask |= 1u << n;
  const char* fe = getenv("CCU_DEMUX_EXACT_EPILOGUE");
  const int force_exact = (fe && fe[0] == '1') ? 1 : 0;
  // work counter of the launch: the 8 bytes behind the tile kernel's (d.maxbq: [0] deepest quality, [8] tile batches)
  unsigned long long* next_item = reinterpret_cast<unsigned long long*>(d.maxbq) + 2;
  err = cudaMemsetAsync(next_item, 0, sizeof(unsigned long long), st);
  if (err != cudaSuccess) return err;
  demux_score_kernel<AC, TK, LAGP><<<grid, kScoreBlock, Cfg::SMEM_BYTES, st>>>(d, plan.njt, plan.nkt, plan.nac, nitems,
                                                                              half_mask, llk_dbg, c0, c1, next_item,
                                                                              force_exact);
  return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------
// K2 for small pools (nv <= 32 and nv*nv*(nA-1) <= 1024 products per droplet -- the classic demuxlet run: 8 or 16
// pooled samples, a few doublet alphas).  The tiled kernel above pads the samples to 32 x 32 per alpha chunk and
// pays its per-item machinery (ring, epilogue of 20 products per thread) for a handful of live products: at 8
// samples and alpha {0, 0.5} it ran at 0.3 % of its own arithmetic.  Here wpc warps (1, 2, 4 or 8) own a droplet
// and a lane owns every (32*wpc)-th product of the list p = (j*nv + k)*(nA-1) + (n-1), i.e. the reference's
// scan order.  Per entry the lane needs f_j, f_k (single-read entries: one 32-double row of `fmat`, staged in shared
// memory) or the genotype rows and the alpha block of the tile (entries with >= 2 informative reads, read
// straight from global memory); same running products, same epilogue rules and the same partial record
// (one per warp, wpc per droplet) as the tiled kernel, so K3 is shared.
// ------------------------------------------------------------------------------------------
constexpr int kSmallWarps = 8;

template <int PPL>
__global__ void __launch_bounds__(kSmallWarps * 32)
    demux_score_small_kernel(DemuxDev d, int wpc, uint32_t half_mask, double2* __restrict__ dbg, int64_t dbg_c0,
                             int64_t dbg_c1) {
  __shared__ double s_f[kSmallWarps][2][32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int wsub = warp % wpc;  // this warp's share of the droplet's products
  const int64_t c = (int64_t)blockIdx.x * (kSmallWarps / wpc) + warp / wpc;
  if (c >= d.nbcs) return;
  const int nv = d.nv, nd = d.nA - 1, P = nv * nv * nd;
  constexpr uint32_t kFull = 0xffffffffu;

  int pj[PPL], pk[PPL], pn[PPL];  // sample j, sample k, alpha index n >= 1 of each product (pn = 0: no product)
  double acc[PPL];
  int aex[PPL];
#pragma unroll
  for (int i = 0; i < PPL; ++i) {
    const int p = wsub * 32 + lane + 32 * wpc * i;
    const bool ok = p < P;
    const int jk = ok ? p / nd : 0;
    pj[i] = jk / nv;
    pk[i] = jk - pj[i] * nv;
    pn[i] = ok ? 1 + (p - jk * nd) : 0;
    acc[i] = 1.0;
    aex[i] = 0;
  }

  const int64_t base = d.cell_ptr[c];
  const int ng = d.vcnt_g[c], na = d.vcnt_a[c];
  const int64_t grow = (int64_t)d.nvp * 3;
  // ---- entries with >= 2 informative reads: sum_l g_j[l] * (sum_m g_k[m] * pG[n][l][m])
  for (int i = 0; i < ng; ++i) {
    const int64_t e = d.vent[base + i];
    const double* g = d.G + (int64_t)d.vsnp[base + i] * grow;
    const double* tile = d.tiles + (int64_t)d.tslot[e] * d.nA * kTileStride;
#pragma unroll
    for (int q = 0; q < PPL; ++q)
      if (pn[q] > 0) {
        const double* gj = g + 3 * pj[q];
        const double* gk = g + 3 * pk[q];
        const double* pg = tile + pn[q] * kTileStride;
        const double k0 = gk[0], k1 = gk[1], k2 = gk[2];
        const double t0 = fma(k2, pg[2], fma(k1, pg[1], k0 * pg[0]));
        const double t1 = fma(k2, pg[5], fma(k1, pg[4], k0 * pg[3]));
        const double t2 = fma(k2, pg[8], fma(k1, pg[7], k0 * pg[6]));
        acc[q] *= fma(gj[2], t2, fma(gj[1], t1, gj[0] * t0));
      }
    if ((i & 15) == 15) {  // a general factor is >= ~2^-34: 16 of them cannot underflow
#pragma unroll
      for (int q = 0; q < PPL; ++q) renorm(acc[q], aex[q]);
    }
  }
#pragma unroll
  for (int q = 0; q < PPL; ++q) renorm(acc[q], aex[q]);
  // ---- single-read entries: f_j + alpha_n (f_k - f_j); the s_j s_k factors come in at the end
  const double* frow = d.fmat + base * d.nvp + lane;  // nvp == 32 here: one 32-sample tile, [entry][32]
  double fnext = (na > 0) ? frow[0] : 0.0;
  for (int i = 0; i < na; ++i) {
    double* S = s_f[warp][i & 1];
    S[lane] = fnext;
    __syncwarp();
    if (i + 1 < na) fnext = frow[(int64_t)(i + 1) * 32];
#pragma unroll
    for (int q = 0; q < PPL; ++q)
      if (pn[q] > 0) {
        const double fj = S[pj[q]], fk = S[pk[q]];
        acc[q] *= fma(d.alpha_c[pn[q]], fk - fj, fj);
      }
    if ((i & 15) == 15) {  // a single-read factor is >= err/3 >= ~2^-34 as well
#pragma unroll
      for (int q = 0; q < PPL; ++q) renorm(acc[q], aex[q]);
    }
  }

  // ---- epilogue: the rules of the tiled kernel's item epilogue, for this droplet's whole product list
  const bool want_dbg = (dbg != nullptr) && c >= dbg_c0 && c < dbg_c1;
  double la = 0.0;
  int le = kLseEmpty;
  Cand b1 = cand_empty(), b2 = cand_empty();
#pragma unroll
  for (int q = 0; q < PPL; ++q)
    if (pn[q] > 0) {
      const int j = pj[q], k = pk[q], n = pn[q];
      double m = acc[q] * (d.pm[c * d.nvp + j] * d.pm[c * d.nvp + k]);
      int e = aex[q] + d.pe[c * d.nvp + j] + d.pe[c * d.nvp + k];
      renorm(m, e);
      if (want_dbg) dbg[(((c - dbg_c0) * nv + j) * nv + k) * d.nA + n] = make_double2(m, (double)e);
      if (j != k && m > 0.0) {
        // alpha == 0.5 is counted once, k < j, with twice the prior (:812-817)
        const double w = ((half_mask >> n) & 1u) ? ((k < j) ? 2.0 : 0.0) : 1.0;
        lse_add(la, le, w * m, e);
        Cand x;
        x.m = m;
        x.e = e;
        x.pos = (j * nv + k) * d.nA + n;
        cand_insert(b1, b2, x);
      }
    }
#pragma unroll
  for (int dd = 16; dd >= 1; dd >>= 1) {
    const double oa = __shfl_xor_sync(kFull, la, dd);
    const int oe = __shfl_xor_sync(kFull, le, dd);
    lse_add(la, le, oa, oe);
    const Cand o1 = cand_shfl_xor(b1, dd), o2 = cand_shfl_xor(b2, dd);
    cand_insert(b1, b2, o1);
    cand_insert(b1, b2, o2);
  }
  if (lane == 0) {
    ItemPartial p;
    p.lse_a = la;
    p.lse_e = (la > 0.0) ? le : kLseEmpty;
    p._pad = 0;
    p.best = b1;
    p.next = b2;
    d.partials[c * wpc + wsub] = p;
  }
}

// products per droplet the small-pool kernel takes (0: use the tiled kernel)
static int small_pool_ppl(const DemuxDev& d, int* wpc_out) {
  if (d.nvp != 32) return 0;
  const int P = d.nv * d.nv * (d.nA - 1);
  if (P <= 0 || P > 1024) return 0;  // beyond that the tiled kernel wins (it needs 1.8 instructions per product and
                                     // SNP where this one needs ~8)
  int wpc = 1;
  while (wpc < kSmallWarps && P > 256 * wpc) wpc *= 2;  // at most 8 products per lane
  int ppl = 2;
  while (ppl * 32 * wpc < P) ppl *= 2;
  if (wpc_out) *wpc_out = wpc;
  return ppl;
}

int score_partials_per_cell(const DemuxDev& d, const ScorePlan& plan) {
  int wpc = 1;
  return small_pool_ppl(d, &wpc) ? wpc : plan.items_per_cell * kScoreWarps;
}

cudaError_t launch_score(const DemuxDev& d, const ScorePlan& plan, int num_sms, double2* llk_dbg,
                         int64_t dbg_c0, int64_t dbg_c1, int* grid_out, cudaStream_t st) {
  if (d.nbcs == 0) return cudaSuccess;
  int wpc = 1;
  if (const int ppl = small_pool_ppl(d, &wpc)) {
    uint32_t half_mask = 0;
    for (int n = 0; n < d.nA && n < 32; ++n)
      if (plan.is_half[n]) half_mask |= 1u << n;
    const int cpb = kSmallWarps / wpc;  // droplets per CTA
    const unsigned blocks = (unsigned)((d.nbcs + cpb - 1) / cpb);
    if (grid_out) *grid_out = (int)blocks;
#define CCU_SMALL(N) \
  demux_score_small_kernel<N><<<blocks, kSmallWarps * 32, 0, st>>>(d, wpc, half_mask, llk_dbg, dbg_c0, dbg_c1)
    switch (ppl) {
      case 2: CCU_SMALL(2); break;
      case 4: CCU_SMALL(4); break;
      case 8: CCU_SMALL(8); break;
      default: CCU_SMALL(16); break;
    }
#undef CCU_SMALL
    return cudaGetLastError();
  }
  switch (plan.ac) {
    case 10: {
      // out-of-step consumer groups where general chunks carry weight (see CCU_LAG_MODE above)
      const bool lag = (CCU_LAG_MODE >= 0) ? (CCU_LAG_MODE > 0) : (d.ngen * 10 > d.nnz);
      if (lag) return launch_score_t<10, 1, 1>(d, plan, num_sms, llk_dbg, dbg_c0, dbg_c1, grid_out, st);
      return launch_score_t<10, 1, 0>(d, plan, num_sms, llk_dbg, dbg_c0, dbg_c1, grid_out, st);
    }
    case 5: return launch_score_t<5, 2, 0>(d, plan, num_sms, llk_dbg, dbg_c0, dbg_c1, grid_out, st);
    case 2: return launch_score_t<2, 4, 0>(d, plan, num_sms, llk_dbg, dbg_c0, dbg_c1, grid_out, st);
    default: return launch_score_t<1, 8, 0>(d, plan, num_sms, llk_dbg, dbg_c0, dbg_c1, grid_out, st);
  }
}

// ------------------------------------------------------------------------------------------
// K3: per-droplet merge (cmd_cram_demuxlet.cpp:788-906)
// ------------------------------------------------------------------------------------------
struct SngCand {
  double v;
  int32_t j;
};
__device__ __forceinline__ bool sng_better(const SngCand& a, const SngCand& b) {
  if (a.v != b.v) return a.v > b.v;
  return (uint32_t)a.j < (uint32_t)b.j;
}
__device__ __forceinline__ void sng_insert(SngCand& b1, SngCand& b2, const SngCand& x) {
  if (sng_better(x, b1)) {
    b2 = b1;
    b1 = x;
  } else if (sng_better(x, b2)) {
    b2 = x;
  }
}

__global__ void __launch_bounds__(256)
    demux_finalize_kernel(DemuxDev d, int nrec, double log_single_prior, double log_doublet_prior1) {
  const int lane = threadIdx.x & 31;
  const int64_t c = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (c >= d.nbcs) return;
  const double LN2 = 0.6931471805599453094;

  // ---- singlets (:804-807, :827-837)
  SngCand s1 = {-1e300, -1}, s2 = {-1e300, -1};
  double smax = 0.0;  // the -1e-300 seed of :791 is exp(.) == 1, i.e. a term at 0
  for (int j = lane; j < d.nv; j += 32) {
    const double v = d.sng_llk[c * d.nv + j];
    SngCand x = {v, j};
    if (v > -1e300) sng_insert(s1, s2, x);
    smax = fmax(smax, v + log_single_prior);
  }
#pragma unroll
  for (int dd = 16; dd >= 1; dd >>= 1) {
    SngCand o1, o2;
    o1.v = __shfl_xor_sync(0xffffffffu, s1.v, dd);
    o1.j = __shfl_xor_sync(0xffffffffu, s1.j, dd);
    o2.v = __shfl_xor_sync(0xffffffffu, s2.v, dd);
    o2.j = __shfl_xor_sync(0xffffffffu, s2.j, dd);
    if (o1.j >= 0) sng_insert(s1, s2, o1);
    if (o2.j >= 0) sng_insert(s1, s2, o2);
    smax = fmax(smax, __shfl_xor_sync(0xffffffffu, smax, dd));
  }

  // ---- doublet partials of this droplet's items
  double la = 0.0;
  int le = kLseEmpty;
  Cand b1 = cand_empty(), b2 = cand_empty();
  for (int i = lane; i < nrec; i += 32) {  // nrec partial records per droplet
    const ItemPartial p = d.partials[c * (int64_t)nrec + i];
    lse_add(la, le, p.lse_a, p.lse_e);
    cand_insert(b1, b2, p.best);
    cand_insert(b1, b2, p.next);
  }
#pragma unroll
  for (int dd = 16; dd >= 1; dd >>= 1) {
    const double oa = __shfl_xor_sync(0xffffffffu, la, dd);
    const int oe = __shfl_xor_sync(0xffffffffu, le, dd);
    lse_add(la, le, oa, oe);
    const Cand o1 = cand_shfl_xor(b1, dd), o2 = cand_shfl_xor(b2, dd);
    cand_insert(b1, b2, o1);
    cand_insert(b1, b2, o2);
  }
  const double td = (la > 0.0) ? (log_doublet_prior1 + log(la) + (double)le * LN2) : -CUDART_INF;
  // At alpha == 0.5 the hypotheses (j,k) and (k,j) are the same model; the reference computes both and
  // lets last-bit rounding of its own summation order decide which one is "best" and which "next"
  // (an exact tie keeps the first scanned, j < k).  That noise cannot be reproduced,
  // so the pair is reported in the canonical orientation j < k whenever best/next are such mirrors.
  if (b1.pos >= 0 && b2.pos >= 0) {
    const int nvA = d.nv * d.nA;
    const int j1 = b1.pos / nvA, k1 = (b1.pos % nvA) / d.nA, n1 = b1.pos % d.nA;
    const int j2 = b2.pos / nvA, k2 = (b2.pos % nvA) / d.nA, n2 = b2.pos % d.nA;
    if (n1 == n2 && d.is_half[n1] && j1 == k2 && k1 == j2 && j1 > k1 && b1.e == b2.e &&
        fabs(b1.m - b2.m) <= 1e-12 * b1.m) {
      const Cand t = b1;
      b1 = b2;
      b2 = t;
    }
  }

  // ---- posterior normalisers (:804-821): log(1 + sum exp(term))
  const double msng = smax;
  const double mall = fmax(msng, td);
  double ssng = 0.0, sall = 0.0;
  for (int j = lane; j < d.nv; j += 32) {
    const double t = d.sng_llk[c * d.nv + j] + log_single_prior;
    ssng += exp(t - msng);
    sall += exp(t - mall);
  }
#pragma unroll
  for (int dd = 16; dd >= 1; dd >>= 1) {
    ssng += __shfl_xor_sync(0xffffffffu, ssng, dd);
    sall += __shfl_xor_sync(0xffffffffu, sall, dd);
  }
  if (lane == 0) {
    ccu_demux_result r;
    r.sBest = s1.j;
    r.sNext = s2.j;
    r.sngBestLLK = s1.v;
    r.sngNextLLK = s2.v;
    r.sngLLK = msng + log(exp(-msng) + ssng);
    r.sumLLK = mall + log(exp(-mall) + sall + exp(td - mall));
    const int nvA = d.nv * d.nA;
    if (b1.pos >= 0) {
      r.dBest1 = b1.pos / nvA;
      r.dBest2 = (b1.pos % nvA) / d.nA;
      r.dBestA = b1.pos % d.nA;
      r.dblBestLLK = log(b1.m) + (double)b1.e * LN2;
    } else {
      r.dBest1 = r.dBest2 = r.dBestA = -1;
      r.dblBestLLK = -1e300;
    }
    if (b2.pos >= 0) {
      r.dNext1 = b2.pos / nvA;
      r.dNext2 = (b2.pos % nvA) / d.nA;
      r.dNextA = b2.pos % d.nA;
      r.dblNextLLK = log(b2.m) + (double)b2.e * LN2;
    } else {
      r.dNext1 = r.dNext2 = r.dNextA = -1;
      r.dblNextLLK = -1e300;
    }
    d.results[c] = r;
  }
}

cudaError_t launch_finalize(const DemuxDev& d, const ScorePlan& plan, cudaStream_t st) {
  if (d.nbcs == 0) return cudaSuccess;
  const double nv = d.nv, nA = d.nA;
  // cmd_cram_demuxlet.cpp:793-794
  const double lsp = log((1.0 - d.doublet_prior) / nv);
  const double ldp1 = log(d.doublet_prior / nv / (nv - 1.) / (nA - 1.));
  int64_t blocks = (d.nbcs * 32 + 255) / 256;
  demux_finalize_kernel<<<(unsigned)blocks, 256, 0, st>>>(d, score_partials_per_cell(d, plan), lsp, ldp1);
  return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------
// <out>.single support: llks00[0] of cmd_cram_demuxlet.cpp:427-440, :762-773 -- the likelihood of the droplet under
// two fully unspecified samples, sum_snp log sum_{l,m} gp0[l] gp0[m] pG[0][l][m] with gp0 the mean genotype
// distribution of the pool at the SNP.  Only computed on request (ccu_demux_get_singlets): the reference itself
// never prints it (dead code there), so it is kept out of the scoring run.
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) demux_gp0_kernel(DemuxDev d, double* __restrict__ gp0) {
  const int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= d.nsnps) return;
  double a0 = 0, a1 = 0, a2 = 0;
  const double* g = d.G + s * (int64_t)d.nvp * 3;
  for (int j = 0; j < d.nv; ++j) {  // :431-434, in sample order
    a0 = __dadd_rn(a0, g[3 * j]);
    a1 = __dadd_rn(a1, g[3 * j + 1]);
    a2 = __dadd_rn(a2, g[3 * j + 2]);
  }
  gp0[s * 4] = a0 / d.nv;
  gp0[s * 4 + 1] = a1 / d.nv;
  gp0[s * 4 + 2] = a2 / d.nv;
  gp0[s * 4 + 3] = 0;
}

__global__ void __launch_bounds__(256) demux_llk00_kernel(DemuxDev d, const double* __restrict__ gp0, double* __restrict__ out) {
  const int lane = threadIdx.x & 31;
  const int64_t c = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (c >= d.nbcs) return;
  const int64_t base = d.cell_ptr[c];
  const int cnt = d.vcnt_g[c] + d.vcnt_a[c];
  double acc = 0.0;
  for (int i = lane; i < cnt; i += 32) {
    // alpha[0] == 0: pG[0][l][m] = v(l) for every m (vaff holds v for every entry)
    const double4 v = reinterpret_cast<const double4*>(d.vaff)[d.vent[base + i]];
    const double* p = gp0 + (int64_t)d.vsnp[base + i] * 4;
    const double vl[3] = {v.x, v.y, v.z};
    double sum = 0.0;
#pragma unroll
    for (int l = 0; l < 3; ++l)
#pragma unroll
      for (int m = 0; m < 3; ++m) sum = __dadd_rn(sum, __dmul_rn(__dmul_rn(p[l], p[m]), vl[l]));  // :764-770
    acc += log(sum);
  }
#pragma unroll
  for (int dd = 16; dd >= 1; dd >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, dd);
  if (lane == 0) out[c] = acc;
}

cudaError_t launch_llk00(const DemuxDev& d, double* gp0_tmp, double* out, cudaStream_t st) {
  if (d.nsnps > 0) demux_gp0_kernel<<<(unsigned)((d.nsnps + 127) / 128), 128, 0, st>>>(d, gp0_tmp);
  if (d.nbcs > 0) demux_llk00_kernel<<<(unsigned)((d.nbcs * 32 + 255) / 256), 256, 0, st>>>(d, gp0_tmp, out);
  return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------
// FP64 roofline denominator: 8 independent DFMA chains per thread
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) fp64_peak_kernel(double* sink, int iters) {
  double a0 = 1.0 + threadIdx.x * 1e-9, a1 = a0 + 1e-3, a2 = a0 + 2e-3, a3 = a0 + 3e-3;
  double a4 = a0 + 4e-3, a5 = a0 + 5e-3, a6 = a0 + 6e-3, a7 = a0 + 7e-3;
  const double x = 0.999999, y = 1e-7;
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      a0 = fma(a0, x, y);
      a1 = fma(a1, x, y);
      a2 = fma(a2, x, y);
      a3 = fma(a3, x, y);
      a4 = fma(a4, x, y);
      a5 = fma(a5, x, y);
      a6 = fma(a6, x, y);
      a7 = fma(a7, x, y);
    }
  }
  const double s = ((a0 + a1) + (a2 + a3)) + ((a4 + a5) + (a6 + a7));
  if (s == 123.456) sink[0] = s;
}

cudaError_t launch_fp64_peak(double* sink, int iters, int num_sms, cudaStream_t st) {
  fp64_peak_kernel<<<num_sms * 4, 256, 0, st>>>(sink, iters);
  return cudaGetLastError();
}

}  // namespace ccu
