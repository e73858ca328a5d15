// ccu_handle.h -- the engine handle shared by the C-ABI translation units (ccu_api.cu: demuxlet,
// ccu_fmx_api.cu: freemuxlet).  Not part of the public ABI.
#pragma once
#include <cuda_runtime.h>
#include <nvtx3/nvToolsExt.h>

#include <cstdarg>
#include <cstdio>
#include <string>
#include <vector>

#include "cramore_cuda.h"

// NVTX range around an entry point of the C ABI: shows up in Nsight Systems / ncu --nvtx next to the kernels it
// launches; a no-op when no tool is attached (header-only NVTX v3).
struct CcuRange {
  explicit CcuRange(const char* name) { nvtxRangePushA(name); }
  ~CcuRange() { nvtxRangePop(); }
};

struct DevBuf {
  void* p = nullptr;
  size_t cap = 0;
  cudaError_t reserve(size_t bytes) {
    if (bytes <= cap) return cudaSuccess;
    if (p) cudaFree(p);
    p = nullptr;
    cap = 0;
    size_t want = bytes + bytes / 8 + 256;
    cudaError_t e = cudaMalloc(&p, want);
    if (e != cudaSuccess) {
      cudaGetLastError();
      e = cudaMalloc(&p, bytes);
      want = bytes;
    }
    if (e == cudaSuccess) cap = want;
    return e;
  }
  void release() {
    if (p) cudaFree(p);
    p = nullptr;
    cap = 0;
  }
  template <typename T>
  T* as() const {
    return reinterpret_cast<T*>(p);
  }
};

enum { EV_RUN0, EV_TILE, EV_COMPACT, EV_SNG, EV_SCORE, EV_FIN, EV_GP0, EV_GP1, EV_X0, EV_X1, EV_F0, EV_F1, EV_COUNT };

struct FmxState;  // ccu_fmx_api.cu

struct ccu_handle {
  int device = 0;
  int num_sms = 0;
  cudaStream_t own_stream = nullptr;
  cudaStream_t stream = nullptr;
  cudaEvent_t ev[EV_COUNT] = {};
  std::string err;
  ccu_timings tm = {};

  // configuration
  bool configured = false;
  int nv = 0, nvp = 0, nA = 0;
  double doublet_prior = 0.5;
  std::vector<double> alpha;
  DevBuf d_alpha, d_ptab, d_half, d_phred;

  // genotypes
  int64_t nsnps = -1;
  int64_t staged_snps = -1;   // SNPs reserved by ccu_demux_genotypes_stage and not committed yet
  bool has_gp_flags = false;
  bool have_avg = false;      // d_avg holds the per-SNP averages of the float matrix (set_genotypes / commit)
  DevBuf d_gpf, d_err, d_hasgp, d_avg, d_G;

  // droplets
  int64_t nbcs = -1, nnz = 0, nreads = 0;
  bool ran = false;
  DevBuf d_cell_ptr, d_snp_idx, d_read_ptr, d_al, d_bq;
  // second set of the five store arrays and a stream of its own: ccu_demux_score uploads the second part of a store
  // into them while the first part is being scored (ccu_api.cu)
  DevBuf d_in2[5];
  cudaStream_t aux_stream = nullptr;
  DevBuf d_tiles, d_vaff, d_eclass, d_vent, d_vsnp, d_vcnt, d_sng, d_fmat, d_pm, d_pe, d_maxbq, d_partials, d_results,
      d_dbg, d_sink, d_flag, d_tslot, d_scan, d_ngen;
  int64_t ngen = 0;  // entries with >= 2 informative reads in the uploaded store (sizes the tile buffer)
  double scratch_fits = 0.0;  // largest scratch need (bytes) that ran in one piece on this handle
  bool validated = false;  // the uploaded CSR passed demux_validate_kernel
  bool partial = false;    // the handle holds the LAST PART of a store that ccu_demux_score ran in parts: the inspection
                           // calls (tiles, llk, singlets, device results) refuse instead of answering for the wrong droplets

  // freemuxlet (owned by ccu_fmx_api.cu)
  FmxState* fmx = nullptr;
};

int ccu_fail(ccu_handle* h, const char* fmt, ...);
void ccu_fmx_release(ccu_handle* h);  // frees h->fmx (called by ccu_destroy)

#define CCU_CUDA(h, call)                                                                           \
  do {                                                                                              \
    cudaError_t e__ = (call);                                                                       \
    if (e__ != cudaSuccess)                                                                         \
      return ccu_fail(h, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e__), __FILE__, __LINE__); \
  } while (0)

inline float ccu_ev_ms(ccu_handle* h, int a, int b) {
  float ms = 0.f;
  if (cudaEventElapsedTime(&ms, h->ev[a], h->ev[b]) != cudaSuccess) {
    cudaGetLastError();
    return 0.f;
  }
  return ms;
}
