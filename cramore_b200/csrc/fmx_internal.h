// fmx_internal.h -- device-side layout and launch wrappers of the freemuxlet kernels
// (fmx_kernels.cu), shared with the C-ABI host code (ccu_fmx_api.cu).  Not part of the public ABI.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "cramore_cuda.h"

namespace ccu {

constexpr int kStatStride = 12;   // doubles per (cluster,SNP): gls[9], nreads, nref, nalt
constexpr int kFmxMaxClust = 64;  // K(K+1)/2 <= 2080 pair likelihoods per droplet

// Reduced scans of one droplet's pair likelihoods (E-step kernel -> decision kernel)
struct FmxScan {
  double sv1, sv2, dv1, dv2;      // best / next singlet and doublet log-likelihood
  int32_t si1, si2, di1, di2;     // cluster of the singlets, pair index j(j+1)/2+k of the doublets (INT_MAX: none)
  double ssum, asum, smax, vmax;  // sum exp(llk + prior - max) over singlets / all, and the two maxima
};

struct FmxDev {
  int32_t K;
  double doublet_prior, geno_error;
  double log_sng_prior, log_dbl_prior;  // log((1-prior)/K), log(prior/K/(K-1)*2)  (cmd_cram_freemuxlet.cpp:462-463)
  int64_t nbcs, nsnps, nnz;
  int64_t max_segment;       // most droplets any SNP has (selects the M-step kernel)
  // droplet store (all droplets)
  const int64_t* cell_ptr;   // [nbcs+1]
  const int32_t* snp_idx;    // [nnz]
  const int64_t* read_ptr;   // [nnz+1]
  const uint8_t* al;
  const uint8_t* bq;
  const double* af;          // [nsnps]
  const double* phred;       // [512] err, mat
  // per-entry pileups (static over the EM loop), cell-major
  double* gl;                // [nnz][9]
  int4* cnt;                 // [nnz] nreads, nref, nalt, 0
  double* logdenom;          // [nnz]
  // SNP-major copy for the M-step
  int64_t* csc_ptr;          // [nsnps+1]
  int32_t* csc_cell;         // [nnz] droplet id, ascending inside a SNP
  double* csc_gl;            // [nnz][9]
  int4* csc_cnt;             // [nnz]
  // EM state
  int32_t* assign;           // [nbcs] cluster of the droplet in the next M-step, or -1
  double* stats;             // [K][nsnps][kStatStride]
  double* cgp;               // [nsnps][K][3] per-cluster genotype posteriors of the current iteration
  double* llk;               // [cells of the shard][K(K+1)/2]
  ccu_fmx_result* results;   // [cells of the shard]
  FmxScan* scan;             // [cells of the shard]
  // shard
  int64_t cell_begin, cell_end, snp_begin, snp_end;
};

cudaError_t launch_fmx_pileup(const FmxDev& d, double alpha, cudaStream_t st);
cudaError_t launch_fmx_lmix(const FmxDev& d, double* llk0, double* llk2, cudaStream_t st);
// out[i] = i
cudaError_t launch_fmx_iota(int32_t* out, int64_t n, cudaStream_t st);
// ent_cell[nnz]: droplet id of every entry
cudaError_t launch_fmx_entry_cells(const FmxDev& d, int32_t* ent_cell, cudaStream_t st);
// after the (snp, entry) pairs were sorted by snp: gather the SNP-major copy and the segment offsets
cudaError_t launch_fmx_build_csc(const FmxDev& d, const int32_t* sorted_snp, const int32_t* sorted_ent,
                                 const int32_t* ent_cell, cudaStream_t st);
cudaError_t launch_fmx_mstep(const FmxDev& d, cudaStream_t st);
cudaError_t launch_fmx_cgp(const FmxDev& d, int apply_geno_error, bool slab_only, cudaStream_t st);
// Peer-memory exchange (one process per GPU, NVLink / NVSwitch): the posterior and assignment arrays of every rank,
// opened through CUDA IPC.  launch_fmx_cgp_push computes this rank's SNP slab of the posteriors and stores every
// value straight into all ranks' arrays (its own included) -- the all-gather happens inside the producing kernel;
// launch_fmx_assign_push copies this rank's droplet range of the assignment vector to all ranks.
constexpr int kMaxPeers = 16;
struct FmxPeers {
  double* post[kMaxPeers];
  int32_t* assign[kMaxPeers];
  uint32_t* flags[kMaxPeers];  // [kMaxPeers + 2]: arrival epochs written by the peers, own epoch, error flag
  int32_t n;
};
cudaError_t launch_fmx_cgp_push(const FmxDev& d, int apply_geno_error, const FmxPeers& peers, cudaStream_t st);
cudaError_t launch_fmx_assign_push(const FmxDev& d, const FmxPeers& peers, cudaStream_t st);
// Stream-ordered barrier across the ranks through the mapped flag arrays: every rank stores its next epoch into its
// slot of every peer's array (release, system scope) and waits until all slots of its own array have reached it.
// The epoch lives in device memory, so the launch can sit in a replayed CUDA graph.  A rank that waits longer than
// ~2 s gives up and raises the error flag (flags[kMaxPeers + 1]) instead of hanging the GPU.
cudaError_t launch_fmx_rank_barrier(const FmxPeers& peers, int rank, cudaStream_t st);
cudaError_t preload_fmx_rank_barrier();
cudaError_t launch_fmx_estep(const FmxDev& d, cudaStream_t st);
// int32 count of droplets whose (type, jBest, kBest) differ between two result arrays
// counts bitwise differences between the shared-reciprocal division of the pileup kernels and __ddiv_rn
cudaError_t launch_fmx_div_selftest(uint64_t seed, int64_t n, unsigned long long* mismatches, cudaStream_t st);
cudaError_t launch_fmx_unpack_stats(const FmxDev& d, ccu_pileup* out, cudaStream_t st);

}  // namespace ccu
