// fmx_internal.h -- device-side layout and launch wrappers of the freemuxlet kernels
// (fmx_kernels.cu), shared with the C-ABI host code (ccu_fmx_api.cu).  Not part of the public ABI.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "cramore_cuda.h"

namespace ccu {

constexpr int kStatStride = 12;   // doubles per (cluster,SNP): gls[9], nreads, nref, nalt
constexpr int kFmxMaxClust = 64;  // K(K+1)/2 <= 2080 pair likelihoods per droplet
constexpr int kFmxAffineMaxK = 16;  // up to here the E-step takes single-read entries by the affine identity (rows2 kernels)
// SNP-major rows of the M-step: 80 bytes, 16-byte aligned; the read counts of an entry take 21 bits each
// (ccu_fmx_upload refuses a (droplet, SNP) entry with 2^21 reads or more)
constexpr int kCscRow = 10;
constexpr int kCntBits = 21;
constexpr unsigned long long kCntMask = (1ull << kCntBits) - 1;
__host__ __device__ inline unsigned long long fmx_pack_counts(int nreads, int nref, int nalt) {
  return (unsigned long long)nreads | ((unsigned long long)nref << kCntBits) | ((unsigned long long)nalt << (2 * kCntBits));
}

// Reduced scans of one droplet's pair likelihoods (E-step kernel -> decision kernel)
struct FmxScan {
  double sv1, sv2, dv1, dv2;      // best / next singlet and doublet log-likelihood
  int32_t si1, si2, di1, di2;     // cluster of the singlets, pair index j(j+1)/2+k of the doublets (INT_MAX: none)
  double ssum, asum, smax, vmax;  // sum exp(llk + prior - max) over singlets / all, and the two maxima
  // form 1 (product domain, the fast epilogue of the row kernels): sv* / dv* are likelihoods scaled by 2^-es / 2^-ed,
  // ssum / asum the plain sums of the scaled singlet / doublet likelihoods (no priors), smax / vmax unused; the
  // decision kernel takes the logs.  form 0: everything above is in the log domain.
  int32_t es, ed, form, _pad;
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
  // E-step view of the same entries (static over the EM loop), per droplet at cell_ptr[c]:
  //   egen[cell_ptr[c] + r], r < gcnt[c]           ids of the entries with >= 2 informative reads (general)
  //   eab / esnp[cell_ptr[c] + gcnt[c] + r]        the others, packed: with at most one informative read the nine
  //                                                likelihoods are affine in g1 + g2 (alpha = 0.5: the read weight is
  //                                                1 - (g1+g2)/4), gl[g1*3+g2] = gl[0] + (g1+g2) * b, so the entry is
  //                                                the pair eab = {gl[0] / 2, b = (gl[8] - gl[0]) / 4} and its SNP
  uint8_t* eclass;           // [nnz] 1 = general
  int32_t* egen;             // [nnz]
  double2* eab;              // [nnz]
  int32_t* esnp;             // [nnz]
  int32_t* gcnt;             // [nbcs]
  uint8_t* snp_gen;          // [nsnps] 1 = some droplet has a general entry at the SNP (the only SNPs whose posterior
                             //         TRIPLES the E-step for K <= kFmxAffineMaxK reads; the others need the mean genotype)
  // SNP-major copy for the M-step
  int64_t* csc_ptr;          // [nsnps+1]
  int32_t* csc_cell;         // [nnz] droplet id, ascending inside a SNP
  double* csc_row;           // [nnz][kCscRow]: gl[0..8], then nreads / nref / nalt packed into the tenth slot
  // EM state
  int32_t* assign;           // [nbcs] cluster of the droplet in the next M-step, or -1
  double* stats;             // [K][nsnps][kStatStride]
  double* cgp;               // [nsnps][K][3] per-cluster genotype posteriors of the current iteration
  double* cgm;               // [nsnps][K] their mean genotype p[1] + 2 p[2] (what a single-read entry needs)
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
// Peer-memory exchange (one handle per GPU, NVLink / NVSwitch): the posterior, mean-genotype and assignment arrays
// and the flag array of every rank ([rank] = this handle's own), mapped through CUDA IPC or, inside one process, peer
// access.  n <= 1: single GPU, nothing is pushed.
constexpr int kMaxPeers = 16;
struct FmxPeers {
  double* post[kMaxPeers];
  double* postm[kMaxPeers];
  int32_t* assign[kMaxPeers];
  uint32_t* flags[kMaxPeers];  // [kMaxPeers + 2]: arrival epochs written by the peers, own epoch, error flag
  int32_t n;
};
// How a kernel of the loop takes part in the cross-rank ordering.  Two phases per iteration: the M-step kernel
// (producer of the posteriors) arrives at its tail and the E-step kernel waits at its head; the decision kernel
// (producer of the assignments) arrives at its tail and the M-step kernel waits at its head.
//   arrive: every CTA fences its stores (system scope) and counts itself in; the last one bumps this rank's epoch
//           and release-stores it into its slot of every peer's flag array;
//   wait:   one thread per peer acquire-spins on the own flag array until every slot has reached the own epoch, then
//           the CTA barrier releases the other threads.  A wait longer than timeout_cycles raises the error flag
//           (flags[kMaxPeers + 1]); every kernel of the loop returns at once when it finds the flag raised.
// wait = arrive = 0: the caller orders the ranks itself (fmx_sync_kernel between the kernels, or a collective).
struct FmxSync {
  int32_t wait, arrive, rank, _pad;
  long long timeout_cycles;
  unsigned int* counter;  // CTA arrival counter (device memory of this handle, zero between kernels)
};
// M-step + posteriors (+ all-gather): rebuilds the statistics of this rank's SNP slab, derives the posteriors and
// mean genotypes the next E-step reads (apply_geno_error: cmd_cram_freemuxlet.cpp:485-489 for that E-step) and
// stores them into every rank's arrays; want_stats = 0 skips the 96-byte statistics records (only the output after
// the last iteration needs them).
cudaError_t launch_fmx_mstep(const FmxDev& d, int apply_geno_error, int want_stats, const FmxPeers& peers, const FmxSync& sync,
                             cudaStream_t st);
// posteriors + mean genotypes from the statistics array (all SNPs, or this slab with zeros elsewhere)
cudaError_t launch_fmx_cgp(const FmxDev& d, int apply_geno_error, bool slab_only, cudaStream_t st);
// mean genotypes from the posterior array (all SNPs)
cudaError_t launch_fmx_cgm(const FmxDev& d, cudaStream_t st);
// one-warp kernel: arrive and / or wait (the barrier between un-fused kernels; the arrival of a rank whose shard
// is empty)
cudaError_t launch_fmx_sync(const FmxPeers& peers, const FmxSync& sync, cudaStream_t st);
cudaError_t preload_fmx_sync();
// per-droplet entry classes -> egen / eab / esnp / gcnt
cudaError_t launch_fmx_compact(const FmxDev& d, cudaStream_t st);
// E-step + decision for this rank's droplets; the decision kernel also stores the assignments into every rank's vector
cudaError_t launch_fmx_estep(const FmxDev& d, const FmxPeers& peers, const FmxSync& sync, cudaStream_t st);
// int32 count of droplets whose (type, jBest, kBest) differ between two result arrays
// counts bitwise differences between the shared-reciprocal division of the pileup kernels and __ddiv_rn
cudaError_t launch_fmx_div_selftest(uint64_t seed, int64_t n, unsigned long long* mismatches, cudaStream_t st);
cudaError_t launch_fmx_unpack_stats(const FmxDev& d, ccu_pileup* out, cudaStream_t st);

}  // namespace ccu
