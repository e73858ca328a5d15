// demux_internal.h -- device-side data layout and launch wrappers shared by the C-ABI host
// code (ccu_api.cu) and the demuxlet kernels (demux_kernels.cu).  Not part of the public ABI.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "cramore_cuda.h"

namespace ccu {

constexpr int kTileStride = 10;  // doubles per alpha in the device tile ([nnz][nA][10]; 9 used, 80 B
                                 // keeps every alpha block 16-byte aligned for cp.async.bulk)
constexpr int kJT = 32;          // samples j per scoring CTA tile
constexpr int kScoreThreads = 512;  // consumer threads of the scoring CTA

// One (best or next) candidate of the doublet scan in the product domain:
// LLK = log(m) + e*ln2, pos = (j*nv + k)*nA + n  (scan order of cmd_cram_demuxlet.cpp:883-906).
struct Cand {
  double m;
  int32_t e;
  int32_t pos;  // -1 = empty
};

constexpr int kScoreWarps = kScoreThreads / 32;  // partial records per work item (one per consumer warp)

// Partial reduction of one consumer warp over one scoring work item (barcode x pair tile x alpha chunk).
struct ItemPartial {
  double lse_a;    // sum over the item's doublet terms of w * m * 2^(e - lse_e), w per :812-817
  int32_t lse_e;   // INT_MIN/2 when the item has no term
  int32_t _pad;
  Cand best, next;
};

struct DemuxDev {
  // configuration
  int32_t nv, nvp, nA;
  double doublet_prior;
  const double* alpha;     // [nA]
  double alpha_c[CCU_MAX_ALPHA];  // the same grid by value: lands in the kernel's constant bank
  const double* ptab;      // [nA*9*2]: p then (1-p), cmd_cram_demuxlet.cpp:673,685
  const uint8_t* is_half;  // [nA] alpha[n] == 0.5
  const double* phred;     // [512]: err[256], mat[256]
  // genotypes
  int64_t nsnps;
  const double* G;         // [nsnps][nvp][3] mixed FP64 posteriors (zero rows/cols where absent)
  const uint8_t* has_gp;   // [nsnps]
  // droplet store
  int64_t nbcs, nnz, nreads;
  int64_t ngen;            // entries with two or more informative reads (counted at upload)
  const int64_t* cell_ptr;
  const int32_t* snp_idx;
  const int64_t* read_ptr;
  const uint8_t* al;
  const uint8_t* bq;
  // intermediates
  double* tiles;           // [general entries][nA][kTileStride] (tslot != NULL: the row of entry e is tslot[e]),
                           //       or [nnz][nA][kTileStride] with every row materialised (tslot == NULL:
                           //       ccu_demux_get_tiles)
  int32_t* tslot;          // [nnz] exclusive prefix sum of eclass: tile row of a general entry
  double* vaff;            // [nnz][4] {pG[0][0][.], pG[0][1][.], pG[0][2][.], 0} of every entry
  uint8_t* eclass;         // [nnz] 0 = at most one informative read (tile affine in the allele
                           //       fraction), 1 = general
  int64_t* vent;           // [nnz] per-cell compacted entry ids with genotypes: general entries
                           //       first, then the affine ones
  int32_t* vsnp;           // [nnz] SNP id of each compacted entry
  int32_t* vcnt_g;         // [nbcs] general entries of the cell
  int32_t* vcnt_a;         // [nbcs] affine entries of the cell
  double* sng_llk;         // [nbcs][nv]
  double* fmat;            // per cell at cell_ptr[c]*nvp: [nvp/32][vcnt_a][32] normalised single-read
                           //       likelihoods f = (sum_l g[l]*pG[0][l][0]) / (sum_l g[l])
  double* pm;              // [nbcs][nvp] mantissa of prod over affine entries of sum_l g[l]
  int32_t* pe;             // [nbcs][nvp] its binary exponent
  int32_t* maxbq;          // [1] largest base quality among informative reads (set by K1)
  ItemPartial* partials;   // [nbcs * items_per_cell * kScoreWarps]
  ccu_demux_result* results;  // [nbcs]
};

// Plan of the scoring kernel for (nv, nA): thread tile TJ x (TK k's x AC alphas).
struct ScorePlan {
  int ac, tk;          // alphas per chunk, k's per thread
  int kt;              // k's per CTA tile = 32*tk
  int njt, nkt, nac;   // tiles per cell
  int items_per_cell;
  size_t smem_bytes;
  uint8_t is_half[CCU_MAX_ALPHA];  // alpha[n] == 0.5
};
ScorePlan make_score_plan(int nv, int nA, const double* alpha);

// launch wrappers (all asynchronous on `st`); return cudaGetLastError()
cudaError_t launch_gp_prepare(const float* gp_f32, const double* snp_err, const uint8_t* has_gp,
                              int64_t nsnps, int nv, int nvp, double* avg_tmp, double* G,
                              cudaStream_t st);
cudaError_t launch_validate(const DemuxDev& d, int64_t nreads, uint32_t* flag, cudaStream_t st);
// number of entries with >= 2 informative reads -> *count_dev (what sizes the tile buffer; once per upload)
cudaError_t launch_count_general(const DemuxDev& d, unsigned long long* count_dev, cudaStream_t st);
// p[i] -= base for i < n (pointer arrays of a store slice uploaded with the whole store's offsets)
cudaError_t launch_rebase(int64_t* p, int64_t n, int64_t base, cudaStream_t st);
// scan_tmp / scan_bytes: scratch of the prefix sum over the entry classes (scan_bytes = 0: return the size needed in
// *scan_need and do nothing)
cudaError_t launch_tile(const DemuxDev& d, bool all_entries, void* scan_tmp, size_t scan_bytes, size_t* scan_need,
                        cudaStream_t st);
cudaError_t launch_compact(const DemuxDev& d, cudaStream_t st);
cudaError_t launch_sample(const DemuxDev& d, cudaStream_t st);
// dbg != NULL: also dump (mantissa, exponent) of every computed llksAB entry of cells [dbg_c0, dbg_c1)
// into [cells][nv][nv][nA] double2; launch_dbg_convert turns them into log-likelihoods
cudaError_t launch_score(const DemuxDev& d, const ScorePlan& plan, int num_sms, double2* dbg,
                         int64_t dbg_c0, int64_t dbg_c1, int* grid_out, cudaStream_t st);
cudaError_t launch_dbg_convert(double2* dbg, double* llk, int64_t n, cudaStream_t st);
cudaError_t launch_finalize(const DemuxDev& d, const ScorePlan& plan, cudaStream_t st);
cudaError_t launch_singlet_to_dbg(const DemuxDev& d, double* llk_dbg, int64_t c0, int64_t c1,
                                  cudaStream_t st);
// llks00[0] per droplet (gp0_tmp: [nsnps][4] scratch); after a run
cudaError_t launch_llk00(const DemuxDev& d, double* gp0_tmp, double* out, cudaStream_t st);
cudaError_t launch_fp64_peak(double* sink, int iters, int num_sms, cudaStream_t st);

}  // namespace ccu
