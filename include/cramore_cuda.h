/*
 * cramore_cuda.h -- C ABI of libcramore_cuda.so, the B200 (sm_100a) droplet-likelihood engine
 * behind `cramore demuxlet` / `cramore freemuxlet`.
 *
 * The reference (hyunminkang/cramore) has no plugin or FFI seam: the engine is inline in
 * cmdCramDemuxlet (cmd_cram_demuxlet.cpp:427-1060) and cmdCramFreemuxlet
 * (cmd_cram_freemuxlet.cpp:116-164, 359-370, 456-653).  This header defines the seam a patched
 * command driver calls once its sc_dropseq_lib_t is populated (cmd_cram_demuxlet.cpp:425,
 * cmd_cram_freemuxlet.cpp:83).  Each entry point cites the reference lines it replaces.
 *
 * Conventions
 *   - plain C types only; every function returns 0 on success, non-zero on error; the message is
 *     available from ccu_last_error().  No exceptions cross the boundary, no CPU fallback exists:
 *     without an sm_100 device ccu_create() fails.
 *   - one handle drives ONE GPU.  Multi-GPU = one process (or handle) per GPU over disjoint
 *     barcode shards, the reference's own --group-list recipe (cmd_cram_demuxlet.cpp:75).
 *   - all array arguments are HOST pointers borrowed for the duration of the call unless the
 *     name says _dev.  The engine owns all device memory.
 *
 * CSR layout of the droplet store (flattened sc_dropseq_lib_t::cell_umis, sc_drop_seq.h:165):
 *   cell_ptr[nbcs+1]  -> first (cell,SNP) entry of each barcode (entries ascend by SNP id)
 *   snp_idx[nnz]      -> SNP id of each entry
 *   read_ptr[nnz+1]   -> first read of each entry (reads in lexicographic UMI order)
 *   al[R], bq[R]      -> allele code {0 ref, 1 alt, 2 other} and capped base quality per unique read
 */
#ifndef CRAMORE_CUDA_H
#define CRAMORE_CUDA_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CCU_MAX_ALPHA 32 /* size of the --alpha grid the engine accepts */

typedef struct ccu_handle ccu_handle;

/* Exactly the per-droplet values consumed by the decision + row code
 * (cmd_cram_demuxlet.cpp:788-791 declarations, :921-1013 consumers). */
typedef struct {
  int32_t sBest, sNext;            /* :827-837 */
  int32_t dBest1, dBest2, dBestA;  /* :883-906 (dBestA indexes the alpha grid) */
  int32_t dNext1, dNext2, dNextA;
  double sngBestLLK, sngNextLLK;
  double dblBestLLK, dblNextLLK;
  double sumLLK, sngLLK;           /* :804-821 */
} ccu_demux_result;

/* Decision of cmd_cram_demuxlet.cpp:921-991.  type: 0 SNG, 1 DBL, 2 AMB. */
typedef struct {
  int32_t type;
  int32_t jBest, kBest, alphaBest;
  int32_t jNext, kNext, alphaNext;
  int32_t _pad;
  double bestLLK, nextLLK;
  double bestPP, sngPP, sngOnlyPP;
} ccu_demux_call;

/* Device timings of the last ccu_demux_run()/ccu_demux_score(), CUDA events on the engine's
 * stream (milliseconds), and the number of kernels it launched. */
typedef struct {
  float ms_genotypes;  /* ccu_demux_set_genotypes kernels (last call) */
  float ms_tile;       /* K1  per-(cell,SNP) tile kernel */
  float ms_singlet;    /* K2s singlet scoring kernel */
  float ms_score;      /* K2  pair x alpha scoring kernel */
  float ms_finalize;   /* K3  per-droplet reduction */
  float ms_run;        /* all kernels of one run, first launch to last completion */
  float ms_h2d, ms_d2h;
  int32_t launches;    /* kernels launched by the last run */
  int32_t score_ctas;  /* grid size of K2 */
  int64_t score_items; /* (barcode x pair-tile x alpha-chunk) work items of K2 */
} ccu_timings;

/* ---- lifetime -------------------------------------------------------------------------- */
int ccu_create(int32_t device, ccu_handle** out);
int ccu_destroy(ccu_handle* h);
/* Message of the last failed call on h (h == NULL: last failed ccu_create). */
const char* ccu_last_error(const ccu_handle* h);
/* Run on the caller's CUDA stream (cudaStream_t passed as void*); NULL = engine-owned stream. */
int ccu_set_stream(ccu_handle* h, void* cuda_stream);
int ccu_synchronize(ccu_handle* h);
int ccu_get_timings(const ccu_handle* h, ccu_timings* t);

/* ---- demuxlet -------------------------------------------------------------------------- */

/* --alpha grid / --doublet-prior / sample count (cmd_cram_demuxlet.cpp:62-63,85-89,103).
 * alpha[0] must be 0, nalpha in [2, CCU_MAX_ALPHA], nv >= 2 (the reference is
 * undefined otherwise). */
int ccu_demux_configure(ccu_handle* h, int32_t nv, int32_t nalpha, const double* alpha,
                        double doublet_prior);

/* Genotype posteriors, replaces the host loop sc_drop_seq.cpp:287-315
 * (== cmd_cram_demuxlet.cpp:241-268): gp is BCFFilteredReader::get_posterior_at output (float,
 * bcf_filtered_reader.h:168-170) as [nsnps][nv][3]; snp_err[s] = offset + (1-offset)*(1-R2)*coeff
 * (clamped to [0,0.999] by the engine like :308-309); has_gp[s]==0 marks sc_snp_t::gps==NULL
 * (NULL pointer = every SNP has genotypes).  The engine forms the per-SNP average and the mixed
 * FP64 matrix on the device, bit-identical to the host loop. */
int ccu_demux_set_genotypes(ccu_handle* h, int64_t nsnps, const float* gp, const double* snp_err,
                            const uint8_t* has_gp);
/* The same from the reference's own per-SNP arrays: gps[s] = sc_snp_t::gps of SNP s (sc_drop_seq.h:115-127: nv*3
 * doubles, genotype error already mixed in by sc_drop_seq.cpp:287-315 / cmd_cram_demuxlet.cpp:241-268), or NULL for
 * a SNP without genotypes.  Lets a patched cmdCramDemuxlet pass `scl.snps[s].gps` through unchanged.
 * The rows are taken as they are (no validation): they must be genotype PROBABILITIES -- non-negative, each sample's
 * three values summing to about 1, as the reference's mixing leaves them.  The scoring kernel renormalises its running
 * products on a budget that assumes every per-SNP factor is at least ~2^-34 (the read-error floor times a row sum
 * near 1); rows that sum to ~0 make a droplet's likelihood exactly 0 (log-likelihood -inf) rather than tiny. */
int ccu_demux_set_genotypes_mixed(ccu_handle* h, int64_t nsnps, const double* const* gps);

/* Multi-GPU form of ccu_demux_set_genotypes (no reference counterpart: the reference is one process and scales by
 * --group-list, cmd_cram_demuxlet.cpp:75, every process parsing the whole VCF).  The genotype matrix is replicated
 * on every GPU, but only ONE copy has to cross PCIe: ccu_demux_genotypes_stage reserves the device arrays
 * (float [nsnps][nv][3], double [nsnps], uint8 [nsnps]) and returns their DEVICE addresses; the caller fills them on
 * the engine's stream -- each rank copies its own SNP slab host->device and an all-gather over NVLink (NCCL, or
 * cudaMemcpyPeerAsync inside one process) completes the matrix -- and ccu_demux_genotypes_commit runs the mixing
 * kernels of sc_drop_seq.cpp:287-315 on what is there.  use_has_gp == 0: every SNP has genotypes. */
int ccu_demux_genotypes_stage(ccu_handle* h, int64_t nsnps, void** gp_dev, void** snp_err_dev, void** has_gp_dev);
int ccu_demux_genotypes_commit(ccu_handle* h, int64_t nsnps, int32_t use_has_gp);
/* One process, several handles (one per device): dst takes the mixed matrix of src (whose genotypes were set by any of
 * the calls above) over NVLink instead of uploading it again.  Both handles configured for the same sample count. */
int ccu_demux_copy_genotypes(ccu_handle* dst, ccu_handle* src);

/* The whole engine for nbcs droplets, replaces cmd_cram_demuxlet.cpp:636-906 (tile :655-725,
 * pairwise accumulation :733-747, posterior normaliser :804-821, top-2 scans :827-906).
 * out[nbcs] in CSR barcode order.  Equivalent to upload + run + download -- with the same results bit for bit, but a
 * large store (4,096 droplets and 2^20 entries or more) is scored in two parts, the second part's host->device copy
 * under the first part's kernels, and a store whose scratch does not fit the free device memory in as many droplet
 * ranges as it takes.  The handle then holds only the LAST part: ccu_demux_get_llk / _get_tiles / _get_singlets /
 * _results_dev refuse it (error message); to inspect a store, use ccu_demux_upload + ccu_demux_run, or set
 * CCU_DEMUX_NO_PIPELINE=1 in the environment. */
int ccu_demux_score(ccu_handle* h, int64_t nbcs, const int64_t* cell_ptr, const int32_t* snp_idx,
                    const int64_t* read_ptr, const uint8_t* al, const uint8_t* bq,
                    ccu_demux_result* out);

/* Staged form of the same call (inputs stay resident in HBM between runs). */
int ccu_demux_upload(ccu_handle* h, int64_t nbcs, const int64_t* cell_ptr, const int32_t* snp_idx,
                     const int64_t* read_ptr, const uint8_t* al, const uint8_t* bq);
int ccu_demux_run(ccu_handle* h);  /* asynchronous on the stream */
int ccu_demux_wait(ccu_handle* h); /* wait for the last run and refresh ccu_get_timings() */
int ccu_demux_download(ccu_handle* h, ccu_demux_result* out);
/* Device address of the ccu_demux_result[nbcs] array of the last run (valid until the next upload): lets a
 * multi-GPU driver gather the shards' records over NVLink (ncclAllGather) instead of through the host. */
void* ccu_demux_results_dev(ccu_handle* h, int64_t* nbcs);

/* Decision + row formatting (host code, cmd_cram_demuxlet.cpp:921-1013). */
int ccu_demux_classify(const ccu_demux_result* r, int32_t nv, int32_t nalpha, const double* alpha,
                       double doublet_prior, ccu_demux_call* call);
/* Returns the length of the row written to buf (newline included); -1: the result names an undefined sample or alpha
 * index; -2: the row does not fit buflen (long sample ids / barcodes: retry with a larger buffer).  Header line:
 * ccu_demux_best_header(). */
int ccu_demux_best_row(char* buf, int32_t buflen, int32_t int_id, const char* barcode,
                       uint32_t nsnps_cell, int32_t nreads_cell, const ccu_demux_result* r, int32_t nv,
                       int32_t nalpha, const double* alpha, double doublet_prior,
                       const char* const* sample_ids);
const char* ccu_demux_best_header(void);

/* Text-level drop-in for <out>.best: re-scores the candidates a row can name -- the two best singlets and the two
 * best doublets the engine reported, plus the mirror orientation (k, j, 1 - alpha) of each of those doublets where
 * the grid holds 1 - alpha (at alpha = 0.5: (k, j)) -- in the reference's own IEEE operation sequence
 * (cmd_cram_demuxlet.cpp:655-747: per-read tile loop, (g_j[l]*g_k[m])*pG l-outer / m-inner, sum of the host libm's
 * log in SNP order) and re-applies the reference's top-2 update rule (:827-837, :883-906) to them, in place.
 * Afterwards the indices, the orientation of mirror pairs and every printed LLK equal the reference's bit for bit
 * (the engine's own values agree to ~1e-13 relative, which cannot decide a last-bit tie).  Host code, threads over
 * droplets (nthreads <= 0: all cores), O(entries x (9 nalpha + candidates)): it decides the print, the search over
 * all hypotheses stays on the GPU.  Needs no handle and no device.
 *   ccu_demux_refine        genotypes as for ccu_demux_set_genotypes (float posteriors + per-SNP error + flags)
 *   ccu_demux_refine_mixed  genotypes as for ccu_demux_set_genotypes_mixed (the reference's scl.snps[s].gps) */
int ccu_demux_refine(int64_t nbcs, const int64_t* cell_ptr, const int32_t* snp_idx, const int64_t* read_ptr,
                     const uint8_t* al, const uint8_t* bq, int32_t nv, int32_t nalpha, const double* alpha,
                     int64_t nsnps, const float* gp, const double* snp_err, const uint8_t* has_gp, int32_t nthreads,
                     ccu_demux_result* inout);
int ccu_demux_refine_mixed(int64_t nbcs, const int64_t* cell_ptr, const int32_t* snp_idx, const int64_t* read_ptr,
                           const uint8_t* al, const uint8_t* bq, int32_t nv, int32_t nalpha, const double* alpha,
                           int64_t nsnps, const double* const* gps, int32_t nthreads, ccu_demux_result* inout);

/* <out>.single / <out>.sing2 (named by the command's documentation; the writers are commented out in the reference,
 * cmd_cram_demuxlet.cpp:516, :552-563, :839-848): every sample's singlet log-likelihood of the last run,
 * llk[nbcs][nv] = llksAB[j][0][0], and -- llk0 != NULL -- llks00[0] per droplet (:427-440, :762-773: both samples
 * unspecified, genotypes drawn from the pool's mean distribution at the SNP).  Parity unpinned by construction: no
 * build of the reference writes these files. */
int ccu_demux_get_singlets(ccu_handle* h, double* llk, double* llk0);

/* Inspection (parity tests): copy device intermediates of the last upload/run back.
 * tiles: [nnz][nalpha*9] (cmd_cram_demuxlet.cpp:655-725); gps: [snp_end-snp_begin][nv][3]
 * (sc_snp_t::gps); llk: [cell_end-cell_begin][nv][nv][nalpha] = llksAB (:746) where the engine
 * computes it ([j][0][0] and every [j][k][n>=1]); other entries are NaN.  ccu_demux_get_llk
 * re-runs the scoring kernel for that barcode range. */
int ccu_demux_get_tiles(ccu_handle* h, double* tiles);
int ccu_demux_get_genotypes(ccu_handle* h, int64_t snp_begin, int64_t snp_end, double* gps);
int ccu_demux_get_llk(ccu_handle* h, int64_t cell_begin, int64_t cell_end, double* llk);


/* ---- freemuxlet ------------------------------------------------------------------------ */

/* snp_droplet_pileup (sc_drop_seq.h:65-75) flattened: one per (cell,SNP) entry or per
 * (cluster,SNP).  logdenom is carried for per-entry pileups only: no live output of the reference
 * reads it (it is dead state of snp_droplet_pileup::merge), so cluster pileups report 0. */
typedef struct {
  int32_t nreads, nref, nalt, _pad;
  double gls[9];
  double logdenom;
} ccu_pileup;

/* Per-droplet outcome of one E-step + decision pass (cmd_cram_freemuxlet.cpp:524-637): exactly the
 * arrays the reference keeps per droplet and prints into <out>.clust1.samples.gz (:709-711).
 * type: 0 SNG, 1 DBL, 2 AMB. */
typedef struct {
  int32_t sBest, sNext, dBest1, dBest2, dNext1, dNext2;
  int32_t type;
  int32_t jBest, kBest, jNext, kNext, _pad;
  double sngBestLLK, sngNextLLK, dblBestLLK, dblNextLLK;
  double sumLLK, sngLLK;
  double bestLLK, nextLLK, bestPP, sngPP, sngOnlyPP;
} ccu_fmx_result;

/* --nsample / --doublet-prior / --geno-error (cmd_cram_freemuxlet.cpp:34-63).  2 <= nclust <= 64. */
int ccu_fmx_configure(ccu_handle* h, int32_t nclust, double doublet_prior, double geno_error);

/* The droplet store (same CSR as demuxlet, ALL droplets) plus the per-SNP allele frequency
 * (sc_snp_t::af, .var.gz column 6).  Computes the per-(cell,SNP) pileup likelihoods at alpha = 0.5
 * (calculate_snp_droplet_pileup, sc_drop_seq.cpp:452-509, as called at
 * cmd_cram_freemuxlet.cpp:133) and the SNP-major copy the M-step folds over. */
int ccu_fmx_upload(ccu_handle* h, int64_t nbcs, int64_t nsnps, const int64_t* cell_ptr, const int32_t* snp_idx,
                   const int64_t* read_ptr, const uint8_t* al, const uint8_t* bq, const double* af);

/* Multi-GPU partition: this handle runs the E-step for droplets [cell_begin, cell_end)
 * and the M-step for SNPs [snp_begin, snp_end).  Default after upload: everything. */
int ccu_fmx_set_shard(ccu_handle* h, int64_t cell_begin, int64_t cell_end, int64_t snp_begin, int64_t snp_end);

/* .lmix scores (cmd_cram_freemuxlet.cpp:116-164): llk0 = DBL.LLK, llk2 = SNG.LLK per droplet [nbcs]. */
int ccu_fmx_lmix(ccu_handle* h, double* llk0, double* llk2);
int ccu_fmx_get_pileups(ccu_handle* h, ccu_pileup* out /* [nnz] */);

/* M-step (cmd_cram_freemuxlet.cpp:359-370 and :639-650): rebuild every cluster pileup as the ordered
 * fold snp_droplet_pileup::merge (sc_drop_seq.h:77-101) over the droplets with assign[c] >= 0 in
 * ascending droplet id.  assign: host int32 [nbcs] (cluster id or -1).  Fills the SNP slab of this
 * shard and zeroes the rest of the statistics array, so that a sum-allreduce over ranks assembles the
 * full array bit-exactly. */
int ccu_fmx_mstep(ccu_handle* h, const int32_t* assign);
/* E-step + decision (cmd_cram_freemuxlet.cpp:465-637) for this shard's droplets.
 * out: host [cell_end-cell_begin]; assign_out (optional): host int32 [cell_end-cell_begin], the
 * M-step input of the next iteration (jBest for confident singlets, else -1, :641-643). */
int ccu_fmx_estep(ccu_handle* h, int32_t apply_geno_error, ccu_fmx_result* out, int32_t* assign_out);
/* llks[pair], [cells of the shard][K(K+1)/2], pair = j(j+1)/2+k (inspection).  An E-step does not keep the matrix: this
 * call runs the E-step again on the cluster posteriors as they stand (right after ccu_fmx_estep: the ones it used) with
 * the epilogue that takes every pair's log, and rewrites results and assignments with the same values. */
int ccu_fmx_get_llk(ccu_handle* h, double* llk);
int ccu_fmx_get_clusters(ccu_handle* h, ccu_pileup* out /* [K][nsnps] */);

/* Device-resident form for the EM loop and its collectives.  The statistics array is
 * double [K][nsnps][12] = {gls[9], nreads, nref, nalt}; the assignment vector is int32 [nbcs].
 * Both pointers stay valid until the next ccu_fmx_upload / ccu_destroy.  A multi-GPU driver runs,
 * per iteration:  ccu_fmx_estep_dev -> all-gather of the assignment vector (each rank wrote its own
 * droplet range) -> ccu_fmx_mstep_dev -> sum-allreduce of the statistics array (NCCL, in place). */
void* ccu_fmx_stats_dev(ccu_handle* h, int64_t* n_doubles);
/* Lighter per-iteration exchange: the E-step reads the statistics only through the per-(SNP, cluster) genotype
 * posteriors of cmd_cram_freemuxlet.cpp:476-489 (3 doubles instead of 12).  ccu_fmx_posteriors_dev computes
 * them for this shard's SNP slab from its slab of the statistics and zeroes the rest of the array
 * double [nsnps][K][3] returned by ccu_fmx_post_dev; after a sum-allreduce of that array
 * ccu_fmx_estep_post_dev runs the E-step + decision from it.  Per iteration:
 * ccu_fmx_posteriors_dev -> allreduce(post) -> ccu_fmx_estep_post_dev -> all-gather(assign) -> ccu_fmx_mstep_dev;
 * the statistics array itself is all-reduced once, after the last iteration (cluster VCF output). */
void* ccu_fmx_post_dev(ccu_handle* h, int64_t* n_doubles);
int ccu_fmx_posteriors_dev(ccu_handle* h, int32_t apply_geno_error);
int ccu_fmx_estep_post_dev(ccu_handle* h);
/* The fused loop (what ccu_fmx_run_em and the multi-GPU drivers run): two calls per iteration.
 *   ccu_fmx_mstep_fused_dev  M-step of this shard's SNP slab; the kernel forms the genotype posteriors (and their mean
 *                            genotypes) of the slab from the registers that hold the finished statistics -- with the
 *                            genotype-error mix of cmd_cram_freemuxlet.cpp:485-489 iff apply_geno_error_next, i.e. the
 *                            flag of the E-step that FOLLOWS -- and stores them into its own arrays and, of them, what
 *                            the peers' E-step reads into the arrays of every other rank: the mean genotypes, and the
 *                            posterior triples at the SNPs where some droplet has an entry with two or more reads
 *                            (every SNP when K > 16).  After a fused loop with peers a rank's posterior array is
 *                            therefore complete for its own slab only; the statistics are the loop's output;
 *                            want_stats = 0 skips the 96-byte statistics records (only the output after the last
 *                            iteration reads them)
 *   ccu_fmx_estep_fused_dev  E-step + decision of this shard's droplets; the decision kernel stores the new
 *                            assignments into the vector of every rank
 * Peer-memory exchange (no reference counterpart: the reference is one process).  With one process (or handle) per
 * GPU on an NVLink / NVSwitch box the two per-iteration exchanges need no collective on the payload: every rank maps
 * the posterior, mean-genotype and assignment arrays of all ranks (CUDA IPC across processes, peer access inside one)
 * and the producing kernels store straight into them.
 *   setup, after ccu_fmx_set_shard: ccu_fmx_peer_handles(h, mine) -> all-gather of the 4 * CCU_IPC_HANDLE_BYTES bytes
 *          (any host channel) -> ccu_fmx_peer_open(h, nranks, rank, all)
 * A rank may read its arrays only after every rank's producing kernel has completed, and may be written to again
 * only after it has finished reading.  in_kernel_sync = 1 makes the kernels do that themselves: the last CTA of a
 * producer release-stores the rank's next epoch into its slot of every peer's flag array, the consumer kernel
 * acquire-spins at its head until every slot of its own array has got there -- no launch, no collective, and the
 * stores of one CTA travel over NVLink while the others still compute.  A wait longer than
 * CCU_FMX_BARRIER_TIMEOUT_S (default 30 s) raises an error flag instead of hanging the GPU, every later kernel of
 * the loop then returns at once, and ccu_fmx_rank_barrier_failed reports it (message: rank, epoch, ranks behind).
 * in_kernel_sync = 0 leaves the ordering to the caller: ccu_fmx_rank_barrier_dev (the same protocol as one one-warp
 * kernel; required when several handles share ONE device, where a kernel spinning at its head could keep the peer it
 * waits for from being scheduled) or a collective on the engine's stream after each of the two calls.
 * Without peers (single GPU) the two calls are simply the loop of cmd_cram_freemuxlet.cpp:457-653.  Results are
 * bit-identical whatever the number of ranks.  ccu_fmx_peer_close unmaps (also done by ccu_destroy). */
#define CCU_IPC_HANDLE_BYTES 64
/* The same exchange inside ONE process: n handles on different devices (or several on one), each configured and
 * uploaded with the same store.  One host thread per handle runs the loop of cmd_cram_freemuxlet.cpp:359-370 +
 * :457-653 on its droplet range and SNP slab (about nnz / n entries each), peers are reached through
 * cudaDeviceEnablePeerAccess, the two per-iteration exchanges happen inside the fused kernels above.
 * cell_ptr / snp_idx: the host arrays that were uploaded (for the partition).  out: host [nbcs];
 * clusters_out: optional host [nclust][nsnps], the cluster pileups after the last M-step.
 * Bit-identical to ccu_fmx_run_em (no early stop: freemux2's stop_when_stable needs a host decision per iteration). */
int ccu_fmx_run_em_multi(ccu_handle** handles, int32_t n, const int64_t* cell_ptr, const int32_t* snp_idx,
                         const int32_t* init_assign, int32_t niter, int32_t geno_error_every_iter, ccu_fmx_result* out,
                         ccu_pileup* clusters_out);
int ccu_fmx_peer_handles(ccu_handle* h, void* out /* [4 * CCU_IPC_HANDLE_BYTES]: posteriors, mean genotypes, assignments, flags */);
int ccu_fmx_peer_open(ccu_handle* h, int32_t nranks, int32_t rank, const void* handles /* [nranks][4][64] */);
int ccu_fmx_rank_barrier_dev(ccu_handle* h);
int ccu_fmx_rank_barrier_failed(ccu_handle* h, int32_t* failed);
int ccu_fmx_peer_close(ccu_handle* h);
int ccu_fmx_mstep_fused_dev(ccu_handle* h, int32_t apply_geno_error_next, int32_t want_stats, int32_t in_kernel_sync);
int ccu_fmx_estep_fused_dev(ccu_handle* h, int32_t in_kernel_sync);
void* ccu_fmx_assign_dev(ccu_handle* h, int64_t* n_int32);
int ccu_fmx_mstep_dev(ccu_handle* h);
int ccu_fmx_estep_dev(ccu_handle* h, int32_t apply_geno_error);
int ccu_fmx_download_results(ccu_handle* h, ccu_fmx_result* out);
/* Single-GPU EM loop cmd_cram_freemuxlet.cpp:457-653: niter iterations of E-step, decision and
 * M-step starting from init_assign (host int32 [nbcs], the clusts[] of :359-370); geno_error is
 * applied on the last iteration only (freemuxlet, :485) or on every iteration (freemux2,
 * cmd_cram_freemux2.cpp:410-415) as every_iter says; stop_when_stable adds freemux2's early stop
 * (:601-604).  out: host [nbcs]; returns the number of iterations run in *iters_run. */
int ccu_fmx_run_em(ccu_handle* h, const int32_t* init_assign, int32_t niter, int32_t geno_error_every_iter,
                   int32_t stop_when_stable, ccu_fmx_result* out, int32_t* iters_run);

/* Device timings (ms) of the last freemuxlet calls. */
typedef struct {
  float ms_pileup, ms_csc, ms_lmix, ms_cgp, ms_estep, ms_mstep;
  int32_t launches;
  int32_t _pad;
} ccu_fmx_timings;
int ccu_fmx_get_timings(const ccu_handle* h, ccu_fmx_timings* t);

/* ---- initial clustering of freemuxlet (cmd_cram_freemuxlet.cpp:166-346) ------------------------------------
 * One record per pair of droplets that share at least one SNP = the non-zero entries of the reference's dense
 * lower-triangular dropDs[id1][id2] matrix (sc_drop_seq.h:45-52), id1 > id2, sorted by (id1, id2). */
typedef struct {
  int32_t id1, id2;       /* droplets, id1 > id2 */
  int32_t nsnps;          /* SNPs covered by both */
  int32_t nread1, nread2; /* reads of id1 / id2 over those SNPs */
  int32_t _pad;
  double llk0, llk2;      /* sum_snp log lk0 (different individuals), log lk2 (same individual), :202-214 */
} ccu_pair_dist;

/* Pairwise distances (:186-221) from the uploaded droplet store, on the device.  Returns the number of pairs,
 * the number of (pair, SNP) records they were folded from and the device time. */
int ccu_fmx_pair_distances(ccu_handle* h, int64_t* npairs, int64_t* nrecords, float* ms);
int ccu_fmx_get_pair_distances(ccu_handle* h, ccu_pair_dist* out /* [npairs] */);
/* Only the pairs that can cast a vote in the clustering below, |llk2 - llk0| > bf_thres (compacted on the device,
 * same order): call once with out == NULL for the count, then with a buffer of that many records.  The votes they
 * give are identical to those of the full list -- the other pairs never touch a vote (:270-277, :313-318). */
int ccu_fmx_get_voting_pairs(ccu_handle* h, double bf_thres, int64_t* nvoting, ccu_pair_dist* out);

/* Host part of the initial clustering (:223-346): greedy assignment in the order of the singlet scores
 * (cell_scores[i] = SNG.LLK - DBL.LLK of .lmix, ties by larger id first, sc_drop_seq.h:190-198) by votes of the
 * already assigned droplets whose Bayes factor against the candidate passes bf_thres, then -- if iter_init > 0 --
 * ten sweeps in std::random_shuffle order.  Every random number comes from rand(), in the reference's call
 * order; seed >= 0 calls srand(seed) first (the reference never seeds: seed 1).  clusts: int32 [nbcs]; with
 * have_init != 0 it holds the --init-cluster assignment (-1: none) and the greedy pass is skipped.  Needs no GPU
 * and no handle. */
int ccu_fmx_initial_clusters(int64_t nbcs, int32_t nclust, const double* cell_scores, int64_t npairs,
                             const ccu_pair_dist* pairs, double bf_thres, double frac_init_clust, int32_t iter_init,
                             int32_t keep_init_missing, int32_t have_init, int32_t seed, int32_t* clusts);

/* freemux2's greedy initialisation (cmd_cram_freemux2.cpp:217-261): droplets in singlet-score order, each assigned to
 * the cluster whose running pileup it fits best (sc_drop_seq.cpp:544-578) and merged into it (sc_drop_seq.h:77-101).
 * plp: the per-entry pileups of ccu_fmx_get_pileups.  Sequential and cheap (entries x nclust): host code, no GPU. */
int ccu_fmx_greedy_clusters(int64_t nbcs, int64_t nsnps, int32_t nclust, const int64_t* cell_ptr, const int32_t* snp_idx,
                            const ccu_pileup* plp, const double* af, const double* cell_scores, double frac_init_clust,
                            double singlet_score_thres, int32_t* clusts);

/* Diagnostic: the pileup kernels divide the nine likelihoods of a pileup by one sum with a shared refined
 * reciprocal (sc_drop_seq.cpp:492-506 and sc_drop_seq.h:85-100 do nine divisions).  Runs `nsets` random
 * operand sets on the device and returns how many of the 9*nsets quotients differ bitwise from IEEE
 * division (must be 0). */
int ccu_fmx_selftest_division(ccu_handle* h, int64_t nsets, uint64_t seed, int64_t* mismatches);

/* Measured FP64 roofline denominator: runs a DFMA-saturating microkernel for about `ms`
 * milliseconds on the handle's device and returns TFLOP/s (2 flops per DFMA). */
int ccu_measure_fp64_peak(ccu_handle* h, float ms, double* tflops);

#ifdef __cplusplus
}
#endif
#endif
