/*
 * oracle.h -- C interface of the CPU oracle (TEST INFRASTRUCTURE, not product code).
 *
 * The oracle is a plain C++14 restatement of the reference's single-threaded
 * demuxlet / freemuxlet likelihood loops, operating on the same flat CSR + GP
 * arrays the CUDA engine takes.  Only tests/, __graft_entry__.smoke() and the
 * cpu_baseline / --impl reference legs of bench.py may load it.
 *
 * PARITY STATUS: PINNED against the reference's own compiled engine.  The reference cannot be
 * built with its own build system here (htslib, an un-vendored external library, is absent), but
 * its likelihood code compiles unmodified from where it lies once the four htslib-facing I/O
 * headers are replaced by the zlib stand-ins of oracle/ref_shim/: `make -C oracle ref` builds
 * oracle/_ref/ref_cramore from /root/reference/{cmd_cram_demuxlet,cmd_cram_freemuxlet,sc_drop_seq,
 * params,Error,PhredHelper}.cpp.  tests/golden/make_ref_fixtures.py runs it on synthetic digital
 * pileups + VCFs and records every value it passes to hprintf at full precision
 * (tests/golden/ref_*.json); tests/test_reference_fixtures.py holds this oracle to those values
 * (log-likelihoods to 1e-12, identical calls, byte-identical .best rows, full freemuxlet EM).
 * Also pinned: the Phred tables against PhredHelper.cpp (oracle/_ref/libphred_ref.so) and the
 * hand-derived known answers of SURVEY.md App. C.  What stays restated (not reference code) is
 * the text I/O inside the shim: VCF GT/GP parsing and the .plp line reader.
 *
 * All file:line citations are relative to /root/reference.
 */
#ifndef CRAMORE_ORACLE_H
#define CRAMORE_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* Mirrors ccu_demux_result in include/cramore_cuda.h (values consumed by
 * cmd_cram_demuxlet.cpp:921-1013). */
typedef struct {
  int32_t sBest, sNext;
  int32_t dBest1, dBest2, dBestA;
  int32_t dNext1, dNext2, dNextA;
  double sngBestLLK, sngNextLLK;
  double dblBestLLK, dblNextLLK;
  double sumLLK, sngLLK;
} orc_demux_result;

/* Derived call for one droplet (cmd_cram_demuxlet.cpp:921-991). type: 0 SNG, 1 DBL, 2 AMB */
typedef struct {
  int32_t type;
  int32_t jBest, kBest, alphaBest;
  int32_t jNext, kNext, alphaNext;
  double bestLLK, nextLLK;
  double bestPP, sngPP, sngOnlyPP;
} orc_demux_call;

/* PhredHelper.cpp:24-41 : err[q] = q>1 ? pow(0.1, q*0.1) : 0.75 ; mat[q] = 1-err[q] */
void orc_phred_tables(double* err256, double* mat256);

/* sc_drop_seq.cpp:5-8 */
double orc_log_add(double la, double lb);

/* sc_drop_seq.cpp:287-315 (== cmd_cram_demuxlet.cpp:241-268): float GP -> double gps with
 * geno-error mixing.  gp: [nsnps][nv][3] float; snp_err: [nsnps] (already clamped offset+coeff
 * term); has_gp: [nsnps] or NULL; gps_out: [nsnps][nv][3] double (rows of SNPs without GP are
 * left zero). */
void orc_mix_genotypes(int64_t nsnps, int32_t nv, const float* gp, const double* snp_err,
                       const uint8_t* has_gp, double* gps_out);

/* cmd_cram_demuxlet.cpp:655-725 : the nA*9 tile of one (cell,SNP) from its reads. */
void orc_demux_tile(int64_t nreads, const uint8_t* al, const uint8_t* bq, int32_t nA,
                    const double* alpha, double* pG);

/* cmd_cram_demuxlet.cpp:636-906 for cells [cell_begin, cell_end).
 * CSR: cell_ptr[nbcs+1] -> entries; snp_idx[nnz]; read_ptr[nnz+1] -> reads; al/bq[R].
 * gps: [nsnps][nv][3] double (post-mixing); has_gp [nsnps] or NULL.
 * out: [cell_end-cell_begin]; llk_out: optional [cells][nv][nv][nA] copy of llksAB;
 * tile_out: optional [entries of those cells][nA*9].
 * returns 0, or -1 on bad arguments (nv<2 or nA<2 or alpha[0]!=0: SURVEY App. A.7-6). */
int orc_demux_score(int64_t nbcs, const int64_t* cell_ptr, const int32_t* snp_idx,
                    const int64_t* read_ptr, const uint8_t* al, const uint8_t* bq,
                    int64_t nsnps, int32_t nv, const double* gps, const uint8_t* has_gp,
                    int32_t nA, const double* alpha, double doublet_prior,
                    int64_t cell_begin, int64_t cell_end,
                    orc_demux_result* out, double* llk_out, double* tile_out);

/* cmd_cram_demuxlet.cpp:921-991 */
void orc_demux_classify(const orc_demux_result* r, int32_t nv, int32_t nA, const double* alpha,
                        double doublet_prior, orc_demux_call* call);

/* cmd_cram_demuxlet.cpp:993-1013 : one .best row (without trailing '\0' issues); returns length
 * written into buf (snprintf semantics).  sample_ids: nv C strings. */
int orc_demux_best_row(char* buf, int buflen, int32_t int_id, const char* barcode, uint32_t nsnps_cell,
                       int32_t nreads_cell, const orc_demux_result* r, int32_t nv, int32_t nA,
                       const double* alpha, double doublet_prior, const char* const* sample_ids);

/* ------------------------------------------------------------------ freemuxlet */

/* snp_droplet_pileup (sc_drop_seq.h:65-75) flattened */
typedef struct {
  int32_t nreads, nref, nalt, _pad;
  double gls[9];
  double logdenom;
} orc_pileup;

/* sc_drop_seq.cpp:452-509 at fixed alpha (freemuxlet uses 0.5, cmd_cram_freemuxlet.cpp:133) */
void orc_fmx_pileup(int64_t nnz, const int64_t* read_ptr, const uint8_t* al, const uint8_t* bq,
                    double alpha, orc_pileup* out);

/* cmd_cram_freemuxlet.cpp:116-164 : per-cell llk0 / llk2 (the .lmix scores) */
void orc_fmx_lmix(int64_t nbcs, const int64_t* cell_ptr, const int32_t* snp_idx,
                  const orc_pileup* plp, const double* af, double* llk0, double* llk2);

/* sc_drop_seq.h:77-101 : dst.merge(src) */
void orc_fmx_merge(orc_pileup* dst, const orc_pileup* src);

/* cmd_cram_freemuxlet.cpp:359-370 / 581-650 M-step: rebuild clust[K][nsnps] (dense; untouched
 * entries stay default gls==1) from cells with assign[c] >= 0, ascending cell id. */
void orc_fmx_mstep(int64_t nbcs, const int64_t* cell_ptr, const int32_t* snp_idx,
                   const orc_pileup* plp, const int32_t* assign, int32_t K, int64_t nsnps,
                   orc_pileup* clust);

typedef struct {
  int32_t sBest, sNext, dBest1, dBest2, dNext1, dNext2;
  int32_t type;          /* 0 SNG, 1 DBL, 2 AMB */
  int32_t jBest, kBest, jNext, kNext, _pad;
  double sngBestLLK, sngNextLLK, dblBestLLK, dblNextLLK;
  double sumLLK, sngLLK;
  double bestLLK, nextLLK, bestPP, sngPP, sngOnlyPP;
} orc_fmx_result;

/* cmd_cram_freemuxlet.cpp:465-650 E-step + per-cell decision for one iteration.
 * apply_geno_error != 0 applies the geno_error mixing (:485-489, :501-505).
 * llk_out optional [nbcs][K*(K+1)/2]. */
void orc_fmx_estep(int64_t nbcs, const int64_t* cell_ptr, const int32_t* snp_idx,
                   const orc_pileup* plp, const double* af, int32_t K, int64_t nsnps,
                   const orc_pileup* clust, double doublet_prior, double geno_error,
                   int32_t apply_geno_error, orc_fmx_result* out, double* llk_out);

/* dropD (sc_drop_seq.h:45-52); dense lower-triangular matrix of cmd_cram_freemuxlet.cpp:172-221 as
 * out[i * nbcs + j], j < i (the other entries stay zero): nsnps, nread1 (droplet i), nread2 (droplet j),
 * llk0, llk2.  Small nbcs only -- this is the reference's own quadratic layout. */
typedef struct {
  int32_t nsnps, nread1, nread2, _pad;
  double llk0, llk2;
} orc_dropd;
void orc_fmx_pair_dist(int64_t nbcs, int64_t nsnps, const int64_t* cell_ptr, const int32_t* snp_idx,
                       const orc_pileup* plp, const double* af, orc_dropd* out);

#ifdef __cplusplus
}
#endif
#endif
