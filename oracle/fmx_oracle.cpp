/*
 * fmx_oracle.cpp -- CPU oracle for the freemuxlet pileup / E-step / M-step path.
 *
 * TEST INFRASTRUCTURE ONLY (see oracle.h).  Restates on flat arrays:
 *     sc_drop_seq.cpp:452-509           calculate_snp_droplet_pileup
 *     sc_drop_seq.h:77-101              snp_droplet_pileup::merge
 *     cmd_cram_freemuxlet.cpp:116-164   per-cell .lmix scores
 *     cmd_cram_freemuxlet.cpp:359-370   initial cluster pileups (M-step form)
 *     cmd_cram_freemuxlet.cpp:456-653   E-step, per-cell decision, M-step rebuild
 * The reference keeps cluster pileups in std::map and default-constructs
 * missing entries (gls == 1, sc_drop_seq.h:72-75); here the K x nsnps array is
 * dense and starts in that default state, which is equivalent.
 * PARITY: pinned against oracle/_ref/ref_cramore (the reference's own sources compiled behind the
 * oracle/ref_shim I/O stand-ins) through tests/golden/ref_*.json -- see oracle.h.
 */
#include "oracle.h"

#include <cmath>
#include <cstring>
#include <vector>

#define ORC_MIN_NORM_GL 1e-6  // sc_drop_seq.h:14

extern "C" {

// sc_drop_seq.cpp:452-509
void orc_fmx_pileup(int64_t nnz, const int64_t* read_ptr, const uint8_t* al, const uint8_t* bq,
                    double alpha, orc_pileup* out) {
  double err[256], mat[256];
  orc_phred_tables(err, mat);
  for (int64_t e = 0; e < nnz; ++e) {
    orc_pileup& p = out[e];
    p.nreads = p.nref = p.nalt = 0;
    p._pad = 0;
    for (int i = 0; i < 9; ++i) p.gls[i] = 1.0;
    p.logdenom = 0;
    double* gls = p.gls;
    double tmp;
    for (int64_t r = read_ptr[e]; r < read_ptr[e + 1]; ++r) {
      uint8_t a = al[r], q = bq[r];
      ++p.nreads;        // :465
      if (a > 1) continue;  // :467
      if (a == 0) ++p.nref;
      else ++p.nalt;
      // :482-490
      gls[0] *= (mat[q] * (a == 0 ? 1.0 : 0.0) + err[q] / 4.);
      gls[1] *= (mat[q] * (a == 0 ? 1. - alpha / 2. : alpha / 2.) + err[q] / 4.);
      gls[2] *= (mat[q] * (a == 0 ? 1.0 - alpha : alpha) + err[q] / 4.);
      gls[3] *= (mat[q] * (a == 0 ? (1. + alpha) / 2. : (1. - alpha) / 2.) + err[q] / 4.);
      gls[4] *= (mat[q] * (a == 0 ? .5 : .5) + err[q] / 4.);
      gls[5] *= (mat[q] * (a == 0 ? (1. - alpha) / 2. : (1. + alpha) / 2.) + err[q] / 4.);
      gls[6] *= (mat[q] * (a == 0 ? alpha : 1. - alpha) + err[q] / 4.);
      gls[7] *= (mat[q] * (a == 0 ? alpha / 2. : 1. - alpha / 2.) + err[q] / 4.);
      gls[8] *= (mat[q] * (a == 0 ? 0.0 : 1.0) + err[q] / 4.);
      tmp = 0;
      for (int i = 0; i < 9; ++i) tmp += gls[i];
      for (int i = 0; i < 9; ++i) gls[i] /= tmp;
      p.logdenom += log(tmp);
    }
    for (int i = 0; i < 9; ++i)
      if (gls[i] < ORC_MIN_NORM_GL) gls[i] = ORC_MIN_NORM_GL;  // :498-501
    tmp = 0;
    for (int i = 0; i < 9; ++i) tmp += gls[i];
    for (int i = 0; i < 9; ++i) gls[i] /= tmp;
    p.logdenom += log(tmp);
  }
}

// cmd_cram_freemuxlet.cpp:131-158
void orc_fmx_lmix(int64_t nbcs, const int64_t* cell_ptr, const int32_t* snp_idx,
                  const orc_pileup* plp, const double* af_arr, double* llk0_out, double* llk2_out) {
  for (int64_t c = 0; c < nbcs; ++c) {
    double llk0 = 0, llk2 = 0;
    for (int64_t e = cell_ptr[c]; e < cell_ptr[c + 1]; ++e) {
      double af = af_arr[snp_idx[e]];
      const double* gls = plp[e].gls;
      double lk0 = 0, lk2 = 0;
      double gps[3];
      gps[0] = (1.0 - af) * (1.0 - af);
      gps[1] = 2.0 * af * (1.0 - af);
      gps[2] = af * af;
      for (int gi = 0; gi < 3; ++gi) {
        lk2 += (gls[gi * 3 + gi] * gps[gi]);
        for (int gj = 0; gj < 3; ++gj) lk0 += (gls[gi * 3 + gj] * gps[gi] * gps[gj]);
      }
      llk0 += log(lk0);
      llk2 += log(lk2);
    }
    llk0_out[c] = llk0;
    llk2_out[c] = llk2;
  }
}

// sc_drop_seq.h:77-101
void orc_fmx_merge(orc_pileup* dst, const orc_pileup* src) {
  dst->nreads += src->nreads;
  dst->nref += src->nref;
  dst->nalt += src->nalt;
  dst->logdenom += src->logdenom;
  double* gls = dst->gls;
  for (int i = 0; i < 9; ++i) gls[i] *= src->gls[i];
  double tmp = 0;
  for (int i = 0; i < 9; ++i) tmp += gls[i];
  dst->logdenom += log(tmp);
  for (int i = 0; i < 9; ++i) gls[i] /= tmp;
  for (int i = 0; i < 9; ++i)
    if (gls[i] < ORC_MIN_NORM_GL) gls[i] = ORC_MIN_NORM_GL;
  tmp = 0;
  for (int i = 0; i < 9; ++i) tmp += gls[i];
  dst->logdenom += log(tmp);
  for (int i = 0; i < 9; ++i) gls[i] /= tmp;
}

// cmd_cram_freemuxlet.cpp:359-370 and the rebuild at :581-650 (ascending cell id)
void orc_fmx_mstep(int64_t nbcs, const int64_t* cell_ptr, const int32_t* snp_idx,
                   const orc_pileup* plp, const int32_t* assign, int32_t K, int64_t nsnps,
                   orc_pileup* clust) {
  for (int64_t i = 0; i < (int64_t)K * nsnps; ++i) {
    orc_pileup& p = clust[i];
    p.nreads = p.nref = p.nalt = p._pad = 0;
    for (int g = 0; g < 9; ++g) p.gls[g] = 1.0;
    p.logdenom = 0;
  }
  for (int64_t c = 0; c < nbcs; ++c) {
    if (assign[c] < 0) continue;
    for (int64_t e = cell_ptr[c]; e < cell_ptr[c + 1]; ++e)
      orc_fmx_merge(&clust[(int64_t)assign[c] * nsnps + snp_idx[e]], &plp[e]);
  }
}

// cmd_cram_freemuxlet.cpp:456-650 (one iteration; the M-step merge is orc_fmx_mstep with
// assign[c] = (type==SNG ? jBest : -1))
void orc_fmx_estep(int64_t nbcs, const int64_t* cell_ptr, const int32_t* snp_idx,
                   const orc_pileup* plp, const double* af_arr, int32_t K, int64_t nsnps,
                   const orc_pileup* clust, double doublet_prior, double geno_error,
                   int32_t apply_geno_error, orc_fmx_result* out, double* llk_out) {
  const int32_t npairs = K * (K + 1) / 2;
  double gp1s[3], gp2s[3], gp0s[3], sum1, sum2;
  double log_single_prior = log((1.0 - doublet_prior) / K);              // :462
  double log_double_prior = log(doublet_prior / K / (K - 1) * 2.0);      // :463
  std::vector<double> llks(npairs), lks(npairs);

  for (int64_t c = 0; c < nbcs; ++c) {
    std::fill(llks.begin(), llks.end(), 0.0);
    for (int64_t e = cell_ptr[c]; e < cell_ptr[c + 1]; ++e) {
      const int64_t s = snp_idx[e];
      double af = af_arr[s];
      gp0s[0] = (1.0 - af) * (1.0 - af);
      gp0s[1] = 2 * af * (1.0 - af);
      gp0s[2] = af * af;
      std::fill(lks.begin(), lks.end(), 0.0);
      double lk;
      const double* glis = plp[e].gls;
      for (int32_t j = 0; j < K; ++j) {
        const orc_pileup& sdp1 = clust[(int64_t)j * nsnps + s];
        gp1s[0] = (1.0 - af) * (1.0 - af) * sdp1.gls[0];
        gp1s[1] = 2 * af * (1.0 - af) * sdp1.gls[4];
        gp1s[2] = af * af * sdp1.gls[8];
        sum1 = gp1s[0] + gp1s[1] + gp1s[2];
        gp1s[0] /= sum1;
        gp1s[1] /= sum1;
        gp1s[2] /= sum1;
        if ((geno_error > 0) && apply_geno_error) {  // :485-489
          gp1s[0] = (1 - geno_error) * gp1s[0] + geno_error * gp0s[0];
          gp1s[1] = (1 - geno_error) * gp1s[1] + geno_error * gp0s[1];
          gp1s[2] = (1 - geno_error) * gp1s[2] + geno_error * gp0s[2];
        }
        for (int32_t k = 0; k < j; ++k) {
          const orc_pileup& sdp2 = clust[(int64_t)k * nsnps + s];
          gp2s[0] = (1.0 - af) * (1.0 - af) * sdp2.gls[0];
          gp2s[1] = 2 * af * (1.0 - af) * sdp2.gls[4];
          gp2s[2] = af * af * sdp2.gls[8];
          sum2 = gp2s[0] + gp2s[1] + gp2s[2];
          gp2s[0] /= sum2;
          gp2s[1] /= sum2;
          gp2s[2] /= sum2;
          if ((geno_error > 0) && apply_geno_error) {
            gp2s[0] = (1 - geno_error) * gp2s[0] + geno_error * gp0s[0];
            gp2s[1] = (1 - geno_error) * gp2s[1] + geno_error * gp0s[1];
            gp2s[2] = (1 - geno_error) * gp2s[2] + geno_error * gp0s[2];
          }
          lk = 0;
          for (int g1 = 0; g1 < 3; ++g1)
            for (int g2 = 0; g2 < 3; ++g2) lk += (glis[g1 * 3 + g2] * gp1s[g1] * gp2s[g2]);  // :511
          lks[j * (j + 1) / 2 + k] = lk;
        }
        lk = 0;
        for (int g1 = 0; g1 < 3; ++g1) lk += (glis[g1 * 3 + g1] * gp1s[g1]);  // :517
        lks[j * (j + 1) / 2 + j] = lk;
      }
      for (int32_t i = 0; i < npairs; ++i) llks[i] += log(lks[i]);  // :521-522
    }

    // :524-579
    orc_fmx_result r;
    memset(&r, 0, sizeof(r));
    r.sBest = r.sNext = r.dBest1 = r.dBest2 = r.dNext1 = r.dNext2 = -1;
    r.sngBestLLK = r.sngNextLLK = r.dblBestLLK = r.dblNextLLK = -1e300;
    double sumLLK = -1e300, sngLLK = -1e300, tmpLLK;
    for (int32_t j = 0; j < K; ++j) {
      for (int32_t k = 0; k < j; ++k) {
        tmpLLK = llks[j * (j + 1) / 2 + k];
        if (tmpLLK > r.dblBestLLK) {
          r.dNext1 = r.dBest1;
          r.dNext2 = r.dBest2;
          r.dblNextLLK = r.dblBestLLK;
          r.dBest1 = j;
          r.dBest2 = k;
          r.dblBestLLK = tmpLLK;
        } else if (tmpLLK > r.dblNextLLK) {
          r.dNext1 = j;
          r.dNext2 = k;
          r.dblNextLLK = tmpLLK;
        }
        sumLLK = orc_log_add(sumLLK, tmpLLK + log_double_prior);
      }
      tmpLLK = llks[j * (j + 1) / 2 + j];
      if (tmpLLK > r.sngBestLLK) {
        r.sNext = r.sBest;
        r.sngNextLLK = r.sngBestLLK;
        r.sBest = j;
        r.sngBestLLK = tmpLLK;
      } else if (tmpLLK > r.sngNextLLK) {
        r.sNext = j;
        r.sngNextLLK = tmpLLK;
      }
      sumLLK = orc_log_add(sumLLK, tmpLLK + log_single_prior);
      sngLLK = orc_log_add(sngLLK, tmpLLK + log_single_prior);
    }
    r.sumLLK = sumLLK;
    r.sngLLK = sngLLK;
    r.sngPP = exp(sngLLK - sumLLK);
    r.sngOnlyPP = exp(r.sngBestLLK + log_single_prior - sngLLK);

    // :584-637
    if (r.dblBestLLK > r.sngBestLLK + 2) {
      r.type = 1;
      r.bestPP = (r.dblBestLLK + log_double_prior - sumLLK);
      r.jBest = r.dBest1;
      r.kBest = r.dBest2;
      r.bestLLK = r.dblBestLLK;
      if (r.dblNextLLK > r.sngBestLLK + 2) {
        r.jNext = r.dNext1;
        r.kNext = r.dNext2;
        r.nextLLK = r.dblNextLLK;
      } else {
        r.jNext = r.kNext = r.sBest;
        r.nextLLK = r.sngBestLLK;
      }
    } else if (r.sngBestLLK > r.sngNextLLK + 2) {
      r.type = 0;
      r.bestPP = (r.sngBestLLK + log_single_prior - sumLLK);
      r.jBest = r.kBest = r.sBest;
      r.bestLLK = r.sngBestLLK;
      if (r.dblBestLLK > r.sngNextLLK + 2) {
        r.jNext = r.dBest1;
        r.kNext = r.dBest2;
        r.nextLLK = r.dblBestLLK;
      } else {
        r.jNext = r.kNext = r.sNext;
        r.nextLLK = r.sngNextLLK;
      }
    } else {
      r.type = 2;
      r.bestPP = (r.sngBestLLK + log_single_prior - sumLLK);
      r.jBest = r.kBest = r.sBest;
      r.bestLLK = r.sngBestLLK;
      if (r.dblBestLLK > r.sngNextLLK + 2) {
        r.jNext = r.dBest1;
        r.kNext = r.dBest2;
        r.nextLLK = r.dblNextLLK;  // :631 (sic: reference uses dblNext here)
      } else {
        r.jNext = r.kNext = r.sNext;
        r.nextLLK = r.sngNextLLK;
      }
    }
    out[c] = r;
    if (llk_out) memcpy(llk_out + c * npairs, llks.data(), sizeof(double) * npairs);
  }
}

}  // extern "C"

// cmd_cram_freemuxlet.cpp:186-221: SNP by SNP, every pair of droplets covering the SNP (the larger id owns the
// row) accumulates log lk0 / log lk2 of the diagonal pileup likelihoods under HWE
void orc_fmx_pair_dist(int64_t nbcs, int64_t nsnps, const int64_t* cell_ptr, const int32_t* snp_idx,
                       const orc_pileup* plp, const double* af_arr, orc_dropd* out) {
  for (int64_t i = 0; i < nbcs * nbcs; ++i) {
    out[i].nsnps = out[i].nread1 = out[i].nread2 = out[i]._pad = 0;
    out[i].llk0 = out[i].llk2 = 0;
  }
  // snp_cell_plps[v]: cells of SNP v in ascending cell id (std::map order)
  std::vector<std::vector<std::pair<int32_t, int64_t>>> by_snp(nsnps);
  for (int64_t c = 0; c < nbcs; ++c)
    for (int64_t e = cell_ptr[c]; e < cell_ptr[c + 1]; ++e) by_snp[snp_idx[e]].push_back({(int32_t)c, e});
  for (int64_t v = 0; v < nsnps; ++v) {
    const auto& cells = by_snp[v];
    for (size_t a = 0; a < cells.size(); ++a) {
      const double* glis = plp[cells[a].second].gls;
      for (size_t b = 0; b < a; ++b) {
        const double* gljs = plp[cells[b].second].gls;
        double af = af_arr[v];
        double lk0 = 0, lk2 = 0;
        double gps[3];
        gps[0] = (1.0 - af) * (1.0 - af);
        gps[1] = 2.0 * af * (1.0 - af);
        gps[2] = af * af;
        for (int gi = 0; gi < 3; ++gi) {
          lk2 += (glis[gi * 3 + gi] * gljs[gi * 3 + gi] * gps[gi]);
          for (int gj = 0; gj < 3; ++gj) lk0 += (glis[gi * 3 + gi] * gljs[gj * 3 + gj] * gps[gi] * gps[gj]);
        }
        orc_dropd& dd = out[(int64_t)cells[a].first * nbcs + cells[b].first];
        ++dd.nsnps;
        dd.nread1 += plp[cells[a].second].nreads;
        dd.nread2 += plp[cells[b].second].nreads;
        dd.llk2 += log(lk2);
        dd.llk0 += log(lk0);
      }
    }
  }
}
