/*
 * demux_oracle.cpp -- CPU oracle for the demuxlet droplet-likelihood engine.
 *
 * TEST INFRASTRUCTURE ONLY (see oracle.h).  Restates, on flat CSR arrays, the
 * single-threaded loop nest of the reference:
 *     cmd_cram_demuxlet.cpp:636-1013  (per-droplet engine + classification + row)
 *     sc_drop_seq.cpp:5-8             (logAdd)
 *     sc_drop_seq.cpp:287-315         (geno-error mixing of float GPs)
 *     PhredHelper.cpp:24-41           (Phred tables)
 * The operation ORDER of every floating-point expression follows the cited
 * lines (this file must be compiled like the reference: -O3, no -march=native,
 * -ffp-contract=off, so no FMA contraction).
 *
 * PARITY: pinned against oracle/_ref/ref_cramore (the reference's own sources compiled behind the
 * oracle/ref_shim I/O stand-ins) through tests/golden/ref_*.json -- see oracle.h.
 */
#include "oracle.h"

#include <cmath>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

namespace {

struct PhredTables {
  double err[256];
  double mat[256];
  PhredTables() {
    // PhredHelper.cpp:29-33
    for (int q = 0; q <= 255; ++q) {
      err[q] = (q > 1) ? pow(0.1, q * 0.1) : 0.75;
      mat[q] = 1. - err[q];
    }
  }
};

const PhredTables& phred() {
  static PhredTables t;
  return t;
}

}  // namespace

extern "C" {

void orc_phred_tables(double* err256, double* mat256) {
  const PhredTables& t = phred();
  for (int q = 0; q < 256; ++q) {
    err256[q] = t.err[q];
    mat256[q] = t.mat[q];
  }
}

// sc_drop_seq.cpp:5-8
double orc_log_add(double la, double lb) {
  if (la > lb) return la + log(1.0 + exp(lb - la));
  return lb + log(1.0 + exp(la - lb));
}

// sc_drop_seq.cpp:287-315
void orc_mix_genotypes(int64_t nsnps, int32_t nv, const float* gp, const double* snp_err,
                       const uint8_t* has_gp, double* gps_out) {
  const int64_t row = (int64_t)nv * 3;
  for (int64_t s = 0; s < nsnps; ++s) {
    double* g = gps_out + s * row;
    if (has_gp != NULL && !has_gp[s]) {
      for (int64_t i = 0; i < row; ++i) g[i] = 0;
      continue;
    }
    const float* f = gp + s * row;
    double avg[3] = {1e-10, 1e-10, 1e-10};  // :289
    for (int64_t i = 0; i < row; ++i) avg[i % 3] += (g[i] = f[i]);  // :291-293
    double sum = avg[0] + avg[1] + avg[2];  // :295
    avg[0] /= sum;
    avg[1] /= sum;
    avg[2] /= sum;
    double err = snp_err[s];
    if (err > 0.999) err = 0.999;  // :308-309
    if (err < 0) err = 0;
    if (err > 0) {  // :311-315
      for (int64_t i = 0; i < row; ++i) g[i] = (1 - err) * g[i] + err * avg[i % 3];
    }
  }
}

// cmd_cram_demuxlet.cpp:655-725
void orc_demux_tile(int64_t nreads, const uint8_t* al, const uint8_t* bq, int32_t nA,
                    const double* alpha, double* pG) {
  const PhredTables& pc = phred();
  const int32_t n9 = nA * 9;
  for (int32_t i = 0; i < n9; ++i) pG[i] = 1.0;  // :657
  for (int64_t r = 0; r < nreads; ++r) {
    if (al[r] == 2) continue;  // :664
    double pR = (al[r] == 0) ? pc.mat[bq[r]] : pc.err[bq[r]] / 3.0;  // :666
    double pA = (al[r] == 1) ? pc.mat[bq[r]] : pc.err[bq[r]] / 3.0;  // :667
    double maxpG = 0;
    for (int32_t k = 0; k < nA; ++k)
      for (int32_t l = 0; l < 3; ++l)
        for (int32_t m = 0; m < 3; ++m) {
          double p = 0.5 * l + (m - l) * 0.5 * alpha[k];  // :673
          double& v = pG[k * 9 + l * 3 + m];
          v *= (pR * (1.0 - p) + pA * p);  // :685
          if (maxpG < v) maxpG = v;
        }
    for (int32_t i = 0; i < n9; ++i) pG[i] /= maxpG;  // :692-699
  }
  double maxpG = 0;  // :704-715
  for (int32_t i = 0; i < n9; ++i) {
    pG[i] += 1e-10;
    if (maxpG < pG[i]) maxpG = pG[i];
  }
  for (int32_t i = 0; i < n9; ++i) pG[i] /= maxpG;  // :718-725
}

int orc_demux_score(int64_t nbcs, const int64_t* cell_ptr, const int32_t* snp_idx,
                    const int64_t* read_ptr, const uint8_t* al, const uint8_t* bq,
                    int64_t nsnps, int32_t nv, const double* gps, const uint8_t* has_gp,
                    int32_t nA, const double* alpha, double doublet_prior,
                    int64_t cell_begin, int64_t cell_end,
                    orc_demux_result* out, double* llk_out, double* tile_out) {
  (void)nsnps;
  if (nv < 2 || nA < 2 || alpha[0] != 0.0) return -1;
  if (cell_begin < 0 || cell_end > nbcs || cell_begin > cell_end) return -1;

  const int64_t nllk = (int64_t)nv * nv * nA;
  std::vector<double> llksAB(nllk);
  std::vector<double> pGs(nA * 9);
  std::vector<double> sumPs(nA);
  int64_t tile_row = 0;

  for (int64_t c = cell_begin; c < cell_end; ++c) {
    std::fill(llksAB.begin(), llksAB.end(), 0.0);  // :643

    for (int64_t e = cell_ptr[c]; e < cell_ptr[c + 1]; ++e) {  // :656
      orc_demux_tile(read_ptr[e + 1] - read_ptr[e], al + read_ptr[e], bq + read_ptr[e], nA, alpha,
                     pGs.data());
      if (tile_out) {
        memcpy(tile_out + tile_row * nA * 9, pGs.data(), sizeof(double) * nA * 9);
        ++tile_row;
      }
      const int32_t isnp = snp_idx[e];
      if (has_gp != NULL && !has_gp[isnp]) continue;  // :733
      const double* g = gps + (int64_t)isnp * nv * 3;
      for (int32_t j = 0; j < nv; ++j) {    // :734
        for (int32_t k = 0; k < nv; ++k) {  // :736
          std::fill(sumPs.begin(), sumPs.end(), 0.0);
          for (int32_t l = 0; l < 3; ++l)
            for (int32_t m = 0; m < 3; ++m) {
              double p = g[j * 3 + l] * g[k * 3 + m];  // :740
              for (int32_t n = 0; n < nA; ++n) sumPs[n] += (p * pGs[n * 9 + l * 3 + m]);  // :742
            }
          double* dst = &llksAB[((int64_t)j * nv + k) * nA];
          for (int32_t n = 0; n < nA; ++n) dst[n] += log(sumPs[n]);  // :746
        }
      }
    }

    // :788-795
    orc_demux_result r;
    r.sBest = r.sNext = r.dBest1 = r.dBest2 = r.dNext1 = r.dNext2 = r.dBestA = r.dNextA = -1;
    r.sngBestLLK = r.sngNextLLK = -1e300;
    r.dblBestLLK = r.dblNextLLK = -1e300;
    double sumLLK = -1e-300, sngLLK = -1e-300;  // :791 (sic)
    double log_single_prior = log((1.0 - doublet_prior) / nv);
    double log_doublet_prior1 = log(doublet_prior / nv / (nv - 1.) / (nA - 1.));
    double log_doublet_prior2 = log(doublet_prior / nv / (nv - 1.) / (nA - 1.) * 2);

    // :804-821
    for (int32_t j = 0; j < nv; ++j) {
      double s = llksAB[(int64_t)j * nv * nA];
      sumLLK = orc_log_add(sumLLK, s + log_single_prior);
      sngLLK = orc_log_add(sngLLK, s + log_single_prior);
      for (int32_t k = 0; k < nv; ++k) {
        if (j == k) continue;
        for (int32_t n = 1; n < nA; ++n) {
          double v = llksAB[((int64_t)j * nv + k) * nA + n];
          if (alpha[n] == 0.5) {
            if (k > j) continue;
            sumLLK = orc_log_add(sumLLK, v + log_doublet_prior2);
          } else {
            sumLLK = orc_log_add(sumLLK, v + log_doublet_prior1);
          }
        }
      }
    }

    // :827-837
    for (int32_t j = 0; j < nv; ++j) {
      double s = llksAB[(int64_t)j * nv * nA];
      if (r.sngBestLLK < s) {
        r.sngNextLLK = r.sngBestLLK;
        r.sNext = r.sBest;
        r.sBest = j;
        r.sngBestLLK = s;
      } else if (r.sngNextLLK < s) {
        r.sNext = j;
        r.sngNextLLK = s;
      }
    }

    // :883-906
    for (int32_t j = 0; j < nv; ++j)
      for (int32_t k = 0; k < nv; ++k) {
        if (j == k) continue;
        for (int32_t n = 1; n < nA; ++n) {
          double v = llksAB[((int64_t)j * nv + k) * nA + n];
          if (r.dblBestLLK < v) {
            r.dNext1 = r.dBest1;
            r.dNext2 = r.dBest2;
            r.dNextA = r.dBestA;
            r.dblNextLLK = r.dblBestLLK;
            r.dBest1 = j;
            r.dBest2 = k;
            r.dBestA = n;
            r.dblBestLLK = v;
          } else if (r.dblNextLLK < v) {
            r.dNext1 = j;
            r.dNext2 = k;
            r.dNextA = n;
            r.dblNextLLK = v;
          }
        }
      }
    r.sumLLK = sumLLK;
    r.sngLLK = sngLLK;
    out[c - cell_begin] = r;
    if (llk_out) memcpy(llk_out + (c - cell_begin) * nllk, llksAB.data(), sizeof(double) * nllk);
  }
  return 0;
}

// cmd_cram_demuxlet.cpp:921-991
void orc_demux_classify(const orc_demux_result* r, int32_t nv, int32_t nA, const double* alpha,
                        double doublet_prior, orc_demux_call* c) {
  double log_single_prior = log((1.0 - doublet_prior) / nv);
  double log_doublet_prior1 = log(doublet_prior / nv / (nv - 1.) / (nA - 1.));
  double log_doublet_prior2 = log(doublet_prior / nv / (nv - 1.) / (nA - 1.) * 2);
  c->jBest = c->kBest = c->jNext = c->kNext = c->alphaBest = c->alphaNext = -1;
  c->bestLLK = c->nextLLK = -1e300;
  if (r->dblBestLLK > r->sngBestLLK + 2) {  // :925
    c->type = 1;
    c->bestPP = exp(r->dblBestLLK +
                    ((alpha[r->dBestA] == 0.5) ? log_doublet_prior2 : log_doublet_prior1) - r->sumLLK);
    c->jBest = r->dBest1;
    c->kBest = r->dBest2;
    c->bestLLK = r->dblBestLLK;
    c->alphaBest = r->dBestA;
    if (r->dblNextLLK > r->sngBestLLK + 2) {
      c->jNext = r->dNext1;
      c->kNext = r->dNext2;
      c->nextLLK = r->dblNextLLK;
      c->alphaNext = r->dNextA;
    } else {
      c->jNext = c->kNext = r->sBest;
      c->nextLLK = r->sngBestLLK;
      c->alphaNext = 0;
    }
  } else {
    c->type = (r->sngBestLLK > r->sngNextLLK + 2) ? 0 : 2;  // :947 / :968
    c->bestPP = r->sngBestLLK + log_single_prior - r->sumLLK;  // :949 (log space, sic)
    c->jBest = c->kBest = r->sBest;
    c->bestLLK = r->sngBestLLK;
    c->alphaBest = 0;
    if (r->dblBestLLK > r->sngNextLLK + 2) {
      c->jNext = r->dBest1;
      c->kNext = r->dBest2;
      c->nextLLK = r->dblBestLLK;
      c->alphaNext = r->dBestA;
    } else {
      c->jNext = c->kNext = r->sNext;
      c->nextLLK = r->sngNextLLK;
      c->alphaNext = 0;
    }
  }
  c->sngPP = exp(r->sngLLK - r->sumLLK);                               // :990
  c->sngOnlyPP = exp(r->sngBestLLK + log_single_prior - r->sngLLK);  // :991
}

// cmd_cram_demuxlet.cpp:993-1013
int orc_demux_best_row(char* buf, int buflen, int32_t int_id, const char* barcode, uint32_t nsnps_cell,
                       int32_t nreads_cell, const orc_demux_result* r, int32_t nv, int32_t nA,
                       const double* alpha, double doublet_prior, const char* const* sample_ids) {
  orc_demux_call c;
  orc_demux_classify(r, nv, nA, alpha, doublet_prior, &c);
  static const char* names[3] = {"SNG", "DBL", "AMB"};
  return snprintf(
      buf, buflen,
      "%d\t%s\t%u\t%d\t%s\t%s,%s,%.2lf\t%.2lf\t%s,%s,%.2lf\t%.2lf\t%.2lf\t%.2lg\t%.2lg\t%s\t%.2lf\t%s\t%."
      "2lf\t%.5lf\t%s,%s,%.2lf\t%.2lf\t%.2lf\n",
      int_id, barcode, nsnps_cell, nreads_cell, names[c.type], sample_ids[c.jBest], sample_ids[c.kBest],
      alpha[c.alphaBest], c.bestLLK, sample_ids[c.jNext], sample_ids[c.kNext], alpha[c.alphaNext],
      c.nextLLK, c.bestLLK - c.nextLLK, c.bestPP, c.sngPP, sample_ids[r->sBest], r->sngBestLLK,
      sample_ids[r->sNext], r->sngNextLLK, c.sngOnlyPP, sample_ids[r->dBest1], sample_ids[r->dBest2],
      alpha[r->dBestA], r->dblBestLLK, r->sngBestLLK - r->dblBestLLK);
}

}  // extern "C"
