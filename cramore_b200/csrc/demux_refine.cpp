// demux_refine.cpp -- text-level drop-in for `.best`: re-scoring of the REPORTED candidates of each droplet in
// the reference's own operation order (cmd_cram_demuxlet.cpp:655-747 per-read tile loop, `(g_j[l]*g_k[m])*pG`
// l-outer / m-inner, sum of the host libm's log in SNP order), followed by the reference's own top-2 update rule
// (:827-837, :883-906) over those candidates.
//
// Why it exists.  The engine's likelihoods equal the reference's to ~1e-13 relative, which is not enough to
// reproduce two things the reference prints: (1) which of the two orientations (j,k,alpha) / (k,j,1-alpha) of one
// and the same doublet model -- at alpha = 0.5: (j,k) / (k,j) -- comes out "best" and which "next": the reference
// evaluates both and lets the last-bit rounding of its summation order decide; (2) the last
// printed digit of an LLK that sits on a rounding boundary of "%.2lf".  Both are functions of the exact IEEE
// sequence the reference executes, so that sequence is run here -- on the host, with the host libm, because the
// reference's numbers are sums of glibc logs -- for the handful of hypotheses a row can name: the two best singlets
// and the two best doublets the engine found, plus the mirror orientations of those doublets.  The engine has
// already done the search over all nv + nv*(nv-1)*(nA-1) hypotheses; this only decides the print.
// Cost: O(entries of the droplet x (9*nA + candidates)) -- independent of nv*nv.  Threads over droplets.
//
// Not a CPU implementation of the engine: it cannot find a best guess, it takes the engine's and re-evaluates it.
#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <thread>
#include <vector>

#include "cramore_cuda.h"

namespace {

struct GenoSource {
  int32_t nv = 0;
  int64_t nsnps = 0;
  // float form (BCFFilteredReader::get_posterior_at + the mixing of sc_drop_seq.cpp:287-315)
  const float* gp = nullptr;
  const uint8_t* has_gp = nullptr;
  std::vector<double> avg;  // [nsnps][4]: avg0, avg1, avg2, clamped err
  // mixed form (sc_snp_t::gps)
  const double* const* gps = nullptr;

  bool present(int64_t s) const { return gps ? gps[s] != nullptr : (has_gp == nullptr || has_gp[s] != 0); }
  // gps[j*3 + l] of SNP s, exactly as the reference holds it
  void row(int64_t s, int32_t j, double* out) const {
    if (gps) {
      const double* g = gps[s] + (size_t)j * 3;
      out[0] = g[0];
      out[1] = g[1];
      out[2] = g[2];
      return;
    }
    const float* f = gp + ((size_t)s * nv + j) * 3;
    const double* a = &avg[(size_t)s * 4];
    const double err = a[3];
    for (int l = 0; l < 3; ++l) {
      double v = (double)f[l];
      if (err > 0) v = (1 - err) * v + err * a[l];  // sc_drop_seq.cpp:311-315
      out[l] = v;
    }
  }
};

// sc_drop_seq.cpp:288-309 for SNPs [b, e)
void fill_avg(GenoSource& g, const double* snp_err, int64_t b, int64_t e) {
  for (int64_t s = b; s < e; ++s) {
    double* o = &g.avg[(size_t)s * 4];
    o[0] = o[1] = o[2] = 1e-10;
    o[3] = 0;
    if (!g.present(s)) continue;
    const float* f = g.gp + (size_t)s * g.nv * 3;
    double a[3] = {1e-10, 1e-10, 1e-10};
    for (int32_t i = 0; i < g.nv * 3; ++i) a[i % 3] += (double)f[i];
    const double sum = a[0] + a[1] + a[2];
    o[0] = a[0] / sum;
    o[1] = a[1] / sum;
    o[2] = a[2] / sum;
    double err = snp_err[s];
    if (err > 0.999) err = 0.999;
    if (err < 0) err = 0;
    o[3] = err;
  }
}

struct Cand {
  int32_t j, k, n;  // llksAB[j][k][n]
  double llk;
};

struct Refiner {
  int32_t nv, nA;
  const double* alpha;
  const int64_t* cell_ptr;
  const int32_t* snp_idx;
  const int64_t* read_ptr;
  const uint8_t* al;
  const uint8_t* bq;
  const GenoSource* geno;
  double phredErr[256], phredMat[256];
  std::vector<int32_t> mirror_of;  // n -> n' with alpha[n'] == 1 - alpha[n] (n' >= 1), or -1

  void init() {
    for (int q = 0; q < 256; ++q) {  // PhredHelper.cpp:29-33
      phredErr[q] = (q > 1) ? pow(0.1, q * 0.1) : 0.75;
      phredMat[q] = 1. - phredErr[q];
    }
    mirror_of.assign(nA, -1);
    for (int n = 1; n < nA; ++n)
      for (int m = 1; m < nA; ++m)
        if (alpha[m] == 1.0 - alpha[n]) {
          mirror_of[n] = m;
          break;
        }
  }

  // One droplet: every candidate's llksAB entry, the reference's way.
  void score(int64_t c, std::vector<Cand>& cands, std::vector<double>& pGs) const {
    for (Cand& x : cands) x.llk = 0.0;  // memset(llksAB, 0, ...), :643
    // alpha blocks the candidates read (for single-read entries only those are normalised)
    uint64_t need = 0;
    for (const Cand& x : cands) need |= 1ull << x.n;
    pGs.resize((size_t)nA * 9);
    for (int64_t e = cell_ptr[c]; e < cell_ptr[c + 1]; ++e) {
      const int64_t snp = snp_idx[e];
      if (!geno->present(snp)) continue;  // :733
      // ---- tile, :655-725
      int ninf = 0;
      for (int64_t r = read_ptr[e]; r < read_ptr[e + 1]; ++r) ninf += (al[r] != 2);
      std::fill(pGs.begin(), pGs.end(), 1.0);
      if (ninf == 1) {
        // one informative read: the running maximum of :686-687 needs every block, the division of :696 only
        // those that will be read (each value is divided once and never multiplied again)
        int64_t r = read_ptr[e];
        while (al[r] == 2) ++r;
        const uint8_t a = al[r], q = bq[r];
        const double pR = (a == 0) ? phredMat[q] : phredErr[q] / 3.0;
        const double pA = (a == 1) ? phredMat[q] : phredErr[q] / 3.0;
        double maxpG = 0;
        for (int32_t k = 0; k < nA; ++k)
          for (int32_t l = 0; l < 3; ++l)
            for (int32_t m = 0; m < 3; ++m) {
              const double p = 0.5 * l + (m - l) * 0.5 * alpha[k];
              double& pG = pGs[k * 9 + l * 3 + m];
              pG *= (pR * (1.0 - p) + pA * p);
              if (maxpG < pG) maxpG = pG;
            }
        for (int32_t k = 0; k < nA; ++k)
          if ((need >> k) & 1)
            for (int t = 0; t < 9; ++t) pGs[k * 9 + t] /= maxpG;
      } else if (ninf > 1) {
        for (int64_t r = read_ptr[e]; r < read_ptr[e + 1]; ++r) {
          const uint8_t a = al[r], q = bq[r];
          if (a == 2) continue;
          const double pR = (a == 0) ? phredMat[q] : phredErr[q] / 3.0;
          const double pA = (a == 1) ? phredMat[q] : phredErr[q] / 3.0;
          double maxpG = 0;
          for (int32_t k = 0; k < nA; ++k)
            for (int32_t l = 0; l < 3; ++l)
              for (int32_t m = 0; m < 3; ++m) {
                const double p = 0.5 * l + (m - l) * 0.5 * alpha[k];
                double& pG = pGs[k * 9 + l * 3 + m];
                pG *= (pR * (1.0 - p) + pA * p);
                if (maxpG < pG) maxpG = pG;
              }
          for (int32_t t = 0; t < nA * 9; ++t) pGs[t] /= maxpG;
        }
      }
      // :704-725.  After the per-read normalisation the largest value is exactly 1 (it was divided by itself) and
      // every other one is <= 1, and x -> x + 1e-10 is monotone under rounding: the maximum of :711-712 is the
      // constant fl(1 + 1e-10) whatever the blocks hold (also with no informative read: all values are 1).
      const double maxpG2 = 1.0 + 1e-10;
      for (int32_t k = 0; k < nA; ++k)
        if ((need >> k) & 1)
          for (int t = 0; t < 9; ++t) {
            double& pG = pGs[k * 9 + t];
            pG += 1e-10;
            pG /= maxpG2;
          }
      // ---- :738-747 for the candidates
      double gj[3], gk[3];
      int32_t cur_j = -1, cur_k = -1;
      for (Cand& x : cands) {
        if (x.j != cur_j) geno->row(snp, cur_j = x.j, gj);
        if (x.k != cur_k) geno->row(snp, cur_k = x.k, gk);
        const double* pg = &pGs[(size_t)x.n * 9];
        double sumP = 0;
        for (int l = 0; l < 3; ++l)
          for (int m = 0; m < 3; ++m) {
            const double p = gj[l] * gk[m];
            sumP += (p * pg[l * 3 + m]);
          }
        x.llk += log(sumP);
      }
    }
  }

  void droplet(int64_t c, ccu_demux_result& r, std::vector<Cand>& cands, std::vector<double>& pGs) const {
    if (cell_ptr[c + 1] == cell_ptr[c]) return;
    auto valid = [&](int32_t j) { return j >= 0 && j < nv; };
    cands.clear();
    // singlets: llksAB[j][0][0]
    const int s_begin = 0;
    if (valid(r.sBest)) cands.push_back({r.sBest, 0, 0, 0.0});
    if (valid(r.sNext) && r.sNext != r.sBest) cands.push_back({r.sNext, 0, 0, 0.0});
    const int d_begin = (int)cands.size();
    auto add_dbl = [&](int32_t j, int32_t k, int32_t n) {
      if (!valid(j) || !valid(k) || j == k || n < 1 || n >= nA) return;
      for (int i = d_begin; i < (int)cands.size(); ++i)
        if (cands[i].j == j && cands[i].k == k && cands[i].n == n) return;
      cands.push_back({j, k, n, 0.0});
    };
    auto add_with_mirror = [&](int32_t j, int32_t k, int32_t n) {
      add_dbl(j, k, n);
      if (n >= 1 && n < nA && mirror_of[n] >= 0) add_dbl(k, j, mirror_of[n]);
    };
    add_with_mirror(r.dBest1, r.dBest2, r.dBestA);
    add_with_mirror(r.dNext1, r.dNext2, r.dNextA);
    if (cands.empty()) return;
    score(c, cands, pGs);
    // scan order of the reference: singlets by j (:827), doublets by (j, k, n) (:883-885)
    std::sort(cands.begin() + s_begin, cands.begin() + d_begin, [](const Cand& a, const Cand& b) { return a.j < b.j; });
    std::sort(cands.begin() + d_begin, cands.end(), [](const Cand& a, const Cand& b) {
      if (a.j != b.j) return a.j < b.j;
      if (a.k != b.k) return a.k < b.k;
      return a.n < b.n;
    });
    if (d_begin > 0) {
      int32_t sBest = -1, sNext = -1;
      double best = -1e300, next = -1e300;
      for (int i = 0; i < d_begin; ++i) {  // :828-837
        const Cand& x = cands[i];
        if (best < x.llk) {
          next = best;
          sNext = sBest;
          sBest = x.j;
          best = x.llk;
        } else if (next < x.llk) {
          sNext = x.j;
          next = x.llk;
        }
      }
      r.sBest = sBest;
      r.sngBestLLK = best;
      if (d_begin > 1) {
        r.sNext = sNext;
        r.sngNextLLK = next;
      }
    }
    if ((int)cands.size() > d_begin) {
      const Cand* b1 = nullptr;
      const Cand* b2 = nullptr;
      double best = -1e300, next = -1e300;
      for (int i = d_begin; i < (int)cands.size(); ++i) {  // :886-903
        const Cand& x = cands[i];
        if (best < x.llk) {
          b2 = b1;
          next = best;
          b1 = &x;
          best = x.llk;
        } else if (next < x.llk) {
          b2 = &x;
          next = x.llk;
        }
      }
      if (b1) {
        r.dBest1 = b1->j;
        r.dBest2 = b1->k;
        r.dBestA = b1->n;
        r.dblBestLLK = best;
      }
      if (b2) {
        r.dNext1 = b2->j;
        r.dNext2 = b2->k;
        r.dNextA = b2->n;
        r.dblNextLLK = next;
      }
    }
  }
};

int refine_impl(int64_t nbcs, const int64_t* cell_ptr, const int32_t* snp_idx, const int64_t* read_ptr, const uint8_t* al,
                const uint8_t* bq, int32_t nv, int32_t nalpha, const double* alpha, GenoSource& geno, const double* snp_err,
                int32_t nthreads, ccu_demux_result* inout) {
  if (nbcs < 0 || !cell_ptr || nv < 2 || nalpha < 2 || nalpha > CCU_MAX_ALPHA || !alpha) return 1;
  if (nbcs == 0) return 0;
  if (!inout) return 1;
  const int64_t nnz = cell_ptr[nbcs];
  if (nnz > 0 && (!snp_idx || !read_ptr)) return 1;
  if (nnz > 0 && read_ptr[nnz] > 0 && (!al || !bq)) return 1;
  for (int64_t e = 0; e < nnz; ++e)
    if (snp_idx[e] < 0 || snp_idx[e] >= geno.nsnps) return 1;
  int nt = nthreads > 0 ? nthreads : (int)std::thread::hardware_concurrency();
  nt = std::max(1, std::min(nt, 256));
  auto parallel = [&](int64_t n, int64_t grain, auto&& fn) {
    std::atomic<int64_t> next(0);
    auto body = [&]() {
      for (;;) {
        const int64_t b = next.fetch_add(grain);
        if (b >= n) return;
        fn(b, std::min(n, b + grain));
      }
    };
    const int use = (int)std::max<int64_t>(1, std::min<int64_t>(nt, (n + grain - 1) / grain));
    if (use == 1) {
      body();
      return;
    }
    std::vector<std::thread> th;
    for (int i = 0; i < use; ++i) th.emplace_back(body);
    for (auto& t : th) t.join();
  };
  if (!geno.gps) {
    geno.avg.resize((size_t)geno.nsnps * 4);
    parallel(geno.nsnps, 2048, [&](int64_t b, int64_t e) { fill_avg(geno, snp_err, b, e); });
  }
  Refiner R;
  R.nv = nv;
  R.nA = nalpha;
  R.alpha = alpha;
  R.cell_ptr = cell_ptr;
  R.snp_idx = snp_idx;
  R.read_ptr = read_ptr;
  R.al = al;
  R.bq = bq;
  R.geno = &geno;
  R.init();
  parallel(nbcs, 16, [&](int64_t b, int64_t e) {
    std::vector<Cand> cands;
    std::vector<double> pGs;
    for (int64_t c = b; c < e; ++c) R.droplet(c, inout[c], cands, pGs);
  });
  return 0;
}

}  // namespace

extern "C" {

int ccu_demux_refine(int64_t nbcs, const int64_t* cell_ptr, const int32_t* snp_idx, const int64_t* read_ptr,
                     const uint8_t* al, const uint8_t* bq, int32_t nv, int32_t nalpha, const double* alpha, int64_t nsnps,
                     const float* gp, const double* snp_err, const uint8_t* has_gp, int32_t nthreads,
                     ccu_demux_result* inout) {
  if (nsnps < 0 || (nsnps > 0 && (!gp || !snp_err))) return 1;
  GenoSource g;
  g.nv = nv;
  g.nsnps = nsnps;
  g.gp = gp;
  g.has_gp = has_gp;
  return refine_impl(nbcs, cell_ptr, snp_idx, read_ptr, al, bq, nv, nalpha, alpha, g, snp_err, nthreads, inout);
}

int ccu_demux_refine_mixed(int64_t nbcs, const int64_t* cell_ptr, const int32_t* snp_idx, const int64_t* read_ptr,
                           const uint8_t* al, const uint8_t* bq, int32_t nv, int32_t nalpha, const double* alpha,
                           int64_t nsnps, const double* const* gps, int32_t nthreads, ccu_demux_result* inout) {
  if (nsnps < 0 || (nsnps > 0 && !gps)) return 1;
  GenoSource g;
  g.nv = nv;
  g.nsnps = nsnps;
  g.gps = gps;
  return refine_impl(nbcs, cell_ptr, snp_idx, read_ptr, al, bq, nv, nalpha, alpha, g, nullptr, nthreads, inout);
}

}  // extern "C"
