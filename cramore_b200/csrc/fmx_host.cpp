// fmx_host.cpp -- host half of freemuxlet's initial clustering (cmd_cram_freemuxlet.cpp:166-346) on the
// sparse pair list that ccu_fmx_pair_distances produces.  No GPU, no handle.
//
// The reference walks a dense nbcs x nbcs matrix; a pair of droplets without a common SNP has llk0 = llk2 = 0
// there and casts no vote, so walking only the recorded pairs gives the same votes.  What has to be kept is the
// ORDER in which the +-1 votes are added to the (tiny random) starting values, because the sums are doubles:
// ascending position in the score order for the greedy pass (:262), ascending droplet id for the sweeps (:311).
// Random numbers are libc rand() in the reference's call order: nclust per visited droplet, and the
// libstdc++ std::random_shuffle of every sweep (rand() % (i + 1) for i = 1 .. n - 1, swap when different).
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <numeric>
#include <vector>

#include "cramore_cuda.h"

namespace {

struct Nbr {
  int32_t other;
  double bf;  // llk2 - llk0
};

}  // namespace

extern "C" int ccu_fmx_initial_clusters(int64_t nbcs, int32_t nclust, const double* cell_scores, int64_t npairs,
                                        const ccu_pair_dist* pairs, double bf_thres, double frac_init_clust,
                                        int32_t iter_init, int32_t keep_init_missing, int32_t have_init,
                                        int32_t seed, int32_t* clusts) {
  if (nbcs < 0 || nclust < 1 || npairs < 0 || (nbcs > 0 && (!cell_scores || !clusts)) || (npairs > 0 && !pairs)) return 1;
  for (int64_t p = 0; p < npairs; ++p)
    if (pairs[p].id1 <= pairs[p].id2 || pairs[p].id2 < 0 || pairs[p].id1 >= nbcs) return 1;
  if (seed >= 0) srand((unsigned)seed);
  const int64_t n = nbcs;

  // symmetric adjacency, every row in ascending id of the other droplet (pairs are sorted by (id1, id2): a
  // counting pass keeps that order on both sides)
  // (a pair whose Bayes factor passes the threshold in neither direction never touches a vote: left out)
  auto votes_at_all = [&](const ccu_pair_dist& p) {
    const double bf = p.llk2 - p.llk0;
    return (bf > bf_thres) || (-bf > bf_thres);
  };
  std::vector<int64_t> ptr(n + 1, 0);
  for (int64_t p = 0; p < npairs; ++p)
    if (votes_at_all(pairs[p])) {
      ++ptr[pairs[p].id1 + 1];
      ++ptr[pairs[p].id2 + 1];
    }
  for (int64_t i = 0; i < n; ++i) ptr[i + 1] += ptr[i];
  std::vector<Nbr> adj(ptr[n]);
  {
    std::vector<int64_t> fill(ptr.begin(), ptr.end() - 1);
    // first the smaller ids (id2 of row id1), then the larger ones: two passes keep each row ascending
    for (int64_t p = 0; p < npairs; ++p)
      if (votes_at_all(pairs[p])) adj[fill[pairs[p].id1]++] = {pairs[p].id2, pairs[p].llk2 - pairs[p].llk0};
    for (int64_t p = 0; p < npairs; ++p)
      if (votes_at_all(pairs[p])) adj[fill[pairs[p].id2]++] = {pairs[p].id1, pairs[p].llk2 - pairs[p].llk0};
  }
  for (int64_t i = 0; i < n; ++i)
    if (!std::is_sorted(adj.begin() + ptr[i], adj.begin() + ptr[i + 1], [](const Nbr& a, const Nbr& b) { return a.other < b.other; }))
      std::sort(adj.begin() + ptr[i], adj.begin() + ptr[i + 1], [](const Nbr& a, const Nbr& b) { return a.other < b.other; });

  std::vector<double> votes(nclust);
  auto elect = [&]() {  // :279-287, :322-330: first maximum
    int32_t elected = 0;
    double maxvote = votes[0];
    for (int32_t j = 1; j < nclust; ++j)
      if (maxvote < votes[j]) {
        elected = j;
        maxvote = votes[j];
      }
    return elected;
  };

  if (!have_init) {
    // droplets by singlet score, larger id first on ties (sc_drop_comp_t)
    std::vector<int32_t> order(n);
    std::iota(order.begin(), order.end(), 0);
    std::sort(order.begin(), order.end(), [&](int32_t a, int32_t b) {
      const double cmp = cell_scores[a] - cell_scores[b];
      if (cmp != 0) return cmp > 0;
      return a > b;
    });
    std::vector<int32_t> rank(n);
    for (int64_t i = 0; i < n; ++i) rank[order[i]] = (int32_t)i;
    for (int64_t i = 0; i < n; ++i) clusts[i] = -1;
    std::vector<std::pair<int32_t, double>> earlier;  // (rank, bf) of the droplets voted on
    for (int64_t i = 0; i < n; ++i) {
      if ((double)i > (double)n * frac_init_clust) continue;  // :249
      const int32_t si = order[i];
      for (int32_t j = 0; j < nclust; ++j) votes[j] = rand() / (RAND_MAX + 1.) / 1000.;  // :258-260
      earlier.clear();
      for (int64_t t = ptr[si]; t < ptr[si + 1]; ++t)
        if (rank[adj[t].other] < (int32_t)i) earlier.emplace_back(rank[adj[t].other], adj[t].bf);
      std::sort(earlier.begin(), earlier.end(), [](const std::pair<int32_t, double>& a, const std::pair<int32_t, double>& b) { return a.first < b.first; });
      for (const auto& e : earlier) {  // :262-278 in the order j = 0 .. i-1
        const int32_t c = clusts[order[e.first]];
        if (-e.second > bf_thres) votes[c] -= 1.0;
        else if (e.second > bf_thres) votes[c] += 1.0;
      }
      clusts[si] = elect();
    }
  }

  if (iter_init > 0) {
    std::vector<int32_t> orand(n);
    for (int iter = 0; iter < 10; ++iter) {  // :296 (the count is a literal 10)
      std::iota(orand.begin(), orand.end(), 0);
      for (int64_t i = 1; i < n; ++i) {  // std::random_shuffle(orand.begin(), orand.end()) of libstdc++
        const int64_t j = rand() % (i + 1);
        if (i != j) std::swap(orand[i], orand[j]);
      }
      for (int64_t i = 0; i < n; ++i) {
        const int32_t si = orand[i];
        for (int32_t j = 0; j < nclust; ++j) votes[j] = rand() / (RAND_MAX + 1.) / 1000.;
        for (int64_t t = ptr[si]; t < ptr[si + 1]; ++t) {  // :311-320, j ascending
          const int32_t c = clusts[adj[t].other];
          if (c >= 0) {
            if (adj[t].bf > bf_thres) ++votes[c];
            else if (adj[t].bf < 0 - bf_thres) --votes[c];
          }
        }
        const int32_t elected = elect();
        if (clusts[si] >= 0 || !keep_init_missing) clusts[si] = elected;  // :332-336
      }
    }
  }
  return 0;
}

// ------------------------------------------------------------------------------------------------------------
// freemux2's greedy initialisation (cmd_cram_freemux2.cpp:217-261): droplets in the order of their singlet
// scores; each goes to the cluster whose running pileup it fits best (largest llk2 - llk0 of
// sc_dropseq_lib_t::calculate_droplet_clust_distance, sc_drop_seq.cpp:544-578, first maximum; a cluster without
// any of the droplet's SNPs scores 0) and is then merged into that cluster's pileup
// (snp_droplet_pileup::merge, sc_drop_seq.h:77-101).  Strictly sequential and O(entries x nclust): host code.
// ------------------------------------------------------------------------------------------------------------
extern "C" int ccu_fmx_greedy_clusters(int64_t nbcs, int64_t nsnps, int32_t nclust, const int64_t* cell_ptr,
                                       const int32_t* snp_idx, const ccu_pileup* plp, const double* af,
                                       const double* cell_scores, double frac_init_clust, double singlet_score_thres,
                                       int32_t* clusts) {
  if (nbcs < 0 || nsnps < 0 || nclust < 1 || !cell_ptr || !clusts || (nbcs > 0 && !cell_scores)) return 1;
  const int64_t n = nbcs;
  if (n > 0 && cell_ptr[n] > 0 && (!snp_idx || !plp || !af)) return 1;
  std::vector<int32_t> order(n);
  std::iota(order.begin(), order.end(), 0);
  std::sort(order.begin(), order.end(), [&](int32_t a, int32_t b) {  // sc_drop_comp_t
    const double cmp = cell_scores[a] - cell_scores[b];
    if (cmp != 0) return cmp > 0;
    return a > b;
  });
  // cluster pileups: gls[9] per (cluster, SNP), allocated on first use (a std::map entry of the reference)
  std::vector<double> cgl((size_t)nclust * nsnps * 9);
  std::vector<uint8_t> present((size_t)nclust * nsnps, 0);
  for (int64_t i = 0; i < n; ++i) clusts[i] = -1;
  for (int64_t i = 0; i < n; ++i) {
    const int32_t si = order[i];
    if ((double)i > (double)n * frac_init_clust) continue;    // :225
    if (cell_scores[si] < singlet_score_thres) continue;      // :226
    int32_t maxClust = 0;
    double maxScore = 0;
    for (int32_t j = 0; j < nclust; ++j) {
      double llk0 = 0, llk2 = 0;
      for (int64_t e = cell_ptr[si]; e < cell_ptr[si + 1]; ++e) {
        const int64_t v = snp_idx[e];
        if (!present[(size_t)j * nsnps + v]) continue;
        const double a = af[v];
        double lk0 = 0, lk2 = 0;
        double gps[3];
        gps[0] = (1.0 - a) * (1.0 - a);
        gps[1] = 2.0 * a * (1.0 - a);
        gps[2] = a * a;
        const double* glis = plp[e].gls;
        const double* gljs = &cgl[((size_t)j * nsnps + v) * 9];
        for (int gi = 0; gi < 3; ++gi) {
          lk2 += (glis[gi * 3 + gi] * gljs[gi * 3 + gi] * gps[gi]);
          for (int gj = 0; gj < 3; ++gj) lk0 += (glis[gi * 3 + gi] * gljs[gj * 3 + gj] * gps[gi] * gps[gj]);
        }
        llk2 += std::log(lk2);
        llk0 += std::log(lk0);
      }
      const double score = llk2 - llk0;
      if (j == 0 || score > maxScore) {  // :236-243: seeded with cluster 0, then strict '>'
        maxClust = j;
        maxScore = score;
      }
    }
    clusts[si] = maxClust;
    for (int64_t e = cell_ptr[si]; e < cell_ptr[si + 1]; ++e) {  // :249-252
      const size_t at = (size_t)maxClust * nsnps + snp_idx[e];
      double* gls = &cgl[at * 9];
      if (!present[at]) {
        present[at] = 1;
        for (int t = 0; t < 9; ++t) gls[t] = 1.0;
      }
      for (int t = 0; t < 9; ++t) gls[t] *= plp[e].gls[t];
      double tmp = 0;
      for (int t = 0; t < 9; ++t) tmp += gls[t];
      for (int t = 0; t < 9; ++t) gls[t] /= tmp;
      for (int t = 0; t < 9; ++t)
        if (gls[t] < 1e-6) gls[t] = 1e-6;  // MIN_NORM_GL
      tmp = 0;
      for (int t = 0; t < 9; ++t) tmp += gls[t];
      for (int t = 0; t < 9; ++t) gls[t] /= tmp;
    }
  }
  return 0;
}
