// demux_host.cpp -- host half of the demuxlet seam: the SNG/DBL/AMB decision and the `.best`
// row writer, i.e. the part of cmdCramDemuxlet that stays on the host once the engine has
// produced a ccu_demux_result per droplet (cmd_cram_demuxlet.cpp:921-1013, header :629).
#include <cmath>
#include <cstdio>

#include "cramore_cuda.h"

extern "C" {

const char* ccu_demux_best_header(void) {
  // cmd_cram_demuxlet.cpp:629
  return "INT_ID\tBARCODE\tNUM.SNPS\tNUM.READS\tDROPLET.TYPE\tBEST.GUESS\tBEST.LLK\tNEXT.GUESS\tNEXT.LLK\t"
         "DIFF.LLK.BEST.NEXT\tBEST.POSTERIOR\tSNG.POSTERIOR\tSNG.BEST.GUESS\tSNG.BEST.LLK\tSNG.NEXT.GUESS\t"
         "SNG.NEXT.LLK\tSNG.ONLY.POSTERIOR\tDBL.BEST.GUESS\tDBL.BEST.LLK\tDIFF.LLK.SNG.DBL\n";
}

int ccu_demux_classify(const ccu_demux_result* r, int32_t nv, int32_t nalpha, const double* alpha,
                       double doublet_prior, ccu_demux_call* c) {
  if (!r || !c || !alpha || nv < 2 || nalpha < 2) return 1;
  // priors, cmd_cram_demuxlet.cpp:793-795
  const double lsp = log((1.0 - doublet_prior) / nv);
  const double ldp1 = log(doublet_prior / nv / (nv - 1.) / (nalpha - 1.));
  const double ldp2 = log(doublet_prior / nv / (nv - 1.) / (nalpha - 1.) * 2);
  c->_pad = 0;
  const bool dbl = r->dblBestLLK > r->sngBestLLK + 2;                 // :925
  const bool sng = !dbl && (r->sngBestLLK > r->sngNextLLK + 2);       // :947
  c->type = dbl ? 1 : (sng ? 0 : 2);
  if (dbl) {
    const bool half = r->dBestA >= 0 && alpha[r->dBestA] == 0.5;
    c->bestPP = exp(r->dblBestLLK + (half ? ldp2 : ldp1) - r->sumLLK);  // :927
    c->jBest = r->dBest1;
    c->kBest = r->dBest2;
    c->alphaBest = r->dBestA;
    c->bestLLK = r->dblBestLLK;
    const bool next_dbl = r->dblNextLLK > r->sngBestLLK + 2;  // :933
    c->jNext = next_dbl ? r->dNext1 : r->sBest;
    c->kNext = next_dbl ? r->dNext2 : r->sBest;
    c->alphaNext = next_dbl ? r->dNextA : 0;
    c->nextLLK = next_dbl ? r->dblNextLLK : r->sngBestLLK;
  } else {
    c->bestPP = r->sngBestLLK + lsp - r->sumLLK;  // :949 / :970, printed in log space by the reference
    c->jBest = c->kBest = r->sBest;
    c->alphaBest = 0;
    c->bestLLK = r->sngBestLLK;
    const bool next_dbl = r->dblBestLLK > r->sngNextLLK + 2;  // :954 / :975
    c->jNext = next_dbl ? r->dBest1 : r->sNext;
    c->kNext = next_dbl ? r->dBest2 : r->sNext;
    c->alphaNext = next_dbl ? r->dBestA : 0;
    c->nextLLK = next_dbl ? r->dblBestLLK : r->sngNextLLK;
  }
  c->sngPP = exp(r->sngLLK - r->sumLLK);                 // :990
  c->sngOnlyPP = exp(r->sngBestLLK + lsp - r->sngLLK);  // :991
  return 0;
}

int ccu_demux_best_row(char* buf, int32_t buflen, int32_t int_id, const char* barcode, uint32_t nsnps_cell,
                       int32_t nreads_cell, const ccu_demux_result* r, int32_t nv, int32_t nalpha,
                       const double* alpha, double doublet_prior, const char* const* ids) {
  ccu_demux_call c;
  if (ccu_demux_classify(r, nv, nalpha, alpha, doublet_prior, &c)) return -1;
  const int idx[] = {c.jBest, c.kBest, c.jNext, c.kNext, r->sBest, r->sNext, r->dBest1, r->dBest2};
  for (int i : idx)
    if (i < 0 || i >= nv) return -1;  // the reference would index sample -1 here (undefined)
  if (c.alphaBest < 0 || c.alphaNext < 0 || r->dBestA < 0) return -1;
  static const char* const kType[3] = {"SNG", "DBL", "AMB"};
  // cmd_cram_demuxlet.cpp:993
  const int n = snprintf(buf, buflen,
                  "%d\t%s\t%u\t%d\t%s\t%s,%s,%.2lf\t%.2lf\t%s,%s,%.2lf\t%.2lf\t%.2lf\t%.2lg\t%.2lg\t%s\t%.2lf\t%s\t"
                  "%.2lf\t%.5lf\t%s,%s,%.2lf\t%.2lf\t%.2lf\n",
                  int_id, barcode, nsnps_cell, nreads_cell, kType[c.type], ids[c.jBest], ids[c.kBest],
                  alpha[c.alphaBest], c.bestLLK, ids[c.jNext], ids[c.kNext], alpha[c.alphaNext], c.nextLLK,
                  c.bestLLK - c.nextLLK, c.bestPP, c.sngPP, ids[r->sBest], r->sngBestLLK, ids[r->sNext],
                  r->sngNextLLK, c.sngOnlyPP, ids[r->dBest1], ids[r->dBest2], alpha[r->dBestA], r->dblBestLLK,
                  r->sngBestLLK - r->dblBestLLK);
  return (n < 0 || n >= buflen) ? -2 : n;  // -2: the row does not fit buflen (nothing usable was written)
}

}  // extern "C"
