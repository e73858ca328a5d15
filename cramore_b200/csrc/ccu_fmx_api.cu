// ccu_fmx_api.cu -- freemuxlet half of the C ABI (include/cramore_cuda.h): device memory, the
// SNP-major index build (CUB radix sort, once per upload) and the EM driver around the kernels of
// fmx_kernels.cu.  No CPU implementation exists behind any entry point.
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstring>
#include <cub/device/device_radix_sort.cuh>
#include <thread>
#include <vector>

#include "ccu_handle.h"
#include "fmx_internal.h"

struct FmxPairs;  // fmx_init.cu
void ccu_fmx_pairs_release(FmxPairs*& p);
int ccu_fmx_pairs_build(ccu_handle* h, const ccu::FmxDev& d, FmxPairs*& ps, int64_t* npairs_out);
int ccu_fmx_pairs_download(ccu_handle* h, FmxPairs* ps, ccu_pair_dist* out);
int ccu_fmx_pairs_download_voting(ccu_handle* h, FmxPairs* ps, double bf_thres, int64_t* nvoting, ccu_pair_dist* out);
float ccu_fmx_pairs_ms(const FmxPairs* ps);
int64_t ccu_fmx_pairs_records(const FmxPairs* ps);

struct FmxState {
  bool configured = false, uploaded = false, have_stats = false, have_estep = false, have_post = false;
  // the posterior / mean-genotype arrays: post_full = they hold every SNP (not only this rank's slab), computed with
  // the genotype-error flag post_apply; cgm_stale = the mean genotypes have to be re-derived from the posteriors
  bool post_full = false, cgm_stale = false;
  int post_apply = -1;
  long long timeout_cycles = 0;
  int32_t K = 0;
  double doublet_prior = 0.5, geno_error = 0.0;
  int64_t nbcs = 0, nsnps = 0, nnz = 0, nreads = 0, max_segment = 0;
  int64_t cell_begin = 0, cell_end = 0, snp_begin = 0, snp_end = 0;
  DevBuf cell_ptr, snp_idx, read_ptr, al, bq, af;
  DevBuf gl, cnt, logdenom, csc_ptr, csc_cell, csc_row;
  DevBuf eclass, egen, eab, esnp, gcnt;
  DevBuf assign, stats, cgp, cgm, llk, results, scan, scratch, flags, counter, snp_gen;
  cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};
  ccu_fmx_timings tm = {};
  FmxPairs* pairs = nullptr;
  // peer-memory exchange: every rank's posterior / assignment array (CUDA IPC mappings; [peer_rank] is our own)
  ccu::FmxPeers peers = {};
  int32_t peer_rank = -1;
  bool peers_are_ipc = true;  // false: pointers of other handles of this process (nothing to unmap)
  void close_peers() {
    for (int r = 0; r < peers.n; ++r)
      if (r != peer_rank && peers_are_ipc) {
        if (peers.post[r]) cudaIpcCloseMemHandle(peers.post[r]);
        if (peers.postm[r]) cudaIpcCloseMemHandle(peers.postm[r]);
        if (peers.assign[r]) cudaIpcCloseMemHandle(peers.assign[r]);
        if (peers.flags[r]) cudaIpcCloseMemHandle(peers.flags[r]);
      }
    peers = {};
    peer_rank = -1;
    peers_are_ipc = true;
  }
  void release() {
    close_peers();
    ccu_fmx_pairs_release(pairs);
    DevBuf* all[] = {&cell_ptr, &snp_idx, &read_ptr, &al, &bq, &af, &gl, &cnt, &logdenom, &csc_ptr, &csc_cell,
                     &csc_row, &eclass, &egen, &eab, &esnp, &gcnt, &snp_gen, &assign, &stats, &cgp, &cgm, &llk, &results,
                     &scan, &scratch, &flags, &counter};
    for (DevBuf* b : all) b->release();
    for (auto& e : ev)
      if (e) cudaEventDestroy(e);
  }
};

void ccu_fmx_release(ccu_handle* h) {
  if (h && h->fmx) {
    h->fmx->release();
    delete h->fmx;
    h->fmx = nullptr;
  }
}

namespace {

ccu::FmxDev fmx_view(const ccu_handle* h) {
  const FmxState* f = h->fmx;
  ccu::FmxDev d;
  memset(&d, 0, sizeof(d));
  d.K = f->K;
  d.doublet_prior = f->doublet_prior;
  d.geno_error = f->geno_error;
  d.log_sng_prior = std::log((1.0 - f->doublet_prior) / f->K);
  d.log_dbl_prior = std::log(f->doublet_prior / f->K / (f->K - 1) * 2.0);
  d.nbcs = f->nbcs;
  d.nsnps = f->nsnps;
  d.nnz = f->nnz;
  d.max_segment = f->max_segment;
  d.cell_ptr = f->cell_ptr.as<int64_t>();
  d.snp_idx = f->snp_idx.as<int32_t>();
  d.read_ptr = f->read_ptr.as<int64_t>();
  d.al = f->al.as<uint8_t>();
  d.bq = f->bq.as<uint8_t>();
  d.af = f->af.as<double>();
  d.phred = h->d_phred.as<double>();
  d.gl = f->gl.as<double>();
  d.cnt = f->cnt.as<int4>();
  d.logdenom = f->logdenom.as<double>();
  d.csc_ptr = f->csc_ptr.as<int64_t>();
  d.csc_cell = f->csc_cell.as<int32_t>();
  d.csc_row = f->csc_row.as<double>();
  d.assign = f->assign.as<int32_t>();
  d.stats = f->stats.as<double>();
  d.snp_gen = f->snp_gen.as<uint8_t>();
  d.cgp = f->cgp.as<double>();
  d.cgm = f->cgm.as<double>();
  d.eclass = f->eclass.as<uint8_t>();
  d.egen = f->egen.as<int32_t>();
  d.eab = f->eab.as<double2>();
  d.esnp = f->esnp.as<int32_t>();
  d.gcnt = f->gcnt.as<int32_t>();
  d.llk = nullptr;  // the pair log-likelihood matrix is written only for ccu_fmx_get_llk (which runs the E-step again)
  d.results = f->results.as<ccu_fmx_result>();
  d.scan = f->scan.as<ccu::FmxScan>();
  d.cell_begin = f->cell_begin;
  d.cell_end = f->cell_end;
  d.snp_begin = f->snp_begin;
  d.snp_end = f->snp_end;
  return d;
}

struct Timer {  // CUDA-event timer on the engine stream
  ccu_handle* h;
  float* dst;
  Timer(ccu_handle* h_, float* dst_) : h(h_), dst(dst_) { cudaEventRecord(h->fmx->ev[0], h->stream); }
  void stop_sync() {
    cudaEventRecord(h->fmx->ev[1], h->stream);
    cudaEventSynchronize(h->fmx->ev[1]);
    float ms = 0;
    if (cudaEventElapsedTime(&ms, h->fmx->ev[0], h->fmx->ev[1]) == cudaSuccess) *dst = ms;
  }
};

int need(ccu_handle* h, bool uploaded, const char* who) {
  if (!h) return 1;
  if (!h->fmx || !h->fmx->configured) return ccu_fail(h, "%s: call ccu_fmx_configure first", who);
  if (uploaded && !h->fmx->uploaded) return ccu_fail(h, "%s: call ccu_fmx_upload first", who);
  return 0;
}

// How the kernels of the loop take part in the cross-rank ordering (FmxSync): in-kernel when asked for and peers are
// mapped, otherwise not at all (single GPU, or the caller orders the ranks itself).
ccu::FmxSync make_sync(const ccu_handle* h, bool in_kernel) {
  const FmxState* f = h->fmx;
  ccu::FmxSync s;
  memset(&s, 0, sizeof(s));
  const bool on = in_kernel && f->peers.n > 1;
  s.wait = on ? 1 : 0;
  s.arrive = on ? 1 : 0;
  s.rank = f->peer_rank < 0 ? 0 : f->peer_rank;
  s.timeout_cycles = f->timeout_cycles;
  s.counter = f->counter.as<unsigned int>();
  return s;
}

ccu::FmxPeers no_peers() {
  ccu::FmxPeers p;
  memset(&p, 0, sizeof(p));
  return p;
}

int reserve_shard_buffers(ccu_handle* h) {
  FmxState* f = h->fmx;
  const int64_t ncell = f->cell_end - f->cell_begin;
  const int64_t npairs = (int64_t)f->K * (f->K + 1) / 2;
  (void)npairs;  // llk[ncell][npairs] is allocated by ccu_fmx_get_llk, the only reader
  CCU_CUDA(h, f->results.reserve((size_t)ncell * sizeof(ccu_fmx_result) + 16));
  CCU_CUDA(h, f->scan.reserve((size_t)ncell * sizeof(ccu::FmxScan) + 16));
  return 0;
}

}  // namespace

extern "C" {

int ccu_fmx_configure(ccu_handle* h, int32_t nclust, double doublet_prior, double geno_error) {
  if (!h) return 1;
  if (nclust < 2 || nclust > ccu::kFmxMaxClust)
    return ccu_fail(h, "ccu_fmx_configure: --nsample %d outside [2,%d]", nclust, ccu::kFmxMaxClust);
  CCU_CUDA(h, cudaSetDevice(h->device));
  if (!h->fmx) {
    h->fmx = new FmxState();
    for (auto& e : h->fmx->ev) cudaEventCreate(&e);
    // how long a kernel of the multi-GPU loop waits for its peers before it raises the error flag instead of
    // hanging the GPU: CCU_FMX_BARRIER_TIMEOUT_S seconds (default 30 -- a late peer_open, a lazy module load or
    // host-thread contention on a healthy run must not trip it)
    int khz = 1900000;
    cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, h->device);
    double secs = 30.0;
    if (const char* v = getenv("CCU_FMX_BARRIER_TIMEOUT_S")) secs = atof(v) > 0 ? atof(v) : secs;
    h->fmx->timeout_cycles = (long long)(secs * 1e3 * (double)khz);
  }
  FmxState* f = h->fmx;
  f->K = nclust;
  f->doublet_prior = doublet_prior;
  f->geno_error = geno_error;
  f->configured = true;
  f->uploaded = f->have_stats = f->have_estep = f->have_post = f->post_full = false;
  f->post_apply = -1;
  return 0;
}

int ccu_fmx_upload(ccu_handle* h, int64_t nbcs, int64_t nsnps, const int64_t* cell_ptr, const int32_t* snp_idx,
                   const int64_t* read_ptr, const uint8_t* al, const uint8_t* bq, const double* af) {
  CcuRange nvtx_range("ccu_fmx_upload");
  if (int rc = need(h, false, "ccu_fmx_upload")) return rc;
  if (nbcs < 0 || nsnps < 0 || !cell_ptr || (nsnps > 0 && !af)) return ccu_fail(h, "ccu_fmx_upload: bad arguments");
  if (cell_ptr[0] != 0) return ccu_fail(h, "ccu_fmx_upload: cell_ptr[0] must be 0");
  const int64_t nnz = cell_ptr[nbcs];
  if (nnz < 0 || nnz > 2147483647LL) return ccu_fail(h, "ccu_fmx_upload: entry count %lld out of range", (long long)nnz);
  if (nnz > 0 && (!snp_idx || !read_ptr || read_ptr[0] != 0)) return ccu_fail(h, "ccu_fmx_upload: bad CSR arrays");
  const int64_t nreads = nnz > 0 ? read_ptr[nnz] : 0;
  if (nreads < 0 || (nreads > 0 && (!al || !bq))) return ccu_fail(h, "ccu_fmx_upload: bad read arrays");
  CCU_CUDA(h, cudaSetDevice(h->device));
  FmxState* f = h->fmx;
  f->close_peers();  // the mapped arrays may be re-allocated below
  cudaStream_t st = h->stream;
  f->nbcs = nbcs;
  f->nsnps = nsnps;
  f->nnz = nnz;
  f->nreads = nreads;
  {  // most droplets per SNP
    std::vector<int32_t> per_snp(nsnps, 0);
    int32_t mx = 0;
    for (int64_t e = 0; e < nnz; ++e) {
      const int32_t s = snp_idx[e];
      if (s < 0 || s >= nsnps) return ccu_fail(h, "ccu_fmx_upload: snp_idx[%lld] = %d out of range", (long long)e, s);
      mx = std::max(mx, ++per_snp[s]);
      if (read_ptr[e + 1] - read_ptr[e] > (int64_t)ccu::kCntMask)  // the packed counts of the SNP-major rows
        return ccu_fail(h, "ccu_fmx_upload: entry %lld has %lld reads (at most %lld per droplet and SNP)", (long long)e,
                        (long long)(read_ptr[e + 1] - read_ptr[e]), (long long)ccu::kCntMask);
    }
    f->max_segment = mx;
  }
  f->cell_begin = 0;
  f->cell_end = nbcs;
  f->snp_begin = 0;
  f->snp_end = nsnps;
  f->uploaded = f->have_stats = f->have_estep = f->have_post = f->post_full = false;
  f->post_apply = -1;
  CCU_CUDA(h, f->cell_ptr.reserve((nbcs + 1) * sizeof(int64_t)));
  CCU_CUDA(h, f->snp_idx.reserve(nnz * sizeof(int32_t) + 16));
  CCU_CUDA(h, f->read_ptr.reserve((nnz + 1) * sizeof(int64_t)));
  CCU_CUDA(h, f->al.reserve(nreads + 16));
  CCU_CUDA(h, f->bq.reserve(nreads + 16));
  CCU_CUDA(h, f->af.reserve(nsnps * sizeof(double) + 16));
  CCU_CUDA(h, f->gl.reserve((size_t)nnz * 9 * sizeof(double) + 16));
  CCU_CUDA(h, f->cnt.reserve((size_t)nnz * sizeof(int4) + 16));
  CCU_CUDA(h, f->logdenom.reserve((size_t)nnz * sizeof(double) + 16));
  CCU_CUDA(h, f->csc_ptr.reserve((nsnps + 1) * sizeof(int64_t)));
  CCU_CUDA(h, f->csc_cell.reserve((size_t)nnz * sizeof(int32_t) + 16));
  CCU_CUDA(h, f->csc_row.reserve((size_t)nnz * ccu::kCscRow * sizeof(double) + 32));
  CCU_CUDA(h, f->assign.reserve(nbcs * sizeof(int32_t) + 16));
  CCU_CUDA(h, f->stats.reserve((size_t)f->K * nsnps * ccu::kStatStride * sizeof(double) + 16));
  CCU_CUDA(h, f->cgp.reserve((size_t)f->K * nsnps * 3 * sizeof(double) + 16));
  CCU_CUDA(h, f->cgm.reserve((size_t)f->K * nsnps * sizeof(double) + 16));
  CCU_CUDA(h, f->eclass.reserve((size_t)nnz + 16));
  CCU_CUDA(h, f->egen.reserve((size_t)nnz * sizeof(int32_t) + 16));
  CCU_CUDA(h, f->eab.reserve((size_t)nnz * sizeof(double2) + 16));
  CCU_CUDA(h, f->esnp.reserve((size_t)nnz * sizeof(int32_t) + 16));
  CCU_CUDA(h, f->gcnt.reserve((size_t)nbcs * sizeof(int32_t) + 16));
  CCU_CUDA(h, f->counter.reserve(16));
  CCU_CUDA(h, cudaMemsetAsync(f->counter.p, 0, 16, st));
  CCU_CUDA(h, f->snp_gen.reserve((size_t)nsnps + 16));
  CCU_CUDA(h, cudaMemsetAsync(f->snp_gen.p, 0, (size_t)nsnps + 16, st));  // set by the compaction kernel
  if (int rc = reserve_shard_buffers(h)) return rc;
  CCU_CUDA(h, cudaMemcpyAsync(f->cell_ptr.p, cell_ptr, (nbcs + 1) * sizeof(int64_t), cudaMemcpyHostToDevice, st));
  if (nnz > 0) {
    CCU_CUDA(h, cudaMemcpyAsync(f->snp_idx.p, snp_idx, nnz * sizeof(int32_t), cudaMemcpyHostToDevice, st));
    CCU_CUDA(h, cudaMemcpyAsync(f->read_ptr.p, read_ptr, (nnz + 1) * sizeof(int64_t), cudaMemcpyHostToDevice, st));
  }
  if (nreads > 0) {
    CCU_CUDA(h, cudaMemcpyAsync(f->al.p, al, nreads, cudaMemcpyHostToDevice, st));
    CCU_CUDA(h, cudaMemcpyAsync(f->bq.p, bq, nreads, cudaMemcpyHostToDevice, st));
  }
  if (nsnps > 0) CCU_CUDA(h, cudaMemcpyAsync(f->af.p, af, nsnps * sizeof(double), cudaMemcpyHostToDevice, st));
  CCU_CUDA(h, cudaMemsetAsync(f->assign.p, 0xff, nbcs * sizeof(int32_t) + 16, st));  // -1: unassigned

  const ccu::FmxDev d = fmx_view(h);
  f->tm.launches = 0;
  {  // F1: alpha = 0.5, cmd_cram_freemuxlet.cpp:133
    Timer t(h, &f->tm.ms_pileup);
    CCU_CUDA(h, ccu::launch_fmx_pileup(d, 0.5, st));
    CCU_CUDA(h, ccu::launch_fmx_compact(d, st));  // the E-step's view: general entries listed, single-read ones packed
    t.stop_sync();
    f->tm.launches += nnz > 0 ? 2 : 0;
  }
  {  // SNP-major index: stable radix sort of (snp, entry id); entries of one SNP stay in ascending droplet order
    Timer t(h, &f->tm.ms_csc);
    // scratch: ent_cell | keys_in(=snp_idx) -> keys_out | vals_in | vals_out | cub temp
    const size_t n = (size_t)nnz;
    size_t temp_bytes = 0;
    int end_bit = 1;
    while ((1LL << end_bit) < nsnps && end_bit < 31) ++end_bit;
    cub::DeviceRadixSort::SortPairs(nullptr, temp_bytes, (const int32_t*)nullptr, (int32_t*)nullptr,
                                    (const int32_t*)nullptr, (int32_t*)nullptr, (int)n, 0, end_bit, st);
    const size_t a4 = (n * sizeof(int32_t) + 255) / 256 * 256;
    CCU_CUDA(h, f->scratch.reserve(4 * a4 + temp_bytes + 256));
    char* base = f->scratch.as<char>();
    int32_t* ent_cell = reinterpret_cast<int32_t*>(base);
    int32_t* keys_out = reinterpret_cast<int32_t*>(base + a4);
    int32_t* vals_in = reinterpret_cast<int32_t*>(base + 2 * a4);
    int32_t* vals_out = reinterpret_cast<int32_t*>(base + 3 * a4);
    void* temp = base + 4 * a4;
    CCU_CUDA(h, ccu::launch_fmx_entry_cells(d, ent_cell, st));
    if (n > 0) {
      CCU_CUDA(h, ccu::launch_fmx_iota(vals_in, (int64_t)n, st));
      CCU_CUDA(h, cub::DeviceRadixSort::SortPairs(temp, temp_bytes, d.snp_idx, keys_out, vals_in, vals_out, (int)n, 0,
                                                  end_bit, st));
    }
    CCU_CUDA(h, ccu::launch_fmx_build_csc(d, keys_out, vals_out, ent_cell, st));
    t.stop_sync();
    f->tm.launches += 4;
  }
  f->uploaded = true;
  return 0;
}

int ccu_fmx_set_shard(ccu_handle* h, int64_t cell_begin, int64_t cell_end, int64_t snp_begin, int64_t snp_end) {
  if (int rc = need(h, true, "ccu_fmx_set_shard")) return rc;
  FmxState* f = h->fmx;
  if (cell_begin < 0 || cell_end > f->nbcs || cell_begin > cell_end || snp_begin < 0 || snp_end > f->nsnps ||
      snp_begin > snp_end)
    return ccu_fail(h, "ccu_fmx_set_shard: range out of bounds");
  CCU_CUDA(h, cudaSetDevice(h->device));
  if (f->cell_begin == cell_begin && f->cell_end == cell_end && f->snp_begin == snp_begin && f->snp_end == snp_end &&
      f->results.p)
    return 0;  // same shard: keep what has been computed for it (a replayed CUDA graph makes no host-side calls)
  f->cell_begin = cell_begin;
  f->cell_end = cell_end;
  f->snp_begin = snp_begin;
  f->snp_end = snp_end;
  f->have_estep = false;
  return reserve_shard_buffers(h);
}

int ccu_fmx_lmix(ccu_handle* h, double* llk0, double* llk2) {
  CcuRange nvtx_range("ccu_fmx_lmix");
  if (int rc = need(h, true, "ccu_fmx_lmix")) return rc;
  FmxState* f = h->fmx;
  if (f->nbcs > 0 && (!llk0 || !llk2)) return ccu_fail(h, "ccu_fmx_lmix: NULL output");
  CCU_CUDA(h, cudaSetDevice(h->device));
  CCU_CUDA(h, f->scratch.reserve(2 * f->nbcs * sizeof(double) + 16));
  double* d0 = f->scratch.as<double>();
  double* d2 = d0 + f->nbcs;
  Timer t(h, &f->tm.ms_lmix);
  CCU_CUDA(h, ccu::launch_fmx_lmix(fmx_view(h), d0, d2, h->stream));
  t.stop_sync();
  if (f->nbcs > 0) {
    CCU_CUDA(h, cudaMemcpy(llk0, d0, f->nbcs * sizeof(double), cudaMemcpyDeviceToHost));
    CCU_CUDA(h, cudaMemcpy(llk2, d2, f->nbcs * sizeof(double), cudaMemcpyDeviceToHost));
  }
  return 0;
}

int ccu_fmx_get_pileups(ccu_handle* h, ccu_pileup* out) {
  if (int rc = need(h, true, "ccu_fmx_get_pileups")) return rc;
  FmxState* f = h->fmx;
  if (f->nnz == 0) return 0;
  if (!out) return ccu_fail(h, "ccu_fmx_get_pileups: NULL output");
  CCU_CUDA(h, cudaSetDevice(h->device));
  CCU_CUDA(h, cudaStreamSynchronize(h->stream));
  std::vector<double> gl((size_t)f->nnz * 9), ld(f->nnz);
  std::vector<int4> cnt(f->nnz);
  CCU_CUDA(h, cudaMemcpy(gl.data(), f->gl.p, gl.size() * sizeof(double), cudaMemcpyDeviceToHost));
  CCU_CUDA(h, cudaMemcpy(ld.data(), f->logdenom.p, ld.size() * sizeof(double), cudaMemcpyDeviceToHost));
  CCU_CUDA(h, cudaMemcpy(cnt.data(), f->cnt.p, cnt.size() * sizeof(int4), cudaMemcpyDeviceToHost));
  for (int64_t e = 0; e < f->nnz; ++e) {
    out[e].nreads = cnt[e].x;
    out[e].nref = cnt[e].y;
    out[e].nalt = cnt[e].z;
    out[e]._pad = 0;
    memcpy(out[e].gls, &gl[e * 9], 9 * sizeof(double));
    out[e].logdenom = ld[e];
  }
  return 0;
}

// M-step of this shard's SNP slab.  The fused kernel also leaves the posteriors / mean genotypes of the slab (without
// the genotype-error mix; ccu_fmx_estep_dev recomputes them when it needs the other flavour or more than the slab).
int ccu_fmx_mstep_dev(ccu_handle* h) {
  if (int rc = need(h, true, "ccu_fmx_mstep")) return rc;
  CCU_CUDA(h, cudaSetDevice(h->device));
  FmxState* f = h->fmx;
  cudaEventRecord(f->ev[2], h->stream);
  CCU_CUDA(h, ccu::launch_fmx_mstep(fmx_view(h), 0, 1, no_peers(), make_sync(h, false), h->stream));
  cudaEventRecord(f->ev[3], h->stream);
  f->tm.launches += 1;
  f->have_stats = true;
  f->post_full = (f->snp_begin == 0 && f->snp_end == f->nsnps);
  f->post_apply = 0;
  f->cgm_stale = false;
  return 0;
}

int ccu_fmx_mstep(ccu_handle* h, const int32_t* assign) {
  if (int rc = need(h, true, "ccu_fmx_mstep")) return rc;
  FmxState* f = h->fmx;
  if (f->nbcs > 0 && !assign) return ccu_fail(h, "ccu_fmx_mstep: NULL assignment vector");
  CCU_CUDA(h, cudaSetDevice(h->device));
  for (int64_t c = 0; c < f->nbcs; ++c)
    if (assign[c] >= f->K) return ccu_fail(h, "ccu_fmx_mstep: assign[%lld] = %d >= nclust", (long long)c, assign[c]);
  if (f->nbcs > 0)
    CCU_CUDA(h, cudaMemcpyAsync(f->assign.p, assign, f->nbcs * sizeof(int32_t), cudaMemcpyHostToDevice, h->stream));
  if (int rc = ccu_fmx_mstep_dev(h)) return rc;
  CCU_CUDA(h, cudaEventSynchronize(f->ev[3]));
  cudaEventElapsedTime(&f->tm.ms_mstep, f->ev[2], f->ev[3]);
  return 0;
}

int ccu_fmx_estep_dev(ccu_handle* h, int32_t apply_geno_error) {
  if (int rc = need(h, true, "ccu_fmx_estep")) return rc;
  FmxState* f = h->fmx;
  if (!f->have_stats) return ccu_fail(h, "ccu_fmx_estep: no cluster statistics yet (run ccu_fmx_mstep first)");
  CCU_CUDA(h, cudaSetDevice(h->device));
  const ccu::FmxDev d = fmx_view(h);
  const int apply = apply_geno_error ? 1 : 0;
  if (!f->post_full || f->post_apply != apply || f->cgm_stale) {  // posteriors of every SNP from the statistics array
    CCU_CUDA(h, ccu::launch_fmx_cgp(d, apply, false, h->stream));
    f->tm.launches += 1;
    f->post_full = true;
    f->post_apply = apply;
    f->cgm_stale = false;
  }
  cudaEventRecord(f->ev[0], h->stream);
  CCU_CUDA(h, ccu::launch_fmx_estep(d, no_peers(), make_sync(h, false), h->stream));
  cudaEventRecord(f->ev[1], h->stream);
  f->tm.launches += 2;  // E-step, decision
  f->have_estep = f->have_post = true;
  return 0;
}

int ccu_fmx_posteriors_dev(ccu_handle* h, int32_t apply_geno_error) {
  if (int rc = need(h, true, "ccu_fmx_posteriors")) return rc;
  FmxState* f = h->fmx;
  if (!f->have_stats) return ccu_fail(h, "ccu_fmx_posteriors: no cluster statistics yet (run ccu_fmx_mstep first)");
  CCU_CUDA(h, cudaSetDevice(h->device));
  CCU_CUDA(h, ccu::launch_fmx_cgp(fmx_view(h), apply_geno_error, true, h->stream));
  f->tm.launches += 1;
  f->have_post = true;
  f->post_full = false;  // the caller completes the array with a sum-all-reduce
  f->post_apply = apply_geno_error ? 1 : 0;
  f->cgm_stale = true;
  return 0;
}

int ccu_fmx_estep_post_dev(ccu_handle* h) {
  if (int rc = need(h, true, "ccu_fmx_estep_post")) return rc;
  FmxState* f = h->fmx;
  if (!f->have_post) return ccu_fail(h, "ccu_fmx_estep_post: no posteriors yet (run ccu_fmx_posteriors_dev first)");
  CCU_CUDA(h, cudaSetDevice(h->device));
  const ccu::FmxDev d = fmx_view(h);
  if (f->cgm_stale) {  // the posterior array was completed by a collective: derive the mean genotypes from it
    CCU_CUDA(h, ccu::launch_fmx_cgm(d, h->stream));
    f->tm.launches += 1;
    f->cgm_stale = false;
  }
  cudaEventRecord(f->ev[0], h->stream);
  CCU_CUDA(h, ccu::launch_fmx_estep(d, no_peers(), make_sync(h, false), h->stream));
  cudaEventRecord(f->ev[1], h->stream);
  f->tm.launches += 2;  // E-step, decision
  f->have_estep = true;
  return 0;
}

// ---- the fused loop: two launches + the decision per iteration, exchange inside the kernels ----
int ccu_fmx_mstep_fused_dev(ccu_handle* h, int32_t apply_geno_error_next, int32_t want_stats, int32_t in_kernel_sync) {
  CcuRange nvtx_range("ccu_fmx_mstep_fused_dev");
  if (int rc = need(h, true, "ccu_fmx_mstep_fused")) return rc;
  CCU_CUDA(h, cudaSetDevice(h->device));
  FmxState* f = h->fmx;
  cudaEventRecord(f->ev[2], h->stream);
  CCU_CUDA(h, ccu::launch_fmx_mstep(fmx_view(h), apply_geno_error_next ? 1 : 0, want_stats ? 1 : 0, f->peers,
                                    make_sync(h, in_kernel_sync != 0), h->stream));
  cudaEventRecord(f->ev[3], h->stream);
  f->tm.launches += 1;
  if (want_stats) f->have_stats = true;
  f->have_post = true;
  f->post_full = f->peers.n > 1 || (f->snp_begin == 0 && f->snp_end == f->nsnps);  // peers fill the other slabs
  f->post_apply = apply_geno_error_next ? 1 : 0;
  f->cgm_stale = false;
  return 0;
}

int ccu_fmx_estep_fused_dev(ccu_handle* h, int32_t in_kernel_sync) {
  CcuRange nvtx_range("ccu_fmx_estep_fused_dev");
  if (int rc = need(h, true, "ccu_fmx_estep_fused")) return rc;
  FmxState* f = h->fmx;
  if (!f->have_post || !f->post_full)
    return ccu_fail(h, "ccu_fmx_estep_fused: the posteriors of every SNP are needed (ccu_fmx_mstep_fused_dev on every rank first)");
  CCU_CUDA(h, cudaSetDevice(h->device));
  cudaEventRecord(f->ev[0], h->stream);
  CCU_CUDA(h, ccu::launch_fmx_estep(fmx_view(h), f->peers, make_sync(h, in_kernel_sync != 0), h->stream));
  cudaEventRecord(f->ev[1], h->stream);
  f->tm.launches += 2;  // E-step, decision
  f->have_estep = true;
  return 0;
}

// ---- peer-memory exchange over NVLink (one process per GPU) ----
int ccu_fmx_peer_handles(ccu_handle* h, void* out) {
  if (int rc = need(h, true, "ccu_fmx_peer_handles")) return rc;
  FmxState* f = h->fmx;
  if (!out) return ccu_fail(h, "ccu_fmx_peer_handles: NULL output");
  if (!f->cgp.p || !f->cgm.p || !f->assign.p) return ccu_fail(h, "ccu_fmx_peer_handles: call ccu_fmx_set_shard first");
  static_assert(sizeof(cudaIpcMemHandle_t) == CCU_IPC_HANDLE_BYTES, "CUDA IPC handle size");
  CCU_CUDA(h, cudaSetDevice(h->device));
  // the flag array of the peer-memory barrier: zeroed here, i.e. before any peer can learn its address
  CCU_CUDA(h, f->flags.reserve((ccu::kMaxPeers + 2) * sizeof(uint32_t)));
  CCU_CUDA(h, cudaMemsetAsync(f->flags.p, 0, (ccu::kMaxPeers + 2) * sizeof(uint32_t), h->stream));
  CCU_CUDA(h, cudaStreamSynchronize(h->stream));
  cudaIpcMemHandle_t hd[4];
  CCU_CUDA(h, cudaIpcGetMemHandle(&hd[0], f->cgp.p));
  CCU_CUDA(h, cudaIpcGetMemHandle(&hd[1], f->cgm.p));
  CCU_CUDA(h, cudaIpcGetMemHandle(&hd[2], f->assign.p));
  CCU_CUDA(h, cudaIpcGetMemHandle(&hd[3], f->flags.p));
  memcpy(out, hd, sizeof(hd));
  return 0;
}

int ccu_fmx_peer_open(ccu_handle* h, int32_t nranks, int32_t rank, const void* handles) {
  if (int rc = need(h, true, "ccu_fmx_peer_open")) return rc;
  FmxState* f = h->fmx;
  if (nranks < 1 || nranks > ccu::kMaxPeers || rank < 0 || rank >= nranks || !handles)
    return ccu_fail(h, "ccu_fmx_peer_open: bad arguments (at most %d ranks)", ccu::kMaxPeers);
  if (!f->cgp.p || !f->assign.p || !f->flags.p) return ccu_fail(h, "ccu_fmx_peer_open: call ccu_fmx_peer_handles first");
  CCU_CUDA(h, cudaSetDevice(h->device));
  f->close_peers();
  const cudaIpcMemHandle_t* hd = static_cast<const cudaIpcMemHandle_t*>(handles);
  f->peer_rank = rank;
  f->peers.n = nranks;
  for (int r = 0; r < nranks; ++r) {
    if (r == rank) {
      f->peers.post[r] = f->cgp.as<double>();
      f->peers.postm[r] = f->cgm.as<double>();
      f->peers.assign[r] = f->assign.as<int32_t>();
      f->peers.flags[r] = f->flags.as<uint32_t>();
      continue;
    }
    void* pp[4] = {nullptr, nullptr, nullptr, nullptr};
    cudaError_t e = cudaSuccess;
    for (int k = 0; k < 4 && e == cudaSuccess; ++k) e = cudaIpcOpenMemHandle(&pp[k], hd[4 * r + k], cudaIpcMemLazyEnablePeerAccess);
    f->peers.post[r] = static_cast<double*>(pp[0]);
    f->peers.postm[r] = static_cast<double*>(pp[1]);
    f->peers.assign[r] = static_cast<int32_t*>(pp[2]);
    f->peers.flags[r] = static_cast<uint32_t*>(pp[3]);
    if (e != cudaSuccess) {
      f->peers.n = r + 1;
      f->close_peers();
      return ccu_fail(h, "ccu_fmx_peer_open: cannot map the arrays of rank %d: %s", r, cudaGetErrorString(e));
    }
  }
  return 0;
}

int ccu_fmx_peer_close(ccu_handle* h) {
  if (!h || !h->fmx) return 1;
  cudaSetDevice(h->device);
  cudaStreamSynchronize(h->stream);
  h->fmx->close_peers();
  return 0;
}

// Full barrier across the ranks as one tiny kernel (arrive + wait): what orders the un-synchronised form of the
// fused calls (in_kernel_sync = 0), e.g. several handles on ONE device, where a kernel that spins at its head could
// keep the peer it waits for from being scheduled.
int ccu_fmx_rank_barrier_dev(ccu_handle* h) {
  if (int rc = need(h, true, "ccu_fmx_rank_barrier")) return rc;
  FmxState* f = h->fmx;
  if (f->peers.n < 1) return ccu_fail(h, "ccu_fmx_rank_barrier: call ccu_fmx_peer_open first");
  CCU_CUDA(h, cudaSetDevice(h->device));
  CCU_CUDA(h, ccu::launch_fmx_sync(f->peers, make_sync(h, true), h->stream));
  f->tm.launches += (f->peers.n > 1) ? 1 : 0;
  return 0;
}

int ccu_fmx_rank_barrier_failed(ccu_handle* h, int32_t* failed) {
  if (int rc = need(h, true, "ccu_fmx_rank_barrier_failed")) return rc;
  FmxState* f = h->fmx;
  if (!failed) return ccu_fail(h, "ccu_fmx_rank_barrier_failed: NULL output");
  *failed = 0;
  if (!f->flags.p) return 0;
  CCU_CUDA(h, cudaSetDevice(h->device));
  uint32_t v[ccu::kMaxPeers + 2] = {};
  CCU_CUDA(h, cudaStreamSynchronize(h->stream));
  CCU_CUDA(h, cudaMemcpy(v, f->flags.p, sizeof(v), cudaMemcpyDeviceToHost));
  if (v[ccu::kMaxPeers + 1]) {
    *failed = 1;
    char late[256] = "";
    size_t n = 0;
    for (int r = 0; r < f->peers.n && n + 24 < sizeof(late); ++r)
      if (r != f->peer_rank && (int32_t)(v[r] - v[ccu::kMaxPeers]) < 0)
        n += (size_t)snprintf(late + n, sizeof(late) - n, " %d(at %u)", r, v[r]);
    ccu_fail(h, "freemuxlet multi-GPU loop: rank %d gave up waiting at epoch %u; ranks behind:%s (time-out: CCU_FMX_BARRIER_TIMEOUT_S)",
             f->peer_rank, v[ccu::kMaxPeers], n ? late : " none (a peer gave up first)");
  }
  return 0;
}

void* ccu_fmx_post_dev(ccu_handle* h, int64_t* n_doubles) {
  if (!h || !h->fmx || !h->fmx->uploaded) return nullptr;
  if (n_doubles) *n_doubles = (int64_t)h->fmx->K * h->fmx->nsnps * 3;
  return h->fmx->cgp.p;
}

int ccu_fmx_download_results(ccu_handle* h, ccu_fmx_result* out) {
  if (int rc = need(h, true, "ccu_fmx_download_results")) return rc;
  FmxState* f = h->fmx;
  if (!f->have_estep) return ccu_fail(h, "ccu_fmx_download_results: no E-step has run");
  const int64_t ncell = f->cell_end - f->cell_begin;
  CCU_CUDA(h, cudaSetDevice(h->device));
  if (ncell > 0) {
    if (!out) return ccu_fail(h, "ccu_fmx_download_results: NULL output");
    CCU_CUDA(h, cudaMemcpyAsync(out, f->results.p, ncell * sizeof(ccu_fmx_result), cudaMemcpyDeviceToHost, h->stream));
  }
  CCU_CUDA(h, cudaStreamSynchronize(h->stream));
  return 0;
}

int ccu_fmx_estep(ccu_handle* h, int32_t apply_geno_error, ccu_fmx_result* out, int32_t* assign_out) {
  if (int rc = ccu_fmx_estep_dev(h, apply_geno_error)) return rc;
  FmxState* f = h->fmx;
  CCU_CUDA(h, cudaEventSynchronize(f->ev[1]));
  cudaEventElapsedTime(&f->tm.ms_estep, f->ev[0], f->ev[1]);
  if (int rc = ccu_fmx_download_results(h, out)) return rc;
  const int64_t ncell = f->cell_end - f->cell_begin;
  if (assign_out && ncell > 0)
    CCU_CUDA(h, cudaMemcpy(assign_out, f->assign.as<int32_t>() + f->cell_begin, ncell * sizeof(int32_t),
                           cudaMemcpyDeviceToHost));
  return 0;
}

int ccu_fmx_get_llk(ccu_handle* h, double* llk) {
  if (int rc = need(h, true, "ccu_fmx_get_llk")) return rc;
  FmxState* f = h->fmx;
  if (!f->have_estep) return ccu_fail(h, "ccu_fmx_get_llk: no E-step has run");
  const int64_t n = (f->cell_end - f->cell_begin) * ((int64_t)f->K * (f->K + 1) / 2);
  if (n == 0) return 0;
  CCU_CUDA(h, cudaSetDevice(h->device));
  // Inspection only: an E-step never stores the matrix (its epilogue works on the products, not on their logs).  The
  // E-step is run again here on the cluster posteriors as they stand -- right after ccu_fmx_estep those are the ones
  // it used -- with the log-domain epilogue, which writes every llk[droplet][pair]; results and assignments are
  // rewritten with the same values.
  CCU_CUDA(h, f->llk.reserve((size_t)n * sizeof(double) + 16));
  ccu::FmxDev d = fmx_view(h);
  d.llk = f->llk.as<double>();
  CCU_CUDA(h, ccu::launch_fmx_estep(d, no_peers(), make_sync(h, false), h->stream));
  CCU_CUDA(h, cudaStreamSynchronize(h->stream));
  CCU_CUDA(h, cudaMemcpy(llk, f->llk.p, n * sizeof(double), cudaMemcpyDeviceToHost));
  return 0;
}

int ccu_fmx_get_clusters(ccu_handle* h, ccu_pileup* out) {
  if (int rc = need(h, true, "ccu_fmx_get_clusters")) return rc;
  FmxState* f = h->fmx;
  if (!f->have_stats) return ccu_fail(h, "ccu_fmx_get_clusters: no M-step has run");
  const int64_t n = (int64_t)f->K * f->nsnps;
  if (n == 0) return 0;
  CCU_CUDA(h, cudaSetDevice(h->device));
  CCU_CUDA(h, f->scratch.reserve(n * sizeof(ccu_pileup) + 16));
  CCU_CUDA(h, ccu::launch_fmx_unpack_stats(fmx_view(h), f->scratch.as<ccu_pileup>(), h->stream));
  CCU_CUDA(h, cudaMemcpyAsync(out, f->scratch.p, n * sizeof(ccu_pileup), cudaMemcpyDeviceToHost, h->stream));
  CCU_CUDA(h, cudaStreamSynchronize(h->stream));
  return 0;
}

void* ccu_fmx_stats_dev(ccu_handle* h, int64_t* n_doubles) {
  if (!h || !h->fmx || !h->fmx->uploaded) return nullptr;
  if (n_doubles) *n_doubles = (int64_t)h->fmx->K * h->fmx->nsnps * ccu::kStatStride;
  return h->fmx->stats.p;
}

void* ccu_fmx_assign_dev(ccu_handle* h, int64_t* n_int32) {
  if (!h || !h->fmx || !h->fmx->uploaded) return nullptr;
  if (n_int32) *n_int32 = h->fmx->nbcs;
  return h->fmx->assign.p;
}

int ccu_fmx_run_em(ccu_handle* h, const int32_t* init_assign, int32_t niter, int32_t geno_error_every_iter,
                   int32_t stop_when_stable, ccu_fmx_result* out, int32_t* iters_run) {
  CcuRange nvtx_range("ccu_fmx_run_em");
  if (int rc = need(h, true, "ccu_fmx_run_em")) return rc;
  FmxState* f = h->fmx;
  if (f->cell_begin != 0 || f->cell_end != f->nbcs || f->snp_begin != 0 || f->snp_end != f->nsnps)
    return ccu_fail(h, "ccu_fmx_run_em: single-GPU loop needs the full shard (use the *_dev calls with collectives)");
  if (niter < 1) return ccu_fail(h, "ccu_fmx_run_em: niter must be >= 1");
  if (f->peers.n > 1) return ccu_fail(h, "ccu_fmx_run_em: peers are mapped on this handle (ccu_fmx_peer_close first)");
  CCU_CUDA(h, cudaSetDevice(h->device));
  for (int64_t c = 0; c < f->nbcs; ++c)
    if (!init_assign || init_assign[c] >= f->K) return ccu_fail(h, "ccu_fmx_run_em: bad initial assignment at droplet %lld", (long long)c);
  if (f->nbcs > 0)
    CCU_CUDA(h, cudaMemcpyAsync(f->assign.p, init_assign, f->nbcs * sizeof(int32_t), cudaMemcpyHostToDevice, h->stream));
  // geno_error is applied by the E-step of the last iteration only (freemuxlet, :485) or of every iteration (freemux2,
  // cmd_cram_freemux2.cpp:410-415); the M-step before it forms the posteriors, so it gets the flag of the E-step
  // that follows.  The 96-byte statistics records are only written by the M-step that may be the last one.
  auto apply_at = [&](int it) { return geno_error_every_iter ? 1 : (it + 1 == niter ? 1 : 0); };
  if (int rc = ccu_fmx_mstep_fused_dev(h, apply_at(0), 0, 0)) return rc;  // cmd_cram_freemuxlet.cpp:359-370
  std::vector<ccu_fmx_result> prev, cur;
  int it = 0;
  for (; it < niter; ++it) {
    if (int rc = ccu_fmx_estep_fused_dev(h, 0)) return rc;
    const bool maybe_last = stop_when_stable || it + 1 == niter;
    if (int rc = ccu_fmx_mstep_fused_dev(h, apply_at(it + 1), maybe_last ? 1 : 0, 0)) return rc;  // :639-650
    if (stop_when_stable) {  // cmd_cram_freemux2.cpp:518-604
      cur.resize(f->nbcs);
      if (int rc = ccu_fmx_download_results(h, cur.data())) return rc;
      int64_t nchanged = 0;
      for (int64_t c = 0; c < f->nbcs; ++c) {
        const bool had = !prev.empty();
        const int ptype = had ? prev[c].type : -1, pj = had ? prev[c].jBest : -1, pk = had ? prev[c].kBest : -1;
        if (cur[c].type == 0) nchanged += (ptype != 0 || pj != cur[c].sBest || pk != cur[c].sBest);
        else nchanged += (ptype != cur[c].type);
      }
      prev.swap(cur);
      if (nchanged == 0) {
        ++it;
        break;
      }
    }
  }
  // device times of the last iteration's two phases
  CCU_CUDA(h, cudaStreamSynchronize(h->stream));
  cudaEventElapsedTime(&f->tm.ms_estep, f->ev[0], f->ev[1]);
  cudaEventElapsedTime(&f->tm.ms_mstep, f->ev[2], f->ev[3]);
  if (iters_run) *iters_run = it;
  return ccu_fmx_download_results(h, out);
}

// ---- in-process multi-GPU EM: one host thread per handle, exchange through peer memory ----
int ccu_fmx_run_em_multi(ccu_handle** hs, int32_t n, const int64_t* cell_ptr, const int32_t* snp_idx,
                         const int32_t* init_assign, int32_t niter, int32_t geno_error_every_iter, ccu_fmx_result* out,
                         ccu_pileup* clusters_out) {
  CcuRange nvtx_range("ccu_fmx_run_em_multi");
  if (!hs || n < 1 || !hs[0]) return 1;
  ccu_handle* h0 = hs[0];
  if (n > ccu::kMaxPeers) return ccu_fail(h0, "ccu_fmx_run_em_multi: at most %d handles", ccu::kMaxPeers);
  if (niter < 1) return ccu_fail(h0, "ccu_fmx_run_em_multi: niter must be >= 1");
  for (int i = 0; i < n; ++i) {
    if (!hs[i]) return ccu_fail(h0, "ccu_fmx_run_em_multi: NULL handle");
    if (int rc = need(hs[i], true, "ccu_fmx_run_em_multi")) return rc;
    const FmxState *a = hs[i]->fmx, *b = h0->fmx;
    if (a->nbcs != b->nbcs || a->nsnps != b->nsnps || a->nnz != b->nnz || a->K != b->K || a->geno_error != b->geno_error ||
        a->doublet_prior != b->doublet_prior)
      return ccu_fail(h0, "ccu_fmx_run_em_multi: handle %d holds a different store or configuration", i);
    for (int j = 0; j < i; ++j)
      if (hs[j] == hs[i]) return ccu_fail(h0, "ccu_fmx_run_em_multi: handle %d given twice", i);
  }
  const FmxState* f0 = h0->fmx;
  const int64_t nbcs = f0->nbcs, nsnps = f0->nsnps, nnz = f0->nnz;
  if (nbcs > 0 && (!cell_ptr || !init_assign || !out)) return ccu_fail(h0, "ccu_fmx_run_em_multi: NULL argument");
  if (nnz > 0 && !snp_idx) return ccu_fail(h0, "ccu_fmx_run_em_multi: NULL snp_idx");
  // droplet ranges and SNP slabs with about nnz / n entries each
  std::vector<int64_t> cb(n + 1, nbcs), sb(n + 1, nsnps);
  cb[0] = sb[0] = 0;
  {
    std::vector<int64_t> per_snp((size_t)nsnps + 1, 0);
    for (int64_t e = 0; e < nnz; ++e) ++per_snp[(size_t)snp_idx[e] + 1];
    for (int64_t v = 0; v < nsnps; ++v) per_snp[(size_t)v + 1] += per_snp[(size_t)v];
    for (int i = 1; i < n; ++i) {
      const int64_t target = nnz * i / n;
      cb[i] = std::lower_bound(cell_ptr, cell_ptr + nbcs + 1, target) - cell_ptr;
      cb[i] = std::min(std::max(cb[i], cb[i - 1]), nbcs);
      sb[i] = std::lower_bound(per_snp.begin(), per_snp.end(), target) - per_snp.begin();
      sb[i] = std::min(std::max(sb[i], sb[i - 1]), nsnps);
    }
  }
  for (int i = 0; i < n; ++i)
    if (int rc = ccu_fmx_set_shard(hs[i], cb[i], cb[i + 1], sb[i], sb[i + 1])) return rc;
  // peer access between the devices, flag arrays, pointer tables
  for (int i = 0; i < n; ++i) {
    ccu_handle* h = hs[i];
    FmxState* f = h->fmx;
    CCU_CUDA(h, cudaSetDevice(h->device));
    f->close_peers();
    for (int j = 0; j < n; ++j)
      if (hs[j]->device != h->device) {
        int can = 0;
        CCU_CUDA(h, cudaDeviceCanAccessPeer(&can, h->device, hs[j]->device));
        if (!can) return ccu_fail(h, "ccu_fmx_run_em_multi: device %d cannot access device %d", h->device, hs[j]->device);
        const cudaError_t e = cudaDeviceEnablePeerAccess(hs[j]->device, 0);
        if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled)
          return ccu_fail(h, "cudaDeviceEnablePeerAccess(%d -> %d): %s", h->device, hs[j]->device, cudaGetErrorString(e));
        cudaGetLastError();
      }
    CCU_CUDA(h, f->flags.reserve((ccu::kMaxPeers + 2) * sizeof(uint32_t)));
    CCU_CUDA(h, cudaMemsetAsync(f->flags.p, 0, (ccu::kMaxPeers + 2) * sizeof(uint32_t), h->stream));
    CCU_CUDA(h, cudaStreamSynchronize(h->stream));
  }
  for (int i = 0; i < n; ++i) {
    FmxState* f = hs[i]->fmx;
    f->peers_are_ipc = false;
    f->peer_rank = i;
    f->peers.n = n;
    for (int j = 0; j < n; ++j) {
      f->peers.post[j] = hs[j]->fmx->cgp.as<double>();
      f->peers.postm[j] = hs[j]->fmx->cgm.as<double>();
      f->peers.assign[j] = hs[j]->fmx->assign.as<int32_t>();
      f->peers.flags[j] = hs[j]->fmx->flags.as<uint32_t>();
    }
  }
  // With one handle per device the kernels of the loop order the ranks themselves (arrival at the tail of the
  // producer, wait at the head of the consumer).  Several handles on ONE device must not do that: a consumer kernel
  // spinning at its head can fill the device and keep the producer it waits for from ever being scheduled; there the
  // ranks are ordered by the one-warp barrier kernel between the launches, which always fits beside other work.
  bool distinct = true;
  for (int i = 0; i < n; ++i)
    for (int j = 0; j < i; ++j) distinct = distinct && hs[i]->device != hs[j]->device;
  const int in_kernel = (distinct && n > 1) ? 1 : 0;
  // Dry pass, one handle after the other and without any cross-rank wait: every kernel of the loop gets loaded into its
  // context and every buffer reaches its final size.  Inside the loop nothing may need a device-wide pause any more
  // -- a lazy kernel load or a cudaMalloc on a device where a peer's kernel is spinning would wait for that kernel,
  // which waits for us.  What the pass writes is overwritten below.
  for (int i = 0; i < n; ++i) {
    ccu_handle* h = hs[i];
    CCU_CUDA(h, cudaSetDevice(h->device));
    if (nbcs > 0) CCU_CUDA(h, cudaMemsetAsync(h->fmx->assign.p, 0xff, nbcs * sizeof(int32_t), h->stream));
    if (ccu_fmx_mstep_fused_dev(h, 1, 1, 0) || ccu_fmx_estep_fused_dev(h, 0))
      return hs[i] == h0 ? 1 : ccu_fail(h0, "ccu_fmx_run_em_multi: handle %d: %s", i, ccu_last_error(h));
    CCU_CUDA(h, ccu::preload_fmx_sync());
    CCU_CUDA(h, cudaStreamSynchronize(h->stream));
  }
  std::vector<int> rcs(n, 0);
  auto work = [&](int i) {
    ccu_handle* h = hs[i];
    FmxState* f = h->fmx;
    auto step = [&](int rc) {
      if (rc && !rcs[i]) rcs[i] = rc;
      return rcs[i] == 0;
    };
    if (cudaSetDevice(h->device) != cudaSuccess) {
      rcs[i] = ccu_fail(h, "cudaSetDevice(%d) failed", h->device);
      return;
    }
    if (nbcs > 0 && cudaMemcpyAsync(f->assign.p, init_assign, nbcs * sizeof(int32_t), cudaMemcpyHostToDevice, h->stream) != cudaSuccess) {
      rcs[i] = ccu_fail(h, "ccu_fmx_run_em_multi: copying the initial assignment failed");
      return;
    }
    auto apply_at = [&](int it) { return geno_error_every_iter ? 1 : (it + 1 == niter ? 1 : 0); };
    // A rank that fails keeps its peers in step (a missing arrival would only be noticed by their time-out)
    auto phase = [&](int rc) {
      if (rcs[i] == 0) step(rc);
      if (!in_kernel || rcs[i] != 0) ccu_fmx_rank_barrier_dev(h);
    };
    phase(ccu_fmx_mstep_fused_dev(h, apply_at(0), 0, in_kernel));
    for (int it = 0; it < niter; ++it) {
      phase(rcs[i] == 0 ? ccu_fmx_estep_fused_dev(h, in_kernel) : 1);
      phase(rcs[i] == 0 ? ccu_fmx_mstep_fused_dev(h, apply_at(it + 1), it + 1 == niter, in_kernel) : 1);
    }
    if (rcs[i] == 0 && cb[i + 1] > cb[i]) step(ccu_fmx_download_results(h, out + cb[i]));
    if (cudaStreamSynchronize(h->stream) != cudaSuccess && rcs[i] == 0) rcs[i] = ccu_fail(h, "stream synchronisation failed");
    int32_t failed = 0;
    if (rcs[i] == 0 && ccu_fmx_rank_barrier_failed(h, &failed) == 0 && failed) rcs[i] = 1;  // message set by the check
  };
  if (n == 1) {
    work(0);
  } else {
    std::vector<std::thread> th;
    for (int i = 0; i < n; ++i) th.emplace_back(work, i);
    for (auto& t : th) t.join();
  }
  for (int i = 0; i < n; ++i)
    if (rcs[i]) {
      if (i != 0) ccu_fail(h0, "ccu_fmx_run_em_multi: handle %d: %s", i, ccu_last_error(hs[i]));
      return rcs[i];
    }
  if (clusters_out) {  // every handle folded its own SNP slab: stitch the slabs together
    const int32_t K = f0->K;
    std::vector<ccu_pileup> tmp((size_t)K * nsnps);
    for (int i = 0; i < n; ++i) {
      if (sb[i + 1] == sb[i]) continue;
      if (int rc = ccu_fmx_get_clusters(hs[i], tmp.data())) return rc;
      for (int32_t k = 0; k < K; ++k)
        memcpy(clusters_out + (size_t)k * nsnps + sb[i], tmp.data() + (size_t)k * nsnps + sb[i],
               (size_t)(sb[i + 1] - sb[i]) * sizeof(ccu_pileup));
    }
  }
  for (int i = 0; i < n; ++i) {
    cudaSetDevice(hs[i]->device);
    hs[i]->fmx->close_peers();
  }
  return 0;
}

int ccu_fmx_pair_distances(ccu_handle* h, int64_t* npairs, int64_t* nrecords, float* ms) {
  CcuRange nvtx_range("ccu_fmx_pair_distances");
  if (int rc = need(h, true, "ccu_fmx_pair_distances")) return rc;
  CCU_CUDA(h, cudaSetDevice(h->device));
  FmxState* f = h->fmx;
  if (int rc = ccu_fmx_pairs_build(h, fmx_view(h), f->pairs, npairs)) return rc;
  if (nrecords) *nrecords = ccu_fmx_pairs_records(f->pairs);
  if (ms) *ms = ccu_fmx_pairs_ms(f->pairs);
  f->tm.launches += 6;
  return 0;
}

int ccu_fmx_get_pair_distances(ccu_handle* h, ccu_pair_dist* out) {
  if (int rc = need(h, true, "ccu_fmx_get_pair_distances")) return rc;
  CCU_CUDA(h, cudaSetDevice(h->device));
  return ccu_fmx_pairs_download(h, h->fmx->pairs, out);
}

int ccu_fmx_get_voting_pairs(ccu_handle* h, double bf_thres, int64_t* nvoting, ccu_pair_dist* out) {
  if (int rc = need(h, true, "ccu_fmx_get_voting_pairs")) return rc;
  CCU_CUDA(h, cudaSetDevice(h->device));
  return ccu_fmx_pairs_download_voting(h, h->fmx->pairs, bf_thres, nvoting, out);
}

int ccu_fmx_selftest_division(ccu_handle* h, int64_t nsets, uint64_t seed, int64_t* mismatches) {
  if (!h) return 1;
  if (!mismatches || nsets < 0) return ccu_fail(h, "ccu_fmx_selftest_division: bad arguments");
  CCU_CUDA(h, cudaSetDevice(h->device));
  unsigned long long* d = nullptr;
  CCU_CUDA(h, cudaMalloc(&d, sizeof(unsigned long long)));
  cudaError_t e = cudaMemsetAsync(d, 0, sizeof(unsigned long long), h->stream);
  if (e == cudaSuccess) e = ccu::launch_fmx_div_selftest(seed, nsets, d, h->stream);
  unsigned long long v = 0;
  if (e == cudaSuccess) e = cudaMemcpyAsync(&v, d, sizeof(v), cudaMemcpyDeviceToHost, h->stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(h->stream);
  cudaFree(d);
  CCU_CUDA(h, e);
  *mismatches = (int64_t)v;
  return 0;
}

int ccu_fmx_get_timings(const ccu_handle* h, ccu_fmx_timings* t) {
  if (!h || !t || !h->fmx) return 1;
  *t = h->fmx->tm;
  return 0;
}

}  // extern "C"
