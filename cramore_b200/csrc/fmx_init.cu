// fmx_init.cu -- freemuxlet's pairwise droplet distances (cmd_cram_freemuxlet.cpp:172-221), the input of
// its initial clustering, on the GPU.
//
// The reference keeps a dense lower-triangular matrix of dropD records (nbcs^2/2 x 32 bytes: 40 GB at 50k
// droplets) and, SNP by SNP, adds log lk0 and log lk2 of every pair of droplets that both cover the SNP.
// Almost all of that matrix stays zero (two droplets with ~200 of 300k SNPs share 0.13 SNPs on average), and a
// pair that shares nothing casts no vote in the clustering, so the engine builds the sparse form:
//
//   1. fmx_pair_count_kernel      records per SNP = n_v (n_v - 1) / 2, exclusive scan -> record offsets
//   2. fmx_pair_emit_kernel       a warp per SNP writes, for every pair of its droplets, the key id1 * nbcs + id2
//                                 (id1 > id2) and the payload (log lk0, log lk2, reads of each droplet); records
//                                 leave in ascending SNP order
//   3. cub::DeviceRadixSort       stable sort of (key, record id): the records of a pair become adjacent and stay
//                                 in ascending SNP order
//   4. cub run-length encode      distinct pairs and their record counts
//   5. fmx_pair_reduce_kernel     a thread per pair adds its records in that order -- the same sequence of
//                                 floating-point additions dropDs[i][j].llk0 += log(lk0) performs
//
// lk0 / lk2 follow the reference's operation order with explicit round-to-nearest intrinsics; the only
// difference left is the last-bit behaviour of log() (device libm vs host libm).
#include <cuda_runtime.h>
#include <stdint.h>

#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_run_length_encode.cuh>
#include <cub/device/device_scan.cuh>
#include <cub/device/device_select.cuh>

#include "ccu_handle.h"
#include "fmx_internal.h"

namespace ccu {

__global__ void fmx_pair_count_kernel(int64_t nsnps, const int64_t* __restrict__ csc_ptr, int64_t* __restrict__ cnt) {
  const int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (v >= nsnps) return;
  const int64_t n = csc_ptr[v + 1] - csc_ptr[v];
  cnt[v] = n * (n - 1) / 2;
}

// one warp per SNP; pair p of the SNP is (a, b), a > b, p = a (a - 1) / 2 + b: the order of
// `for it: for jt < it` (cmd_cram_freemuxlet.cpp:191-193)
__global__ void __launch_bounds__(256)
    fmx_pair_emit_kernel(FmxDev d, const int64_t* __restrict__ rec_ptr, uint64_t* __restrict__ keys,
                         uint32_t* __restrict__ rec_id, double2* __restrict__ rec_llk, int2* __restrict__ rec_reads) {
  const int lane = threadIdx.x & 31;
  const int64_t v = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (v >= d.nsnps) return;
  const int64_t b0 = d.csc_ptr[v], n = d.csc_ptr[v + 1] - b0;
  const int64_t np = n * (n - 1) / 2, r0 = rec_ptr[v];
  if (np == 0) return;
  const double af = d.af[v];
  double gps[3];  // :197-200
  gps[0] = __dmul_rn(1.0 - af, 1.0 - af);
  gps[1] = __dmul_rn(__dmul_rn(2.0, af), 1.0 - af);
  gps[2] = __dmul_rn(af, af);
  for (int64_t p = lane; p < np; p += 32) {
    int64_t a = (int64_t)((1.0 + sqrt(1.0 + 8.0 * (double)p)) * 0.5);
    while (a * (a - 1) / 2 > p) --a;
    while ((a + 1) * a / 2 <= p) ++a;
    const int64_t b = p - a * (a - 1) / 2;
    const int64_t ia = b0 + a, ib = b0 + b;
    const double* ra = d.csc_row + ia * kCscRow;
    const double* rb = d.csc_row + ib * kCscRow;
    const double gi[3] = {ra[0], ra[4], ra[8]};  // gls[0], gls[4], gls[8]
    const double gj[3] = {rb[0], rb[4], rb[8]};
    double lk0 = 0.0, lk2 = 0.0;
#pragma unroll
    for (int x = 0; x < 3; ++x) {  // :202-207
      const double gix = gi[x];
      lk2 = __dadd_rn(lk2, __dmul_rn(__dmul_rn(gix, gj[x]), gps[x]));
#pragma unroll
      for (int y = 0; y < 3; ++y)
        lk0 = __dadd_rn(lk0, __dmul_rn(__dmul_rn(__dmul_rn(gix, gj[y]), gps[x]), gps[y]));
    }
    const int64_t r = r0 + p;
    keys[r] = (uint64_t)d.csc_cell[ia] * (uint64_t)d.nbcs + (uint64_t)d.csc_cell[ib];  // droplets ascend inside a SNP
    rec_id[r] = (uint32_t)r;
    rec_llk[r] = make_double2(log(lk0), log(lk2));
    // nreads of the two entries (low 21 bits of the packed counts)
    rec_reads[r] = make_int2((int)((unsigned long long)__double_as_longlong(ra[9]) & kCntMask),
                             (int)((unsigned long long)__double_as_longlong(rb[9]) & kCntMask));
  }
}

__global__ void fmx_pair_reduce_kernel(int64_t npairs, int64_t nbcs, const uint64_t* __restrict__ ukeys,
                                       const int64_t* __restrict__ run_ptr, const uint32_t* __restrict__ sorted_id,
                                       const double2* __restrict__ rec_llk, const int2* __restrict__ rec_reads,
                                       ccu_pair_dist* __restrict__ out) {
  const int64_t u = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (u >= npairs) return;
  ccu_pair_dist o;
  o.id1 = (int32_t)(ukeys[u] / (uint64_t)nbcs);
  o.id2 = (int32_t)(ukeys[u] % (uint64_t)nbcs);
  o.nsnps = 0;
  o.nread1 = o.nread2 = 0;
  o._pad = 0;
  o.llk0 = o.llk2 = 0.0;
  for (int64_t t = run_ptr[u]; t < run_ptr[u + 1]; ++t) {  // ascending SNP: :209-214
    const uint32_t r = sorted_id[t];
    const double2 l = rec_llk[r];
    const int2 rd = rec_reads[r];
    ++o.nsnps;
    o.nread1 += rd.x;
    o.nread2 += rd.y;
    o.llk2 = __dadd_rn(o.llk2, l.y);
    o.llk0 = __dadd_rn(o.llk0, l.x);
  }
  out[u] = o;
}

struct PairVotes {  // a pair casts a vote iff its Bayes factor passes the threshold in either direction
  double thres;
  __host__ __device__ bool operator()(const ccu_pair_dist& p) const {
    const double bf = p.llk2 - p.llk0;
    return (bf > thres) || (-bf > thres);
  }
};

}  // namespace ccu

namespace {

struct PairState {
  DevBuf cnt, rec_ptr, keys, keys_alt, ids, ids_alt, llk, reads, ukeys, ucnt, run_ptr, nruns, temp, out, sel;
  int64_t npairs = -1, nrecords = 0;
  float ms = 0;
  cudaEvent_t e0 = nullptr, e1 = nullptr;
  void release() {
    if (e0) cudaEventDestroy(e0);
    if (e1) cudaEventDestroy(e1);
    e0 = e1 = nullptr;
    DevBuf* all[] = {&cnt, &rec_ptr, &keys, &keys_alt, &ids, &ids_alt, &llk, &reads, &ukeys, &ucnt, &run_ptr, &nruns,
                     &temp, &out, &sel};
    for (DevBuf* b : all) b->release();
    npairs = -1;
  }
};

}  // namespace

// owned by the handle through an opaque pointer kept in ccu_fmx_api.cu
struct FmxPairs : PairState {};

void ccu_fmx_pairs_release(FmxPairs*& p) {
  if (p) {
    p->release();
    delete p;
    p = nullptr;
  }
}

int ccu_fmx_pairs_build(ccu_handle* h, const ccu::FmxDev& d, FmxPairs*& ps, int64_t* npairs_out) {
  if (!ps) ps = new FmxPairs();
  cudaStream_t st = h->stream;
  const int64_t nsnps = d.nsnps;
  if ((double)d.nbcs * (double)d.nbcs >= 1.8e19) return ccu_fail(h, "ccu_fmx_pair_distances: too many droplets");
  if (!ps->e0) CCU_CUDA(h, cudaEventCreate(&ps->e0));
  if (!ps->e1) CCU_CUDA(h, cudaEventCreate(&ps->e1));
  cudaEvent_t e0 = ps->e0, e1 = ps->e1;
  ps->npairs = -1;
  CCU_CUDA(h, cudaEventRecord(e0, st));
  CCU_CUDA(h, ps->cnt.reserve((nsnps + 1) * sizeof(int64_t)));
  CCU_CUDA(h, ps->rec_ptr.reserve((nsnps + 1) * sizeof(int64_t)));
  CCU_CUDA(h, cudaMemsetAsync(ps->cnt.p, 0, (nsnps + 1) * sizeof(int64_t), st));
  if (nsnps > 0)
    ccu::fmx_pair_count_kernel<<<(unsigned)((nsnps + 255) / 256), 256, 0, st>>>(nsnps, d.csc_ptr, ps->cnt.as<int64_t>());
  size_t tb = 0;
  CCU_CUDA(h, cub::DeviceScan::ExclusiveSum(nullptr, tb, ps->cnt.as<int64_t>(), ps->rec_ptr.as<int64_t>(), (int)(nsnps + 1), st));
  CCU_CUDA(h, ps->temp.reserve(tb + 16));
  CCU_CUDA(h, cub::DeviceScan::ExclusiveSum(ps->temp.p, tb, ps->cnt.as<int64_t>(), ps->rec_ptr.as<int64_t>(), (int)(nsnps + 1), st));
  int64_t nrec = 0;
  CCU_CUDA(h, cudaMemcpyAsync(&nrec, ps->rec_ptr.as<int64_t>() + nsnps, sizeof(int64_t), cudaMemcpyDeviceToHost, st));
  CCU_CUDA(h, cudaStreamSynchronize(st));
  ps->nrecords = nrec;
  if (nrec > 2147483647LL)
    return ccu_fail(h, "ccu_fmx_pair_distances: %lld (droplet pair, SNP) records exceed 2^31-1", (long long)nrec);
  int64_t nruns = 0;
  if (nrec > 0) {
    CCU_CUDA(h, ps->keys.reserve(nrec * sizeof(uint64_t)));
    CCU_CUDA(h, ps->keys_alt.reserve(nrec * sizeof(uint64_t)));
    CCU_CUDA(h, ps->ids.reserve(nrec * sizeof(uint32_t)));
    CCU_CUDA(h, ps->ids_alt.reserve(nrec * sizeof(uint32_t)));
    CCU_CUDA(h, ps->llk.reserve(nrec * sizeof(double2)));
    CCU_CUDA(h, ps->reads.reserve(nrec * sizeof(int2)));
    ccu::fmx_pair_emit_kernel<<<(unsigned)((nsnps * 32 + 255) / 256), 256, 0, st>>>(
        d, ps->rec_ptr.as<int64_t>(), ps->keys.as<uint64_t>(), ps->ids.as<uint32_t>(), ps->llk.as<double2>(),
        ps->reads.as<int2>());
    CCU_CUDA(h, cudaGetLastError());
    int end_bit = 1;
    while (end_bit < 64 && ((double)d.nbcs * (double)d.nbcs) > ldexp(1.0, end_bit)) ++end_bit;
    cub::DoubleBuffer<uint64_t> kb(ps->keys.as<uint64_t>(), ps->keys_alt.as<uint64_t>());
    cub::DoubleBuffer<uint32_t> vb(ps->ids.as<uint32_t>(), ps->ids_alt.as<uint32_t>());
    tb = 0;
    CCU_CUDA(h, cub::DeviceRadixSort::SortPairs(nullptr, tb, kb, vb, (int)nrec, 0, end_bit, st));
    CCU_CUDA(h, ps->temp.reserve(tb + 16));
    CCU_CUDA(h, cub::DeviceRadixSort::SortPairs(ps->temp.p, tb, kb, vb, (int)nrec, 0, end_bit, st));
    // distinct pairs
    CCU_CUDA(h, ps->ukeys.reserve(nrec * sizeof(uint64_t)));
    CCU_CUDA(h, ps->ucnt.reserve((nrec + 1) * sizeof(int64_t)));
    CCU_CUDA(h, ps->run_ptr.reserve((nrec + 1) * sizeof(int64_t)));
    CCU_CUDA(h, ps->nruns.reserve(sizeof(int64_t)));
    tb = 0;
    CCU_CUDA(h, cub::DeviceRunLengthEncode::Encode(nullptr, tb, kb.Current(), ps->ukeys.as<uint64_t>(),
                                                   ps->ucnt.as<int64_t>(), ps->nruns.as<int64_t>(), (int)nrec, st));
    CCU_CUDA(h, ps->temp.reserve(tb + 16));
    CCU_CUDA(h, cub::DeviceRunLengthEncode::Encode(ps->temp.p, tb, kb.Current(), ps->ukeys.as<uint64_t>(),
                                                   ps->ucnt.as<int64_t>(), ps->nruns.as<int64_t>(), (int)nrec, st));
    CCU_CUDA(h, cudaMemcpyAsync(&nruns, ps->nruns.p, sizeof(int64_t), cudaMemcpyDeviceToHost, st));
    CCU_CUDA(h, cudaStreamSynchronize(st));
    CCU_CUDA(h, cudaMemsetAsync(ps->ucnt.as<int64_t>() + nruns, 0, sizeof(int64_t), st));
    tb = 0;
    CCU_CUDA(h, cub::DeviceScan::ExclusiveSum(nullptr, tb, ps->ucnt.as<int64_t>(), ps->run_ptr.as<int64_t>(), (int)(nruns + 1), st));
    CCU_CUDA(h, ps->temp.reserve(tb + 16));
    CCU_CUDA(h, cub::DeviceScan::ExclusiveSum(ps->temp.p, tb, ps->ucnt.as<int64_t>(), ps->run_ptr.as<int64_t>(), (int)(nruns + 1), st));
    CCU_CUDA(h, ps->out.reserve(nruns * sizeof(ccu_pair_dist) + 16));
    ccu::fmx_pair_reduce_kernel<<<(unsigned)((nruns + 255) / 256), 256, 0, st>>>(
        nruns, d.nbcs, ps->ukeys.as<uint64_t>(), ps->run_ptr.as<int64_t>(), vb.Current(), ps->llk.as<double2>(),
        ps->reads.as<int2>(), ps->out.as<ccu_pair_dist>());
    CCU_CUDA(h, cudaGetLastError());
  }
  CCU_CUDA(h, cudaEventRecord(e1, st));
  CCU_CUDA(h, cudaEventSynchronize(e1));
  cudaEventElapsedTime(&ps->ms, e0, e1);
  ps->npairs = nruns;
  if (npairs_out) *npairs_out = nruns;
  return 0;
}

int ccu_fmx_pairs_download(ccu_handle* h, FmxPairs* ps, ccu_pair_dist* out) {
  if (!ps || ps->npairs < 0) return ccu_fail(h, "ccu_fmx_get_pair_distances: call ccu_fmx_pair_distances first");
  if (ps->npairs == 0) return 0;
  if (!out) return ccu_fail(h, "ccu_fmx_get_pair_distances: NULL output");
  CCU_CUDA(h, cudaMemcpyAsync(out, ps->out.p, ps->npairs * sizeof(ccu_pair_dist), cudaMemcpyDeviceToHost, h->stream));
  CCU_CUDA(h, cudaStreamSynchronize(h->stream));
  return 0;
}

float ccu_fmx_pairs_ms(const FmxPairs* ps) { return ps ? ps->ms : 0.f; }
int64_t ccu_fmx_pairs_records(const FmxPairs* ps) { return ps ? ps->nrecords : 0; }

// the pairs whose |llk2 - llk0| passes bf_thres (the only ones the votes of :262-278 / :311-320 look at),
// compacted on the device in (id1, id2) order; out == NULL only counts them
int ccu_fmx_pairs_download_voting(ccu_handle* h, FmxPairs* ps, double bf_thres, int64_t* nvoting, ccu_pair_dist* out) {
  if (!ps || ps->npairs < 0) return ccu_fail(h, "ccu_fmx_get_voting_pairs: call ccu_fmx_pair_distances first");
  if (!nvoting) return ccu_fail(h, "ccu_fmx_get_voting_pairs: NULL count");
  *nvoting = 0;
  if (ps->npairs == 0) return 0;
  cudaStream_t st = h->stream;
  CCU_CUDA(h, ps->sel.reserve(ps->npairs * sizeof(ccu_pair_dist) + 16));
  CCU_CUDA(h, ps->nruns.reserve(sizeof(int64_t)));
  size_t tb = 0;
  ccu::PairVotes pred{bf_thres};
  CCU_CUDA(h, cub::DeviceSelect::If(nullptr, tb, ps->out.as<ccu_pair_dist>(), ps->sel.as<ccu_pair_dist>(),
                                    ps->nruns.as<int64_t>(), (int)ps->npairs, pred, st));
  CCU_CUDA(h, ps->temp.reserve(tb + 16));
  CCU_CUDA(h, cub::DeviceSelect::If(ps->temp.p, tb, ps->out.as<ccu_pair_dist>(), ps->sel.as<ccu_pair_dist>(),
                                    ps->nruns.as<int64_t>(), (int)ps->npairs, pred, st));
  int64_t n = 0;
  CCU_CUDA(h, cudaMemcpyAsync(&n, ps->nruns.p, sizeof(int64_t), cudaMemcpyDeviceToHost, st));
  CCU_CUDA(h, cudaStreamSynchronize(st));
  *nvoting = n;
  if (out && n > 0) {
    CCU_CUDA(h, cudaMemcpyAsync(out, ps->sel.p, n * sizeof(ccu_pair_dist), cudaMemcpyDeviceToHost, st));
    CCU_CUDA(h, cudaStreamSynchronize(st));
  }
  return 0;
}
