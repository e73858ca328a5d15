// ccu_api.cu -- host side of libcramore_cuda.so: the C ABI declared in include/cramore_cuda.h.
// Owns device memory, the stream and the event timers; every kernel lives in demux_kernels.cu /
// fmx_kernels.cu.  There is deliberately no CPU implementation behind any entry point: if no
// sm_100 device is usable, ccu_create() fails and nothing else can be called.
#include <cuda_runtime.h>

#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <limits>
#include <string>
#include <chrono>
#include <vector>

#include "ccu_handle.h"
#include "cramore_cuda.h"
#include "demux_internal.h"

namespace {

std::string g_create_error;

}  // namespace

int ccu_fail(ccu_handle* h, const char* fmt, ...) {
  char buf[1024];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  if (h)
    h->err = buf;
  else
    g_create_error = buf;
  return 1;
}

namespace {

#define fail ccu_fail
#define ev_ms ccu_ev_ms

ccu::DemuxDev dev_view(const ccu_handle* h) {
  ccu::DemuxDev d;
  memset(&d, 0, sizeof(d));
  d.nv = h->nv;
  d.nvp = h->nvp;
  d.nA = h->nA;
  d.doublet_prior = h->doublet_prior;
  d.alpha = h->d_alpha.as<double>();
  for (int n = 0; n < CCU_MAX_ALPHA; ++n) d.alpha_c[n] = (n < (int)h->alpha.size()) ? h->alpha[n] : 0.0;
  d.ptab = h->d_ptab.as<double>();
  d.is_half = h->d_half.as<uint8_t>();
  d.phred = h->d_phred.as<double>();
  d.nsnps = h->nsnps;
  d.G = h->d_G.as<double>();
  d.has_gp = h->has_gp_flags ? h->d_hasgp.as<uint8_t>() : nullptr;
  d.nbcs = h->nbcs;
  d.nnz = h->nnz;
  d.nreads = h->nreads;
  d.ngen = h->ngen;
  d.cell_ptr = h->d_cell_ptr.as<int64_t>();
  d.snp_idx = h->d_snp_idx.as<int32_t>();
  d.read_ptr = h->d_read_ptr.as<int64_t>();
  d.al = h->d_al.as<uint8_t>();
  d.bq = h->d_bq.as<uint8_t>();
  d.tiles = h->d_tiles.as<double>();
  d.tslot = h->d_tslot.as<int32_t>();
  d.vaff = h->d_vaff.as<double>();
  d.vent = h->d_vent.as<int64_t>();
  d.vsnp = h->d_vsnp.as<int32_t>();
  d.eclass = h->d_eclass.as<uint8_t>();
  d.vcnt_g = h->d_vcnt.as<int32_t>();
  d.vcnt_a = h->d_vcnt.as<int32_t>() + (h->nbcs > 0 ? h->nbcs : 0);
  d.sng_llk = h->d_sng.as<double>();
  d.fmat = h->d_fmat.as<double>();
  d.pm = h->d_pm.as<double>();
  d.pe = h->d_pe.as<int32_t>();
  d.maxbq = h->d_maxbq.as<int32_t>();
  d.partials = h->d_partials.as<ccu::ItemPartial>();
  d.results = h->d_results.as<ccu_demux_result>();
  return d;
}

}  // namespace

extern "C" {

int ccu_create(int32_t device, ccu_handle** out) {
  if (!out) return fail(nullptr, "ccu_create: out is NULL");
  *out = nullptr;
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0)
    return fail(nullptr, "ccu_create: no CUDA device available (%s); this engine has no CPU fallback",
                e == cudaSuccess ? "device count 0" : cudaGetErrorString(e));
  if (device < 0 || device >= ndev) return fail(nullptr, "ccu_create: device %d out of range [0,%d)", device, ndev);
  cudaDeviceProp prop;
  if ((e = cudaGetDeviceProperties(&prop, device)) != cudaSuccess)
    return fail(nullptr, "ccu_create: cudaGetDeviceProperties: %s", cudaGetErrorString(e));
  if (prop.major != 10)
    return fail(nullptr, "ccu_create: device %d is sm_%d%d; libcramore_cuda is built for sm_100a only", device,
                prop.major, prop.minor);
  if ((e = cudaSetDevice(device)) != cudaSuccess)
    return fail(nullptr, "ccu_create: cudaSetDevice: %s", cudaGetErrorString(e));
  ccu_handle* h = new ccu_handle();
  h->device = device;
  h->num_sms = prop.multiProcessorCount;
  if ((e = cudaStreamCreateWithFlags(&h->own_stream, cudaStreamNonBlocking)) != cudaSuccess) {
    delete h;
    return fail(nullptr, "ccu_create: cudaStreamCreate: %s", cudaGetErrorString(e));
  }
  h->stream = h->own_stream;
  for (int i = 0; i < EV_COUNT; ++i) cudaEventCreate(&h->ev[i]);

  // Phred tables, PhredHelper.cpp:29-33, computed with the host libm exactly like the reference
  std::vector<double> ph(512);
  for (int q = 0; q < 256; ++q) {
    ph[q] = (q > 1) ? pow(0.1, q * 0.1) : 0.75;
    ph[256 + q] = 1. - ph[q];
  }
  if (h->d_phred.reserve(512 * sizeof(double)) != cudaSuccess ||
      cudaMemcpy(h->d_phred.p, ph.data(), 512 * sizeof(double), cudaMemcpyHostToDevice) != cudaSuccess ||
      h->d_sink.reserve(64) != cudaSuccess) {
    ccu_destroy(h);
    return fail(nullptr, "ccu_create: device allocation failed");
  }
  *out = h;
  return 0;
}

int ccu_destroy(ccu_handle* h) {
  if (!h) return 0;
  cudaSetDevice(h->device);
  cudaStreamSynchronize(h->stream);
  DevBuf* bufs[] = {&h->d_alpha, &h->d_ptab, &h->d_half, &h->d_phred, &h->d_gpf, &h->d_err, &h->d_hasgp,
                    &h->d_avg, &h->d_G, &h->d_cell_ptr, &h->d_snp_idx, &h->d_read_ptr, &h->d_al, &h->d_bq,
                    &h->d_tiles, &h->d_vaff, &h->d_eclass, &h->d_vent, &h->d_vsnp, &h->d_vcnt, &h->d_sng, &h->d_fmat, &h->d_pm, &h->d_pe, &h->d_maxbq, &h->d_partials, &h->d_results, &h->d_dbg, &h->d_flag,
                    &h->d_sink, &h->d_tslot, &h->d_scan, &h->d_ngen};
  for (DevBuf* b : bufs) b->release();
  for (DevBuf& b : h->d_in2) b.release();
  if (h->aux_stream) cudaStreamDestroy(h->aux_stream);
  ccu_fmx_release(h);
  for (int i = 0; i < EV_COUNT; ++i)
    if (h->ev[i]) cudaEventDestroy(h->ev[i]);
  if (h->own_stream) cudaStreamDestroy(h->own_stream);
  delete h;
  return 0;
}

const char* ccu_last_error(const ccu_handle* h) { return h ? h->err.c_str() : g_create_error.c_str(); }

int ccu_set_stream(ccu_handle* h, void* cuda_stream) {
  if (!h) return 1;
  h->stream = cuda_stream ? reinterpret_cast<cudaStream_t>(cuda_stream) : h->own_stream;
  return 0;
}

int ccu_synchronize(ccu_handle* h) {
  if (!h) return 1;
  CCU_CUDA(h, cudaSetDevice(h->device));
  CCU_CUDA(h, cudaStreamSynchronize(h->stream));
  return 0;
}

int ccu_get_timings(const ccu_handle* h, ccu_timings* t) {
  if (!h || !t) return 1;
  *t = h->tm;
  return 0;
}

int ccu_demux_configure(ccu_handle* h, int32_t nv, int32_t nalpha, const double* alpha, double doublet_prior) {
  if (!h) return 1;
  if (nv < 2) return fail(h, "ccu_demux_configure: need at least 2 samples (nv=%d)", nv);
  if (nalpha < 2 || nalpha > CCU_MAX_ALPHA)
    return fail(h, "ccu_demux_configure: alpha grid size %d outside [2,%d]", nalpha, CCU_MAX_ALPHA);
  if (!alpha || alpha[0] != 0.0)
    return fail(h, "ccu_demux_configure: the first --alpha must be 0 (singlets are read from it)");
  if ((double)nv * nv * nalpha >= 2147483647.0) return fail(h, "ccu_demux_configure: nv*nv*nalpha overflows int32");
  CCU_CUDA(h, cudaSetDevice(h->device));
  h->nv = nv;
  h->scratch_fits = 0.0;
  h->nvp = (nv + 31) / 32 * 32;
  h->nA = nalpha;
  h->doublet_prior = doublet_prior;
  h->alpha.assign(alpha, alpha + nalpha);
  // allele-fraction table, cmd_cram_demuxlet.cpp:673: p = 0.5*l + (m-l)*0.5*alpha ; and 1.0-p (:685)
  std::vector<double> ptab(nalpha * 18);
  std::vector<uint8_t> half(nalpha);
  for (int k = 0; k < nalpha; ++k) {
    half[k] = (alpha[k] == 0.5) ? 1 : 0;
    for (int l = 0; l < 3; ++l)
      for (int m = 0; m < 3; ++m) {
        volatile double t1 = 0.5 * l;
        volatile double t2 = (m - l) * 0.5;
        volatile double t3 = t2 * alpha[k];
        volatile double p = t1 + t3;
        volatile double omp = 1.0 - p;
        ptab[k * 9 + l * 3 + m] = p;
        ptab[nalpha * 9 + k * 9 + l * 3 + m] = omp;
      }
  }
  CCU_CUDA(h, h->d_alpha.reserve(nalpha * sizeof(double)));
  CCU_CUDA(h, h->d_ptab.reserve(ptab.size() * sizeof(double)));
  CCU_CUDA(h, h->d_half.reserve(nalpha));
  CCU_CUDA(h, cudaMemcpyAsync(h->d_alpha.p, alpha, nalpha * sizeof(double), cudaMemcpyHostToDevice, h->stream));
  CCU_CUDA(h, cudaMemcpyAsync(h->d_ptab.p, ptab.data(), ptab.size() * sizeof(double), cudaMemcpyHostToDevice,
                              h->stream));
  CCU_CUDA(h, cudaMemcpyAsync(h->d_half.p, half.data(), nalpha, cudaMemcpyHostToDevice, h->stream));
  CCU_CUDA(h, cudaStreamSynchronize(h->stream));
  h->configured = true;
  h->nsnps = -1;
  h->nbcs = -1;
  h->ran = false;
  return 0;
}

// Device-side staging of the genotype inputs (include/cramore_cuda.h): reserve the float matrix, the per-SNP error
// and flag arrays and hand their device addresses to the caller, who fills them on the engine's stream -- a plain
// host->device copy, or (multi-GPU) a copy of one SNP slab per rank completed by an all-gather over NVLink.
int ccu_demux_genotypes_stage(ccu_handle* h, int64_t nsnps, void** gp_dev, void** snp_err_dev, void** has_gp_dev) {
  if (!h) return 1;
  if (!h->configured) return fail(h, "ccu_demux_genotypes_stage: call ccu_demux_configure first");
  if (nsnps < 0) return fail(h, "ccu_demux_genotypes_stage: bad arguments");
  if (nsnps > 2147483647LL) return fail(h, "ccu_demux_genotypes_stage: more than 2^31-1 SNPs");
  CCU_CUDA(h, cudaSetDevice(h->device));
  const size_t nf = (size_t)nsnps * h->nv * 3;
  CCU_CUDA(h, h->d_gpf.reserve(nf * sizeof(float) + 16));
  CCU_CUDA(h, h->d_err.reserve(nsnps * sizeof(double) + 16));
  CCU_CUDA(h, h->d_avg.reserve(nsnps * 4 * sizeof(double) + 16));
  CCU_CUDA(h, h->d_G.reserve((size_t)nsnps * h->nvp * 3 * sizeof(double) + 16));
  CCU_CUDA(h, h->d_hasgp.reserve(nsnps + 16));
  if (gp_dev) *gp_dev = h->d_gpf.p;
  if (snp_err_dev) *snp_err_dev = h->d_err.p;
  if (has_gp_dev) *has_gp_dev = h->d_hasgp.p;
  h->nsnps = -1;  // nothing usable until the commit
  h->staged_snps = nsnps;
  return 0;
}

int ccu_demux_genotypes_commit(ccu_handle* h, int64_t nsnps, int32_t use_has_gp) {
  CcuRange nvtx_range("ccu_demux_genotypes_commit");
  if (!h) return 1;
  if (!h->configured) return fail(h, "ccu_demux_genotypes_commit: call ccu_demux_configure first");
  if (nsnps < 0 || nsnps != h->staged_snps)
    return fail(h, "ccu_demux_genotypes_commit: %lld SNPs were not staged by ccu_demux_genotypes_stage", (long long)nsnps);
  CCU_CUDA(h, cudaSetDevice(h->device));
  h->has_gp_flags = use_has_gp != 0;
  CCU_CUDA(h, cudaEventRecord(h->ev[EV_GP0], h->stream));
  CCU_CUDA(h, ccu::launch_gp_prepare(h->d_gpf.as<float>(), h->d_err.as<double>(),
                                     use_has_gp ? h->d_hasgp.as<uint8_t>() : nullptr, nsnps, h->nv, h->nvp,
                                     h->d_avg.as<double>(), h->d_G.as<double>(), h->stream));
  CCU_CUDA(h, cudaEventRecord(h->ev[EV_GP1], h->stream));
  CCU_CUDA(h, cudaStreamSynchronize(h->stream));
  h->tm.ms_genotypes = ev_ms(h, EV_GP0, EV_GP1);
  h->nsnps = nsnps;
  h->have_avg = true;
  return 0;
}

int ccu_demux_set_genotypes(ccu_handle* h, int64_t nsnps, const float* gp, const double* snp_err,
                            const uint8_t* has_gp) {
  CcuRange nvtx_range("ccu_demux_set_genotypes");
  if (!h) return 1;
  if (!h->configured) return fail(h, "ccu_demux_set_genotypes: call ccu_demux_configure first");
  if (nsnps < 0 || (nsnps > 0 && (!gp || !snp_err))) return fail(h, "ccu_demux_set_genotypes: bad arguments");
  if (int rc = ccu_demux_genotypes_stage(h, nsnps, nullptr, nullptr, nullptr)) return rc;
  const size_t nf = (size_t)nsnps * h->nv * 3;
  CCU_CUDA(h, cudaEventRecord(h->ev[EV_X0], h->stream));
  if (nsnps > 0) {
    CCU_CUDA(h, cudaMemcpyAsync(h->d_gpf.p, gp, nf * sizeof(float), cudaMemcpyHostToDevice, h->stream));
    CCU_CUDA(h, cudaMemcpyAsync(h->d_err.p, snp_err, nsnps * sizeof(double), cudaMemcpyHostToDevice, h->stream));
    if (has_gp) CCU_CUDA(h, cudaMemcpyAsync(h->d_hasgp.p, has_gp, nsnps, cudaMemcpyHostToDevice, h->stream));
  }
  return ccu_demux_genotypes_commit(h, nsnps, has_gp != nullptr);
}

// The reference keeps sc_snp_t::gps (sc_drop_seq.h:115-127): nv*3 doubles per SNP, genotype error already mixed
// in (sc_drop_seq.cpp:287-315), or NULL where the VCF had no usable record.  A patched cmdCramDemuxlet can hand
// those pointers over as they are.
int ccu_demux_set_genotypes_mixed(ccu_handle* h, int64_t nsnps, const double* const* gps) {
  CcuRange nvtx_range("ccu_demux_set_genotypes_mixed");
  if (!h) return 1;
  if (!h->configured) return fail(h, "ccu_demux_set_genotypes_mixed: call ccu_demux_configure first");
  if (nsnps < 0 || (nsnps > 0 && !gps)) return fail(h, "ccu_demux_set_genotypes_mixed: bad arguments");
  if (nsnps > 2147483647LL) return fail(h, "ccu_demux_set_genotypes_mixed: more than 2^31-1 SNPs");
  CCU_CUDA(h, cudaSetDevice(h->device));
  const size_t row = (size_t)h->nvp * 3;
  CCU_CUDA(h, h->d_G.reserve((size_t)nsnps * row * sizeof(double) + 16));
  CCU_CUDA(h, h->d_hasgp.reserve(nsnps + 16));
  h->has_gp_flags = true;
  CCU_CUDA(h, cudaEventRecord(h->ev[EV_X0], h->stream));
  CCU_CUDA(h, cudaEventRecord(h->ev[EV_GP0], h->stream));
  // staged through a bounded host buffer: rows padded to nvp samples, zero rows where gps[s] == NULL
  const int64_t chunk = std::max<int64_t>(1, (int64_t)((64u << 20) / (row * sizeof(double))));
  std::vector<double> stage((size_t)std::min(chunk, std::max<int64_t>(nsnps, 1)) * row);
  std::vector<uint8_t> flags((size_t)nsnps);
  for (int64_t b = 0; b < nsnps; b += chunk) {
    const int64_t e = std::min(nsnps, b + chunk);
    std::fill(stage.begin(), stage.begin() + (size_t)(e - b) * row, 0.0);
    for (int64_t s = b; s < e; ++s) {
      flags[(size_t)s] = gps[s] != nullptr;
      if (gps[s]) memcpy(&stage[(size_t)(s - b) * row], gps[s], (size_t)h->nv * 3 * sizeof(double));
    }
    CCU_CUDA(h, cudaMemcpyAsync(h->d_G.as<double>() + (size_t)b * row, stage.data(), (size_t)(e - b) * row * sizeof(double),
                                cudaMemcpyHostToDevice, h->stream));
    CCU_CUDA(h, cudaStreamSynchronize(h->stream));  // the staging buffer is reused
  }
  if (nsnps > 0) CCU_CUDA(h, cudaMemcpyAsync(h->d_hasgp.p, flags.data(), nsnps, cudaMemcpyHostToDevice, h->stream));
  CCU_CUDA(h, cudaEventRecord(h->ev[EV_GP1], h->stream));
  CCU_CUDA(h, cudaStreamSynchronize(h->stream));
  h->tm.ms_genotypes = ev_ms(h, EV_GP0, EV_GP1);
  h->nsnps = nsnps;
  h->have_avg = false;
  h->staged_snps = -1;
  return 0;
}

// One process driving several GPUs: the genotype matrix is uploaded and mixed once (on `src`) and every other handle
// takes its copy over NVLink (cudaMemcpyPeerAsync on its own stream), instead of one PCIe upload per device.
int ccu_demux_copy_genotypes(ccu_handle* dst, ccu_handle* src) {
  CcuRange nvtx_range("ccu_demux_copy_genotypes");
  if (!dst || !src) return 1;
  if (dst == src) return 0;
  if (!dst->configured || !src->configured || src->nsnps < 0)
    return fail(dst, "ccu_demux_copy_genotypes: configure both handles and set the genotypes of the source first");
  if (dst->nv != src->nv) return fail(dst, "ccu_demux_copy_genotypes: the handles are configured for %d and %d samples", dst->nv, src->nv);
  CCU_CUDA(dst, cudaSetDevice(src->device));
  CCU_CUDA(dst, cudaStreamSynchronize(src->stream));  // the source matrix is complete
  CCU_CUDA(dst, cudaSetDevice(dst->device));
  const int64_t nsnps = src->nsnps;
  const size_t gbytes = (size_t)nsnps * src->nvp * 3 * sizeof(double);
  CCU_CUDA(dst, dst->d_G.reserve(gbytes + 16));
  CCU_CUDA(dst, dst->d_hasgp.reserve(nsnps + 16));
  CCU_CUDA(dst, cudaEventRecord(dst->ev[EV_GP0], dst->stream));
  if (nsnps > 0) {
    CCU_CUDA(dst, cudaMemcpyPeerAsync(dst->d_G.p, dst->device, src->d_G.p, src->device, gbytes, dst->stream));
    if (src->has_gp_flags)
      CCU_CUDA(dst, cudaMemcpyPeerAsync(dst->d_hasgp.p, dst->device, src->d_hasgp.p, src->device, (size_t)nsnps, dst->stream));
  }
  CCU_CUDA(dst, cudaEventRecord(dst->ev[EV_GP1], dst->stream));
  CCU_CUDA(dst, cudaStreamSynchronize(dst->stream));
  dst->tm.ms_genotypes = ev_ms(dst, EV_GP0, EV_GP1);
  dst->has_gp_flags = src->has_gp_flags;
  dst->have_avg = false;
  dst->staged_snps = -1;
  dst->nsnps = nsnps;
  return 0;
}

// Device buffers of one droplet store (inputs and scratch; the tile buffer is sized after the upload, when the
// number of entries with two or more reads is known).  No-op for a store no larger than one reserved before.
static int demux_reserve(ccu_handle* h, int64_t nbcs, int64_t nnz, int64_t nreads) {
  const ccu::ScorePlan plan = ccu::make_score_plan(h->nv, h->nA, h->alpha.data());
  CCU_CUDA(h, h->d_cell_ptr.reserve((nbcs + 1) * sizeof(int64_t)));
  CCU_CUDA(h, h->d_snp_idx.reserve(nnz * sizeof(int32_t) + 16));
  CCU_CUDA(h, h->d_read_ptr.reserve((nnz + 1) * sizeof(int64_t)));
  CCU_CUDA(h, h->d_al.reserve(nreads + 16));
  CCU_CUDA(h, h->d_bq.reserve(nreads + 16));
  if (nnz > 2147483647LL) return fail(h, "ccu_demux_upload: more than 2^31-1 (cell,SNP) entries in one call (ccu_demux_score splits a store into ranges)");
  CCU_CUDA(h, h->d_tslot.reserve((size_t)nnz * sizeof(int32_t) + 16));
  CCU_CUDA(h, h->d_ngen.reserve(16));
  {
    size_t need = 0;
    ccu::DemuxDev q;
    memset(&q, 0, sizeof(q));
    q.nnz = nnz;
    CCU_CUDA(h, ccu::launch_tile(q, false, nullptr, 0, &need, h->stream));
    CCU_CUDA(h, h->d_scan.reserve(need + 16));
  }
  CCU_CUDA(h, h->d_vent.reserve(nnz * sizeof(int64_t) + 16));
  CCU_CUDA(h, h->d_vsnp.reserve(nnz * sizeof(int32_t) + 16));
  CCU_CUDA(h, h->d_vcnt.reserve(2 * nbcs * sizeof(int32_t) + 16));
  CCU_CUDA(h, h->d_sng.reserve((size_t)nbcs * h->nv * sizeof(double) + 16));
  CCU_CUDA(h, h->d_eclass.reserve(nnz + 16));
  CCU_CUDA(h, h->d_vaff.reserve((size_t)nnz * 4 * sizeof(double) + 32));
  CCU_CUDA(h, h->d_fmat.reserve((size_t)nnz * h->nvp * sizeof(double) + 16));
  CCU_CUDA(h, h->d_pm.reserve((size_t)nbcs * h->nvp * sizeof(double) + 16));
  CCU_CUDA(h, h->d_pe.reserve((size_t)nbcs * h->nvp * sizeof(int32_t) + 16));
  CCU_CUDA(h, h->d_maxbq.reserve(32));  // deepest quality, the tile kernel's batch counter, K2's item counter
  CCU_CUDA(h, h->d_partials.reserve((size_t)nbcs * plan.items_per_cell * ccu::kScoreWarps * sizeof(ccu::ItemPartial) + 16));
  CCU_CUDA(h, h->d_results.reserve(nbcs * sizeof(ccu_demux_result) + 16));
  return 0;
}

// The upload proper.  ent_base / read_base: a slice of a larger store handed over with that store's offsets (cell_ptr
// starts at ent_base, read_ptr at read_base, the other arrays already point at the slice): the pointer arrays are
// rebased on the device, so the host never rewrites them.
static int demux_upload_impl(ccu_handle* h, int64_t nbcs, const int64_t* cell_ptr, const int32_t* snp_idx,
                             const int64_t* read_ptr, const uint8_t* al, const uint8_t* bq, int64_t ent_base,
                             int64_t read_base) {
  if (!h) return 1;
  if (!h->configured) return fail(h, "ccu_demux_upload: call ccu_demux_configure first");
  if (nbcs < 0 || !cell_ptr) return fail(h, "ccu_demux_upload: bad arguments");
  const int64_t nnz = cell_ptr[nbcs] - cell_ptr[0];
  if (cell_ptr[0] != ent_base) return fail(h, "ccu_demux_upload: cell_ptr[0] must be 0");
  if (nnz < 0 || (nnz > 0 && (!snp_idx || !read_ptr))) return fail(h, "ccu_demux_upload: bad CSR arrays");
  const int64_t nreads = nnz > 0 ? read_ptr[nnz] - read_ptr[0] : 0;
  if (nnz > 0 && read_ptr[0] != read_base) return fail(h, "ccu_demux_upload: read_ptr[0] must be 0");
  if (nreads < 0 || (nreads > 0 && (!al || !bq))) return fail(h, "ccu_demux_upload: bad read arrays");
  CCU_CUDA(h, cudaSetDevice(h->device));
  if (int rc = demux_reserve(h, nbcs, nnz, nreads)) return rc;
  CCU_CUDA(h, cudaEventRecord(h->ev[EV_X0], h->stream));
  CCU_CUDA(h, cudaMemcpyAsync(h->d_cell_ptr.p, cell_ptr, (nbcs + 1) * sizeof(int64_t), cudaMemcpyHostToDevice,
                              h->stream));
  if (nnz > 0) {
    CCU_CUDA(h, cudaMemcpyAsync(h->d_snp_idx.p, snp_idx, nnz * sizeof(int32_t), cudaMemcpyHostToDevice, h->stream));
    CCU_CUDA(h, cudaMemcpyAsync(h->d_read_ptr.p, read_ptr, (nnz + 1) * sizeof(int64_t), cudaMemcpyHostToDevice,
                                h->stream));
  }
  if (nreads > 0) {
    CCU_CUDA(h, cudaMemcpyAsync(h->d_al.p, al, nreads, cudaMemcpyHostToDevice, h->stream));
    CCU_CUDA(h, cudaMemcpyAsync(h->d_bq.p, bq, nreads, cudaMemcpyHostToDevice, h->stream));
  }
  CCU_CUDA(h, cudaEventRecord(h->ev[EV_X1], h->stream));
  CCU_CUDA(h, ccu::launch_rebase(h->d_cell_ptr.as<int64_t>(), nbcs + 1, ent_base, h->stream));
  if (nnz > 0) CCU_CUDA(h, ccu::launch_rebase(h->d_read_ptr.as<int64_t>(), nnz + 1, read_base, h->stream));
  // Full 9 x nAlpha tiles are only kept for the entries with two or more informative reads (one entry in two thousand
  // in sparse single-cell data, a third with skewed coverage): count them, on the device, and size the buffer to that
  // (80 * nAlpha bytes per entry: 3.5 GB at C2 if it were kept for every entry).
  unsigned long long ngen = 0;
  {
    ccu::DemuxDev q;
    memset(&q, 0, sizeof(q));
    q.nnz = nnz;
    q.read_ptr = h->d_read_ptr.as<int64_t>();
    q.al = h->d_al.as<uint8_t>();
    CCU_CUDA(h, ccu::launch_count_general(q, h->d_ngen.as<unsigned long long>(), h->stream));
    CCU_CUDA(h, cudaMemcpyAsync(&ngen, h->d_ngen.p, sizeof(ngen), cudaMemcpyDeviceToHost, h->stream));
  }
  CCU_CUDA(h, cudaStreamSynchronize(h->stream));
  h->ngen = (int64_t)ngen;
  CCU_CUDA(h, h->d_tiles.reserve((size_t)h->ngen * h->nA * ccu::kTileStride * sizeof(double) + 16));
  h->tm.ms_h2d = ev_ms(h, EV_X0, EV_X1);
  h->nbcs = nbcs;
  h->nnz = nnz;
  h->nreads = nreads;
  h->ran = false;
  h->validated = false;
  return 0;
}

int ccu_demux_upload(ccu_handle* h, int64_t nbcs, const int64_t* cell_ptr, const int32_t* snp_idx,
                     const int64_t* read_ptr, const uint8_t* al, const uint8_t* bq) {
  CcuRange nvtx_range("ccu_demux_upload");
  if (h) h->partial = false;
  return demux_upload_impl(h, nbcs, cell_ptr, snp_idx, read_ptr, al, bq, 0, 0);
}

int ccu_demux_run(ccu_handle* h) {
  CcuRange nvtx_range("ccu_demux_run");
  if (!h) return 1;
  if (!h->configured || h->nsnps < 0 || h->nbcs < 0)
    return fail(h, "ccu_demux_run: configure, set_genotypes and upload must precede run");
  CCU_CUDA(h, cudaSetDevice(h->device));
  const ccu::DemuxDev d = dev_view(h);
  const ccu::ScorePlan plan = ccu::make_score_plan(h->nv, h->nA, h->alpha.data());
  int grid = 0;
  cudaStream_t st = h->stream;
  int nvalidate = 0;
  if (!h->validated) {
    nvalidate = (h->nbcs > 0 || h->nnz > 0) ? 1 : 0;
    // The kernels index the genotype matrix and the read arrays with what the caller handed in: check it on the
    // device first (pointer arrays non-decreasing and closed, SNP ids inside the genotype matrix) and refuse to run
    // rather than read out of bounds.  One pass over the index arrays and one 4-byte read-back per upload.
    uint32_t bad = 0;
    CCU_CUDA(h, h->d_flag.reserve(16));
    CCU_CUDA(h, cudaMemsetAsync(h->d_flag.p, 0, 4, st));
    CCU_CUDA(h, ccu::launch_validate(d, h->nreads, h->d_flag.as<uint32_t>(), st));
    CCU_CUDA(h, cudaMemcpyAsync(&bad, h->d_flag.p, 4, cudaMemcpyDeviceToHost, st));
    CCU_CUDA(h, cudaStreamSynchronize(st));
    if (bad)
      return fail(h, "ccu_demux_run: malformed droplet store (%s%s%s)", (bad & 1) ? "cell_ptr not non-decreasing from 0 to nnz; " : "",
                  (bad & 2) ? "SNP id outside [0, nsnps); " : "", (bad & 4) ? "read_ptr not non-decreasing from 0 to the number of reads" : "");
    h->validated = true;
  }
  CCU_CUDA(h, cudaEventRecord(h->ev[EV_RUN0], st));
  CCU_CUDA(h, ccu::launch_tile(d, false, h->d_scan.p, h->d_scan.cap, nullptr, st));
  CCU_CUDA(h, cudaEventRecord(h->ev[EV_TILE], st));
  CCU_CUDA(h, ccu::launch_compact(d, st));
  CCU_CUDA(h, cudaEventRecord(h->ev[EV_COMPACT], st));
  CCU_CUDA(h, ccu::launch_sample(d, st));
  CCU_CUDA(h, cudaEventRecord(h->ev[EV_SNG], st));
  CCU_CUDA(h, ccu::launch_score(d, plan, h->num_sms, nullptr, 0, 0, &grid, st));
  CCU_CUDA(h, cudaEventRecord(h->ev[EV_SCORE], st));
  CCU_CUDA(h, ccu::launch_finalize(d, plan, st));
  CCU_CUDA(h, cudaEventRecord(h->ev[EV_FIN], st));
  h->tm.launches = (h->nbcs > 0 ? 4 : 0) + (h->nnz > 0 ? 3 : 0) + nvalidate;  // + entry, prefix sum, tile kernels
  h->tm.score_ctas = grid;
  h->tm.score_items = h->nbcs * plan.items_per_cell;
  h->ran = true;
  return 0;
}

static int collect_run_timings(ccu_handle* h) {
  CCU_CUDA(h, cudaStreamSynchronize(h->stream));
  h->tm.ms_tile = ev_ms(h, EV_RUN0, EV_TILE);
  h->tm.ms_singlet = ev_ms(h, EV_COMPACT, EV_SNG);
  h->tm.ms_score = ev_ms(h, EV_SNG, EV_SCORE);
  h->tm.ms_finalize = ev_ms(h, EV_SCORE, EV_FIN);
  h->tm.ms_run = ev_ms(h, EV_RUN0, EV_FIN);
  return 0;
}

int ccu_demux_wait(ccu_handle* h) {
  if (!h) return 1;
  if (!h->ran) return fail(h, "ccu_demux_wait: nothing has been run");
  CCU_CUDA(h, cudaSetDevice(h->device));
  return collect_run_timings(h);
}

void* ccu_demux_results_dev(ccu_handle* h, int64_t* nbcs) {
  if (!h || !h->ran || h->partial) return nullptr;
  if (nbcs) *nbcs = h->nbcs;
  return h->d_results.p;
}

int ccu_demux_download(ccu_handle* h, ccu_demux_result* out) {
  CcuRange nvtx_range("ccu_demux_download");
  if (!h) return 1;
  if (!h->ran) return fail(h, "ccu_demux_download: nothing has been run");
  if (h->nbcs > 0 && !out) return fail(h, "ccu_demux_download: out is NULL");
  CCU_CUDA(h, cudaSetDevice(h->device));
  if (collect_run_timings(h)) return 1;
  CCU_CUDA(h, cudaEventRecord(h->ev[EV_X0], h->stream));
  if (h->nbcs > 0)
    CCU_CUDA(h, cudaMemcpyAsync(out, h->d_results.p, h->nbcs * sizeof(ccu_demux_result), cudaMemcpyDeviceToHost,
                                h->stream));
  CCU_CUDA(h, cudaEventRecord(h->ev[EV_X1], h->stream));
  CCU_CUDA(h, cudaStreamSynchronize(h->stream));
  h->tm.ms_d2h = ev_ms(h, EV_X0, EV_X1);
  return 0;
}

// Device scratch one droplet range needs (the per-entry tiles and single-read likelihood rows dominate).
static double demux_scratch_bytes(const ccu_handle* h, int64_t ncell, int64_t nnz, int64_t nreads) {
  const ccu::ScorePlan plan = ccu::make_score_plan(h->nv, h->nA, h->alpha.data());
  // full tiles are kept for the entries with >= 2 informative reads only; how many that is is known after the upload,
  // so the estimate charges a quarter of the entries (real data: well under that; beyond it the allocation itself
  // fails cleanly and the caller can lower CCU_DEMUX_MAX_SCRATCH_MB)
  const double per_entry = 0.25 * (double)h->nA * ccu::kTileStride * 8 + (double)h->nvp * 8 + 8 + 4 + 1 + 32 + 4 + 8 + 4;
  const double per_cell = (double)h->nv * 8 + (double)h->nvp * 12 + 16 + sizeof(ccu_demux_result) +
                          (double)plan.items_per_cell * ccu::kScoreWarps * sizeof(ccu::ItemPartial);
  return per_entry * (double)nnz + per_cell * (double)ncell + 2.0 * (double)nreads;
}

// A store in two parts, the second part's upload under the first part's kernels.  Every kernel of a pass needs the whole
// of its store, so inside ONE pass the host->device copy cannot hide; droplets are independent, so a small first part
// (an eighth of the entries) is uploaded and started, and while it is scored the rest crosses PCIe on a second stream
// into a second set of store arrays (pointer arrays sent with the whole store's offsets and rebased on the device).
// Results are bitwise those of the single pass.  All buffers are reserved for the larger part before anything runs:
// growing one later would free it under the running kernels (cudaFree waits -- correct, but serial).
static int demux_score_two_parts(ccu_handle* h, int64_t nbcs, const int64_t* cell_ptr, const int32_t* snp_idx,
                                 const int64_t* read_ptr, const uint8_t* al, const uint8_t* bq, ccu_demux_result* out,
                                 int64_t n0) {
  const int64_t e0 = cell_ptr[n0], e1 = cell_ptr[nbcs];
  const int64_t r0 = read_ptr[e0], r1 = read_ptr[e1];
  const int64_t nb1 = nbcs - n0, nnz1 = e1 - e0, nr1 = r1 - r0;
  if (!h->aux_stream) CCU_CUDA(h, cudaStreamCreateWithFlags(&h->aux_stream, cudaStreamNonBlocking));
  if (int rc = demux_reserve(h, std::max(n0, nb1), std::max(e0, nnz1), std::max(r0, nr1))) return rc;
  CCU_CUDA(h, h->d_in2[0].reserve((nb1 + 1) * sizeof(int64_t)));
  CCU_CUDA(h, h->d_in2[1].reserve(nnz1 * sizeof(int32_t) + 16));
  CCU_CUDA(h, h->d_in2[2].reserve((nnz1 + 1) * sizeof(int64_t)));
  CCU_CUDA(h, h->d_in2[3].reserve(nr1 + 16));
  CCU_CUDA(h, h->d_in2[4].reserve(nr1 + 16));
  // ---- part 0: upload, start
  if (int rc = demux_upload_impl(h, n0, cell_ptr, snp_idx, read_ptr, al, bq, 0, 0)) return rc;
  {  // room for part 1's tiles at part 0's rate of entries with two or more reads (and a half more)
    const double rate = e0 > 0 ? (double)h->ngen / (double)e0 : 0.0;
    const size_t guess = std::min<size_t>((size_t)(1.5 * rate * (double)nnz1) + 4096, (size_t)nnz1);
    CCU_CUDA(h, h->d_tiles.reserve(guess * h->nA * ccu::kTileStride * sizeof(double) + 16));
  }
  if (int rc = ccu_demux_run(h)) return rc;
  ccu_timings sum = h->tm;
  // ---- part 1's arrays into the other set, on the auxiliary stream, while part 0 is scored
  DevBuf* cur[5] = {&h->d_cell_ptr, &h->d_snp_idx, &h->d_read_ptr, &h->d_al, &h->d_bq};
  for (int i = 0; i < 5; ++i) std::swap(*cur[i], h->d_in2[i]);
  cudaStream_t main_stream = h->stream;
  h->stream = h->aux_stream;
  const int rc1 = demux_upload_impl(h, nb1, cell_ptr + n0, snp_idx + e0, read_ptr + e0, al + r0, bq + r0, e0, r0);
  h->stream = main_stream;
  if (rc1) return rc1;
  const float h2d1 = h->tm.ms_h2d;
  // ---- part 0's results (its kernels read the first set and the scratch; nothing above touched either)
  if (collect_run_timings(h)) return 1;
  sum.ms_tile = h->tm.ms_tile;
  sum.ms_singlet = h->tm.ms_singlet;
  sum.ms_score = h->tm.ms_score;
  sum.ms_finalize = h->tm.ms_finalize;
  sum.ms_run = h->tm.ms_run;
  CCU_CUDA(h, cudaMemcpyAsync(out, h->d_results.p, n0 * sizeof(ccu_demux_result), cudaMemcpyDeviceToHost, h->stream));
  CCU_CUDA(h, cudaStreamSynchronize(h->stream));
  // ---- part 1
  if (int rc = ccu_demux_run(h)) return rc;
  if (int rc = ccu_demux_download(h, out + n0)) return rc;
  sum.ms_tile += h->tm.ms_tile;
  sum.ms_singlet += h->tm.ms_singlet;
  sum.ms_score += h->tm.ms_score;
  sum.ms_finalize += h->tm.ms_finalize;
  sum.ms_run += h->tm.ms_run;
  sum.ms_h2d += h2d1;
  sum.ms_d2h = h->tm.ms_d2h;
  sum.launches += h->tm.launches;
  sum.score_items += h->tm.score_items;
  sum.score_ctas = h->tm.score_ctas;
  h->tm = sum;
  h->partial = true;
  return 0;
}

int ccu_demux_score(ccu_handle* h, int64_t nbcs, const int64_t* cell_ptr, const int32_t* snp_idx,
                    const int64_t* read_ptr, const uint8_t* al, const uint8_t* bq, ccu_demux_result* out) {
  CcuRange nvtx_range("ccu_demux_score");
  if (!h) return 1;
  if (!h->configured) return fail(h, "ccu_demux_score: call ccu_demux_configure first");
  if (nbcs < 0 || !cell_ptr) return fail(h, "ccu_demux_score: bad arguments");
  const auto t_enter = std::chrono::steady_clock::now();
  // Droplets are independent: when the scratch of the whole store does not fit into the device memory that is
  // free right now, the store is scored in contiguous droplet ranges that do (CCU_DEMUX_MAX_SCRATCH_MB caps the
  // budget, for tests).
  CCU_CUDA(h, cudaSetDevice(h->device));
  const int64_t nnz = cell_ptr[nbcs] - cell_ptr[0];
  const int64_t nreads = (nnz > 0 && read_ptr) ? read_ptr[nnz] - read_ptr[0] : 0;
  const double need = (nbcs == 0) ? 0.0 : demux_scratch_bytes(h, nbcs, nnz, nreads);
  const char* cap = getenv("CCU_DEMUX_MAX_SCRATCH_MB");
  // A store no larger than one that already ran in one piece on this handle fits again (its buffers are still
  // held): the free-memory query -- a driver call whose latency is not ours to bound -- is skipped for it.
  double budget = h->scratch_fits;
  if (cap || need > budget) {
    size_t free_b = 0, total_b = 0;
    CCU_CUDA(h, cudaMemGetInfo(&free_b, &total_b));
    const size_t held = h->d_tiles.cap + h->d_fmat.cap + h->d_vaff.cap + h->d_vent.cap + h->d_partials.cap;  // reusable
    budget = 0.85 * ((double)free_b + (double)held);
    if (cap) budget = std::min(budget, atof(cap) * 1048576.0);
  }
  if (h) h->partial = false;
  if (need <= budget && !cap && nbcs >= 4096 && nnz >= (1 << 20) && out && snp_idx && read_ptr && cell_ptr[0] == 0 &&
      read_ptr[0] == 0 && !getenv("CCU_DEMUX_NO_PIPELINE")) {
    // first part: the droplets that hold the first eighth of the entries
    const int64_t n0 = std::lower_bound(cell_ptr, cell_ptr + nbcs, cell_ptr[nbcs] / 8) - cell_ptr;
    // (a store whose arrays are not in order takes the single pass, whose device-side validation names the defect)
    const bool sane = n0 > 0 && n0 < nbcs && cell_ptr[n0] > 0 && cell_ptr[n0] < cell_ptr[nbcs] &&
                      read_ptr[cell_ptr[n0]] >= 0 && read_ptr[cell_ptr[n0]] <= read_ptr[cell_ptr[nbcs]];
    if (sane) {
      const int rc = demux_score_two_parts(h, nbcs, cell_ptr, snp_idx, read_ptr, al, bq, out, n0);
      if (rc == 0 && need > h->scratch_fits) h->scratch_fits = need;
      return rc;
    }
  }
  if (nbcs == 0 || need <= budget) {
    const bool trace = getenv("CCU_TRACE") != nullptr;  // host wall clock of the phases, to stderr
    const auto t0 = std::chrono::steady_clock::now();
    if (int rc = ccu_demux_upload(h, nbcs, cell_ptr, snp_idx, read_ptr, al, bq)) return rc;
    const auto t1 = std::chrono::steady_clock::now();
    if (int rc = ccu_demux_run(h)) return rc;
    const auto t2 = std::chrono::steady_clock::now();
    const int rc = ccu_demux_download(h, out);
    if (rc == 0 && !cap && need > h->scratch_fits) h->scratch_fits = need;
    if (trace) {
      const auto t3 = std::chrono::steady_clock::now();
      auto ms = [](auto a, auto b) { return std::chrono::duration<double, std::milli>(b - a).count(); };
      fprintf(stderr, "[ccu] score: mem-info %.3f ms, upload %.3f ms, run (launch) %.3f ms, wait+download %.3f ms\n",
              ms(t_enter, t0), ms(t0, t1), ms(t1, t2), ms(t2, t3));
    }
    return rc;
  }
  if (cell_ptr[0] != 0 || (nnz > 0 && (!snp_idx || !read_ptr || read_ptr[0] != 0)))
    return fail(h, "ccu_demux_score: bad CSR arrays");
  ccu_timings sum = {};
  std::vector<int64_t> cp, rp;
  int64_t b = 0, nbatch = 0;
  while (b < nbcs) {
    int64_t e = b + 1;  // at least one droplet per range; grow while the range still fits
    while (e < nbcs && demux_scratch_bytes(h, e + 1 - b, cell_ptr[e + 1] - cell_ptr[b],
                                           read_ptr[cell_ptr[e + 1]] - read_ptr[cell_ptr[b]]) <= budget) {
      // grow geometrically, then settle with the linear test above
      const int64_t step = std::max<int64_t>(1, (e - b) / 2);
      const int64_t e2 = std::min(nbcs, e + step);
      if (demux_scratch_bytes(h, e2 - b, cell_ptr[e2] - cell_ptr[b], read_ptr[cell_ptr[e2]] - read_ptr[cell_ptr[b]]) <= budget) e = e2;
      else ++e;
    }
    const int64_t e0 = cell_ptr[b], e1 = cell_ptr[e], r0 = read_ptr[e0];
    cp.assign(cell_ptr + b, cell_ptr + e + 1);
    for (auto& x : cp) x -= e0;
    rp.assign(read_ptr + e0, read_ptr + e1 + 1);
    for (auto& x : rp) x -= r0;
    if (int rc = ccu_demux_upload(h, e - b, cp.data(), snp_idx + e0, rp.data(), al + r0, bq + r0)) return rc;
    if (int rc = ccu_demux_run(h)) return rc;
    if (int rc = ccu_demux_download(h, out + b)) return rc;
    sum.ms_tile += h->tm.ms_tile;
    sum.ms_singlet += h->tm.ms_singlet;
    sum.ms_score += h->tm.ms_score;
    sum.ms_finalize += h->tm.ms_finalize;
    sum.ms_run += h->tm.ms_run;
    sum.ms_h2d += h->tm.ms_h2d;
    sum.ms_d2h += h->tm.ms_d2h;
    sum.launches += h->tm.launches;
    sum.score_items += h->tm.score_items;
    sum.score_ctas = h->tm.score_ctas;
    b = e;
    ++nbatch;
  }
  sum.ms_genotypes = h->tm.ms_genotypes;
  h->tm = sum;
  h->partial = nbatch > 1;
  return 0;
}

int ccu_demux_get_tiles(ccu_handle* h, double* tiles) {
  if (!h) return 1;
  if (!h->ran) return fail(h, "ccu_demux_get_tiles: nothing has been run");
  if (h->partial) return fail(h, "ccu_demux_get_tiles: the last ccu_demux_score ran its store in parts and the handle holds only the last one: use ccu_demux_upload + ccu_demux_run (or set CCU_DEMUX_NO_PIPELINE=1) to inspect a store");
  CCU_CUDA(h, cudaSetDevice(h->device));
  const size_t rowp = (size_t)h->nA * ccu::kTileStride;
  std::vector<double> tmp((size_t)h->nnz * rowp);
  // inspection only: the tile of EVERY entry (the engine itself keeps those of the general entries), in a scratch buffer
  CCU_CUDA(h, h->d_dbg.reserve(tmp.size() * sizeof(double) + 16));
  ccu::DemuxDev d = dev_view(h);
  d.tiles = h->d_dbg.as<double>();
  CCU_CUDA(h, ccu::launch_tile(d, true, h->d_scan.p, h->d_scan.cap, nullptr, h->stream));
  CCU_CUDA(h, cudaStreamSynchronize(h->stream));
  if (h->nnz > 0) CCU_CUDA(h, cudaMemcpy(tmp.data(), h->d_dbg.p, tmp.size() * sizeof(double), cudaMemcpyDeviceToHost));
  for (int64_t e = 0; e < h->nnz; ++e)
    for (int n = 0; n < h->nA; ++n)
      memcpy(tiles + ((size_t)e * h->nA + n) * 9, tmp.data() + e * rowp + n * ccu::kTileStride, 9 * sizeof(double));
  return 0;
}

// The per-(droplet, sample) singlet log-likelihoods llksAB[j][0][0] of the last run (what <out>.single lists) and,
// optionally, llks00[0] per droplet.
int ccu_demux_get_singlets(ccu_handle* h, double* llk, double* llk0) {
  if (!h) return 1;
  if (!h->ran) return fail(h, "ccu_demux_get_singlets: nothing has been run");
  if (h->partial) return fail(h, "ccu_demux_get_singlets: the last ccu_demux_score ran its store in parts and the handle holds only the last one: use ccu_demux_upload + ccu_demux_run (or set CCU_DEMUX_NO_PIPELINE=1) to inspect a store");
  if (h->nbcs > 0 && !llk) return fail(h, "ccu_demux_get_singlets: NULL output");
  CCU_CUDA(h, cudaSetDevice(h->device));
  if (h->nbcs == 0) return 0;
  CCU_CUDA(h, cudaMemcpyAsync(llk, h->d_sng.p, (size_t)h->nbcs * h->nv * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
  if (llk0) {
    CCU_CUDA(h, h->d_dbg.reserve(((size_t)h->nsnps * 4 + (size_t)h->nbcs) * sizeof(double) + 16));
    double* gp0 = h->d_dbg.as<double>();
    double* out = gp0 + (size_t)h->nsnps * 4;
    CCU_CUDA(h, ccu::launch_llk00(dev_view(h), gp0, out, h->stream));
    CCU_CUDA(h, cudaMemcpyAsync(llk0, out, (size_t)h->nbcs * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
  }
  CCU_CUDA(h, cudaStreamSynchronize(h->stream));
  return 0;
}

int ccu_demux_get_genotypes(ccu_handle* h, int64_t snp_begin, int64_t snp_end, double* gps) {
  if (!h) return 1;
  if (h->nsnps < 0) return fail(h, "ccu_demux_get_genotypes: no genotypes set");
  if (snp_begin < 0 || snp_end > h->nsnps || snp_begin > snp_end) return fail(h, "ccu_demux_get_genotypes: bad range");
  CCU_CUDA(h, cudaSetDevice(h->device));
  const int64_t n = snp_end - snp_begin;
  if (n == 0) return 0;
  CCU_CUDA(h, cudaStreamSynchronize(h->stream));
  CCU_CUDA(h, cudaMemcpy2D(gps, (size_t)h->nv * 24, h->d_G.as<double>() + snp_begin * h->nvp * 3, (size_t)h->nvp * 24,
                           (size_t)h->nv * 24, n, cudaMemcpyDeviceToHost));
  return 0;
}

int ccu_demux_get_llk(ccu_handle* h, int64_t cell_begin, int64_t cell_end, double* llk) {
  if (!h) return 1;
  if (!h->ran) return fail(h, "ccu_demux_get_llk: run the engine first");
  if (h->partial) return fail(h, "ccu_demux_get_llk: the last ccu_demux_score ran its store in parts and the handle holds only the last one: use ccu_demux_upload + ccu_demux_run (or set CCU_DEMUX_NO_PIPELINE=1) to inspect a store");
  if (cell_begin < 0 || cell_end > h->nbcs || cell_begin > cell_end) return fail(h, "ccu_demux_get_llk: bad range");
  const int64_t nc = cell_end - cell_begin;
  if (nc == 0) return 0;
  CCU_CUDA(h, cudaSetDevice(h->device));
  const size_t n = (size_t)nc * h->nv * h->nv * h->nA;
  // d_dbg: n double2 (mantissa, exponent) followed by n doubles of converted log-likelihoods
  CCU_CUDA(h, h->d_dbg.reserve(n * 3 * sizeof(double)));
  double2* raw2 = h->d_dbg.as<double2>();
  double* llk_dev = h->d_dbg.as<double>() + 2 * n;
  CCU_CUDA(h, cudaMemsetAsync(raw2, 0xff, n * sizeof(double2), h->stream));  // all-ones = NaN marker
  const ccu::DemuxDev d = dev_view(h);
  const ccu::ScorePlan plan = ccu::make_score_plan(h->nv, h->nA, h->alpha.data());
  CCU_CUDA(h, ccu::launch_score(d, plan, h->num_sms, raw2, cell_begin, cell_end, nullptr, h->stream));
  CCU_CUDA(h, ccu::launch_dbg_convert(raw2, llk_dev, (int64_t)n, h->stream));
  CCU_CUDA(h, ccu::launch_singlet_to_dbg(d, llk_dev, cell_begin, cell_end, h->stream));
  CCU_CUDA(h, cudaMemcpyAsync(llk, llk_dev, n * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
  CCU_CUDA(h, cudaStreamSynchronize(h->stream));
  return 0;
}

int ccu_measure_fp64_peak(ccu_handle* h, float ms, double* tflops) {
  if (!h || !tflops) return 1;
  CCU_CUDA(h, cudaSetDevice(h->device));
  int iters = 2000;
  double best = 0;
  float total = 0;
  for (int rep = 0; rep < 64 && (rep < 3 || total < ms); ++rep) {
    CCU_CUDA(h, cudaEventRecord(h->ev[EV_X0], h->stream));
    CCU_CUDA(h, ccu::launch_fp64_peak(h->d_sink.as<double>(), iters, h->num_sms, h->stream));
    CCU_CUDA(h, cudaEventRecord(h->ev[EV_X1], h->stream));
    CCU_CUDA(h, cudaStreamSynchronize(h->stream));
    float t = ev_ms(h, EV_X0, EV_X1);
    total += t;
    const double flops = (double)h->num_sms * 4 * 256 * (double)iters * 64 * 2;
    if (rep > 0 && t > 0) best = fmax(best, flops / (t * 1e-3) / 1e12);
    if (t < 2.f) iters *= 2;
  }
  *tflops = best;
  return 0;
}

}  // extern "C"
