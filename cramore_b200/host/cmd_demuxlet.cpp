// cmd_demuxlet.cpp -- `cramore demuxlet` on the B200 engine: the command-line surface and output files of
// cmdCramDemuxlet (cmd_cram_demuxlet.cpp:6-1060) for the --plp path.  Ingestion is host C++ (host_io.h);
// everything between "ingestion finished" (:425) and "row printed" (:993) runs in libcramore_cuda.so.
// --sam reads BAM through the built-in reader of bam_io.h (SAM text and CRAM need htslib, which this build does not
// link: they are rejected with an error).
#include <chrono>
#include <cmath>
#include <thread>

#include "cli_params.h"
#include "cramore_cuda.h"
#include "bam_io.h"
#include "host_io.h"

using namespace hostio;

namespace {

void dump_vec(const std::string& fn, const void* p, size_t bytes) {
  FILE* f = fopen(fn.c_str(), "wb");
  if (!f) fatal("Cannot write %s", fn.c_str());
  if (bytes) fwrite(p, 1, bytes, f);
  fclose(f);
}

template <typename T>
void load_vec(const std::string& fn, std::vector<T>& v) {
  FILE* f = fopen(fn.c_str(), "rb");
  if (!f) fatal("Cannot read %s (written by --gpu-dump-inputs)", fn.c_str());
  fseek(f, 0, SEEK_END);
  const long bytes = ftell(f);
  fseek(f, 0, SEEK_SET);
  if (bytes < 0 || bytes % (long)sizeof(T) != 0) fatal("%s: unexpected size", fn.c_str());
  v.resize((size_t)bytes / sizeof(T));
  if (bytes > 0 && fread(v.data(), 1, (size_t)bytes, f) != (size_t)bytes) fatal("%s: short read", fn.c_str());
  fclose(f);
}

void load_lines(const std::string& fn, std::vector<std::string>& v) {
  LineReader lr(fn);
  while (lr.next_line()) v.push_back(lr.line);
}

std::vector<int> parse_devices(const std::string& s) {
  std::vector<int> v;
  size_t p = 0;
  while (p < s.size()) {
    size_t e = s.find(',', p);
    if (e == std::string::npos) e = s.size();
    v.push_back(atoi(s.substr(p, e - p).c_str()));
    p = e + 1;
  }
  if (v.empty()) v.push_back(0);
  return v;
}

}  // namespace

int cmdDemuxlet(int argc, char** argv) {
  // defaults of cmd_cram_demuxlet.cpp:7-36
  std::string samFile, tagGroup("CB"), tagUMI("UB"), plpPrefix, vcfFile, field("GP"), r2Info("R2"), smList, outPrefix,
      groupList, gpuDevices("0"), dumpPrefix, loadPrefix, timingsFile;
  bool gpuNoRefine = false, writeSingle = false;
  double genoErrorOffset = 0.10, genoErrorCoeffR2 = 0.00, minCallRate = 0.5, doublet_prior = 0.5;
  int32_t minMAC = 1, samVerbose = 1000000, vcfVerbose = 10000, capBQ = 20, minBQ = 13, minMQ = 20, minTD = 0,
          exclFlag = 0x0f04, minTotalReads = 0, minUMIs = 0, minCoveredSNPs = 0;
  std::vector<std::string> smIDs;
  std::vector<double> gridAlpha;

  cli::Params pl;
  pl.group("Options for input SAM/BAM/CRAM");
  pl.add("sam", &samFile, "Input SAM/BAM/CRAM file. Must be sorted by coordinates and indexed");
  pl.add("tag-group", &tagGroup, "Tag representing readgroup or cell barcodes, in the case to partition the BAM file into multiple groups. For 10x genomics, use CB");
  pl.add("tag-UMI", &tagUMI, "Tag representing UMIs. For 10x genomiucs, use UB");
  pl.group("Options for input Pileup format");
  pl.add("plp", &plpPrefix, "Input pileup format");
  pl.group("Options for input VCF/BCF");
  pl.add("vcf", &vcfFile, "Input VCF/BCF file, containing the individual genotypes (GT), posterior probability (GP), or genotype likelihood (PL)");
  pl.add("field", &field, "FORMAT field to extract the genotype, likelihood, or posterior from");
  pl.add("geno-error-offset", &genoErrorOffset, "Offset of genotype error rate. [error] = [offset] + [1-offset]*[coeff]*[1-r2]");
  pl.add("geno-error-coeff", &genoErrorCoeffR2, "Slope of genotype error rate. [error] = [offset] + [1-offset]*[coeff]*[1-r2]");
  pl.add("r2-info", &r2Info, "INFO field name representing R2 value. Used for representing imputation quality");
  pl.add("min-mac", &minMAC, "Minimum minor allele frequency");
  pl.add("min-callrate", &minCallRate, "Minimum call rate");
  pl.add("sm", &smIDs, "List of sample IDs to compare to (default: use all)");
  pl.add("sm-list", &smList, "File containing the list of sample IDs to compare");
  pl.group("Output Options");
  pl.add("out", &outPrefix, "Output file prefix");
  pl.add("alpha", &gridAlpha, "Grid of alpha to search for (default is 0, 0.5; the first value must be 0: singlets are read from it)");
  pl.add("doublet-prior", &doublet_prior, "Prior of doublet");
  pl.add("sam-verbose", &samVerbose, "Verbose message frequency for SAM/BAM/CRAM");
  pl.add("vcf-verbose", &vcfVerbose, "Verbose message frequency for VCF/BCF");
  pl.group("Read filtering Options");
  pl.add("cap-BQ", &capBQ, "Maximum base quality (higher BQ will be capped)");
  pl.add("min-BQ", &minBQ, "Minimum base quality to consider (lower BQ will be skipped)");
  pl.add("min-MQ", &minMQ, "Minimum mapping quality to consider (lower MQ will be ignored)");
  pl.add("min-TD", &minTD, "Minimum distance to the tail (lower will be ignored)");
  pl.add("excl-flag", &exclFlag, "SAM/BAM FLAGs to be excluded");
  pl.group("Cell/droplet filtering options");
  pl.add("group-list", &groupList, "List of tag readgroup/cell barcode to consider in this run. All other barcodes will be ignored. This is useful for parallelized run");
  pl.add("min-total", &minTotalReads, "Minimum number of total reads for a droplet/cell to be considered");
  pl.add("min-umi", &minUMIs, "Minimum number of UMIs for a droplet/cell to be considered");
  pl.add("min-snp", &minCoveredSNPs, "Minimum number of SNPs with coverage for a droplet/cell to be considered");
  pl.group("GPU engine options (additions of this build)");
  pl.add("gpu-devices", &gpuDevices, "Comma-separated CUDA device ids; droplets are sharded across them");
  pl.add("gpu-no-refine", &gpuNoRefine, "Skip the host pass that re-scores the reported best/next guesses in the reference's own operation order (rows then agree with the reference to ~1e-13 in the likelihoods instead of byte for byte)");
  pl.add("write-single", &writeSingle, "Also write [out].single: one row per droplet and sample with the singlet log-likelihood (the file the reference's documentation names; its writer is commented out there, cmd_cram_demuxlet.cpp:516-563)");
  pl.add("gpu-timings", &timingsFile, "Write the device times of the engine's kernels (first device) and the host time of the refinement pass to this file as one JSON object");
  pl.add("gpu-dump-inputs", &dumpPrefix, "Write the flattened engine inputs to PREFIX.* and exit without touching a GPU");
  pl.add("gpu-load-inputs", &loadPrefix, "Binary sidecar cache: read the engine inputs written by --gpu-dump-inputs instead of parsing --plp/--vcf (droplet and variant filters are those of the dump)");
  if (!pl.read(argc, argv)) return 1;

  if (gridAlpha.empty()) {  // :85-89
    gridAlpha.push_back(0);
    gridAlpha.push_back(0.5);
  }
  const int32_t nAlpha = (int32_t)gridAlpha.size();
  // What the engine cannot take is refused here, before any file is read (the reference is undefined for the first
  // two: its doublet indices stay -1 and index the sample table)
  if (nAlpha < 2) fatal("--alpha : the grid needs 0 and at least one doublet fraction (e.g. --alpha 0 --alpha 0.5)");
  if (gridAlpha[0] != 0.0) fatal("--alpha : the first value must be 0 (singlet likelihoods are read from it); got %lg", gridAlpha[0]);
  if (nAlpha > CCU_MAX_ALPHA) fatal("--alpha : at most %d values are supported (got %d)", CCU_MAX_ALPHA, nAlpha);
  if (!samFile.empty() && !plpPrefix.empty()) fatal("with --plp option, neither --sam option cannot be used");
  if (outPrefix.empty()) fatal("Missing required option(s) : --out");
  if (loadPrefix.empty() && ((plpPrefix.empty() && samFile.empty()) || vcfFile.empty()))
    fatal("Missing required option(s) : --sam (or --plp), --vcf, --out");

  // what the engine and the .best writer need: filled by ingestion, or by the binary sidecar cache
  std::vector<int64_t> cell_ptr, read_ptr;
  std::vector<int32_t> snp_idx, row_int_id;
  std::vector<uint8_t> al, bq, has_gp;
  std::vector<float> gp_f32;
  std::vector<double> snp_err;
  std::vector<std::string> row_bcs, sample_ids;
  std::vector<int32_t> row_totl, row_uniq;  // RD.TOTL / RD.UNIQ of <out>.single (empty after --gpu-load-inputs)
  int32_t nv_total = 0;
  int64_t nsnps_total = 0;
  if (!loadPrefix.empty()) {
    load_vec(loadPrefix + ".cell_ptr.i64", cell_ptr);
    load_vec(loadPrefix + ".snp_idx.i32", snp_idx);
    load_vec(loadPrefix + ".read_ptr.i64", read_ptr);
    load_vec(loadPrefix + ".al.u8", al);
    load_vec(loadPrefix + ".bq.u8", bq);
    load_vec(loadPrefix + ".gp.f32", gp_f32);
    load_vec(loadPrefix + ".snp_err.f64", snp_err);
    load_vec(loadPrefix + ".has_gp.u8", has_gp);
    load_vec(loadPrefix + ".row_int_id.i32", row_int_id);
    load_lines(loadPrefix + ".barcodes.txt", row_bcs);
    load_lines(loadPrefix + ".samples.txt", sample_ids);
    nv_total = (int32_t)sample_ids.size();
    nsnps_total = (int64_t)snp_err.size();
    if (cell_ptr.empty() || cell_ptr.size() != row_bcs.size() + 1 || row_int_id.size() != row_bcs.size() ||
        read_ptr.size() != snp_idx.size() + 1 || (int64_t)snp_idx.size() != cell_ptr.back() || al.size() != bq.size() ||
        (int64_t)al.size() != read_ptr.back() || has_gp.size() != snp_err.size() || nv_total < 1 ||
        gp_f32.size() != (size_t)nsnps_total * nv_total * 3)
      fatal("--gpu-load-inputs %s: the files do not belong together", loadPrefix.c_str());
    notice("Loaded engine inputs for %zu droplets, %lld variants, %d samples from %s.*", row_bcs.size(),
           (long long)nsnps_total, nv_total, loadPrefix.c_str());
  } else {

    std::set<std::string> sm_ids(smIDs.begin(), smIDs.end());
    if (!smList.empty()) {
      LineReader lr(smList);
      while (lr.read_fields() > 0) sm_ids.insert(lr.fields[0]);
    }
    VcfReader vr(vcfFile, sm_ids);
    vr.minMAC = minMAC;
    vr.minCallRate = minCallRate;
    vr.maxAlleles = 2;
    vr.prefetch(field.c_str(), 0);  // what both ingestion paths ask for (all but the first SNP of the --sam path)
    const int32_t nv = vr.nsamples();

    PlpFilters flt;
    flt.minRead = minTotalReads;
    flt.minUMI = minUMIs;
    flt.minSNP = minCoveredSNPs;
    flt.capBQ = capBQ;
    flt.minBQ = minBQ;
    if (!groupList.empty()) {
      LineReader lr(groupList);
      while (lr.read_fields() > 0) flt.valid_bcs.insert(lr.fields[0]);
      notice("Loaded %zu valid barcodes from %s", flt.valid_bcs.size(), groupList.c_str());
    }

    if (!vr.read()) fatal("Cannot read any single variant from %s", vcfFile.c_str());
    std::vector<float> tmp_gp(nv * 3);
    DropletStore st;
    if (!samFile.empty()) {
      // ---- BAM x VCF join, cmd_cram_demuxlet.cpp:126-395 (join_sam_vcf keeps the reference's streaming order) ----
      SamOptions so;
      so.tagGroup = tagGroup;
      so.tagUMI = tagUMI;
      so.capBQ = capBQ;
      so.minBQ = minBQ;
      so.minMQ = minMQ;
      so.minTD = minTD;
      so.exclFlag = (uint32_t)exclFlag;
      so.valid_bcs = flt.valid_bcs;
      SamJoinStats js;
      join_sam_vcf(samFile, so, vr, st, [&](int32_t /*snp*/, bool first) {
        // the first SNP is added before the loop with the default genotype-error offset of parse_posteriors and
        // without the error mix (:177, :204-209); the others at :238-268
        if (!vr.parse_posteriors(field.c_str(), first ? 1e-4 : 0, tmp_gp.data()))
          fatal("Cannot parse posterior probability at %s:%d", vr.chroms[vr.cur.rid].c_str(), vr.cur.pos1);
        gp_f32.insert(gp_f32.end(), tmp_gp.begin(), tmp_gp.end());
        double err = 0;
        if (!first) {
          err = genoErrorOffset;  // :256-262
          if (genoErrorCoeffR2 > 0) {
            float r2;
            if (!vr.info_float(r2Info.c_str(), &r2))
              fatal("Cannot extract %s (1 float value) from INFO field at %s:%d. Cannot use --geno-error-coeff", r2Info.c_str(),
                    vr.chroms[vr.cur.rid].c_str(), vr.cur.pos1);
            err += (1 - genoErrorOffset) * (1 - r2) * genoErrorCoeffR2;
          }
        }
        snp_err.push_back(err);
        has_gp.push_back(1);
      }, js);
      notice("Finished reading %lld reads from %s (%lld skipped by the read filter, %lld outside --group-list, %lld overlapping SNPs, %lld redundant)",
             (long long)js.n_read, samFile.c_str(), (long long)js.n_skip, (long long)js.nReadsSkipBCD, (long long)js.nReadsPass,
             (long long)js.nReadsRedundant);
    } else
    // ---- .var.gz x VCF join, sc_drop_seq.cpp:108-116 and :226-330 ----
    load_plp(plpPrefix, flt, st, [&](int32_t /*idx*/, const char* chr, const SnpInfo& s) {
      bool found = false, passed = false;
      while (!(found || passed)) {
        if (vr.eof) passed = true;
        else if (vr.cur.rid > s.rid) passed = true;
        else if (vr.cur.rid == s.rid) {
          if (vr.cur.pos1 > s.pos) passed = true;
          else if (vr.cur.pos1 == s.pos) {
            if (vr.cur.ref[0] != s.ref || vr.cur.alt[0] != s.alt) passed = true;
            else found = true;
          }
        }
        if (passed || found) break;
        vr.read();
      }
      gp_f32.resize(gp_f32.size() + (size_t)nv * 3, 0.0f);
      if (found) {
        if (!vr.parse_posteriors(field.c_str(), 0, tmp_gp.data()))
          fatal("Cannot parse posterior probability at %s:%d", chr, s.pos);
        memcpy(&gp_f32[gp_f32.size() - (size_t)nv * 3], tmp_gp.data(), sizeof(float) * nv * 3);
        double err = genoErrorOffset;  // :298-306
        if (genoErrorCoeffR2 > 0) {
          float r2;
          if (!vr.info_float(r2Info.c_str(), &r2))
            fatal("Cannot extract %s (1 float value) from INFO field at %s:%d. Cannot use --geno-error-coeff", r2Info.c_str(), chr, s.pos);
          err += (1 - genoErrorOffset) * (1 - r2) * genoErrorCoeffR2;
        }
        snp_err.push_back(err);  // clamped to [0, 0.999] by the engine (:308-309)
        has_gp.push_back(1);
      } else {
        snp_err.push_back(0.0);
        has_gp.push_back(0);  // SNP kept without genotypes: sc_snp_t::gps == NULL (:278-282)
      }
    });
    notice("Finished reading %lld variants from %s (%lld skipped by the variant filter)", (long long)vr.nRead, vcfFile.c_str(),
           (long long)vr.nSkip);

    // ---- droplets that produce a row, in std::map barcode order (:636-653) ----
    const int32_t nbcs = st.nbcs();
    std::vector<int32_t> order(nbcs);
    for (int32_t i = 0; i < nbcs; ++i) order[i] = i;
    std::sort(order.begin(), order.end(), [&](int32_t a, int32_t b) { return st.bcs[a] < st.bcs[b]; });
    std::vector<int32_t> row_cell;
    for (int32_t rank = 0; rank < nbcs; ++rank) {
      const int32_t i = order[rank];
      const int32_t ns = (int32_t)(st.cell_ptr[i + 1] - st.cell_ptr[i]);
      if (st.totl_reads[i] < minTotalReads || st.uniq_reads[i] < minUMIs || ns < minCoveredSNPs) continue;
      if (ns == 0) continue;
      row_cell.push_back(i);
      row_int_id.push_back(rank);
    }
    // CSR of those droplets, in row order
    cell_ptr.assign(1, 0);
    read_ptr.assign(1, 0);
    for (int32_t i : row_cell) {
      for (int64_t e = st.cell_ptr[i]; e < st.cell_ptr[i + 1]; ++e) {
        snp_idx.push_back(st.snp_idx[e]);
        al.insert(al.end(), st.al.begin() + st.read_ptr[e], st.al.begin() + st.read_ptr[e + 1]);
        bq.insert(bq.end(), st.bq.begin() + st.read_ptr[e], st.bq.begin() + st.read_ptr[e + 1]);
        read_ptr.push_back((int64_t)al.size());
      }
      cell_ptr.push_back((int64_t)snp_idx.size());
    }
    for (int32_t i : row_cell) row_bcs.push_back(st.bcs[i]);
    for (int32_t i : row_cell) {
      row_totl.push_back((int32_t)st.totl_reads[i]);
      row_uniq.push_back((int32_t)st.uniq_reads[i]);
    }
    for (int32_t j = 0; j < nv; ++j) sample_ids.push_back(vr.sample_id(j));
    nv_total = nv;
    nsnps_total = st.nsnps();

    if (!dumpPrefix.empty()) {  // test hook: the flattened inputs, bit for bit
      dump_vec(dumpPrefix + ".cell_ptr.i64", cell_ptr.data(), cell_ptr.size() * 8);
      dump_vec(dumpPrefix + ".snp_idx.i32", snp_idx.data(), snp_idx.size() * 4);
      dump_vec(dumpPrefix + ".read_ptr.i64", read_ptr.data(), read_ptr.size() * 8);
      dump_vec(dumpPrefix + ".al.u8", al.data(), al.size());
      dump_vec(dumpPrefix + ".bq.u8", bq.data(), bq.size());
      dump_vec(dumpPrefix + ".gp.f32", gp_f32.data(), gp_f32.size() * 4);
      dump_vec(dumpPrefix + ".snp_err.f64", snp_err.data(), snp_err.size() * 8);
      dump_vec(dumpPrefix + ".has_gp.u8", has_gp.data(), has_gp.size());
      dump_vec(dumpPrefix + ".row_int_id.i32", row_int_id.data(), row_int_id.size() * 4);
      Writer w(dumpPrefix + ".barcodes.txt", false);
      for (const std::string& b : row_bcs) w.printf("%s\n", b.c_str());
      Writer ws(dumpPrefix + ".samples.txt", false);
      for (const std::string& b : sample_ids) ws.printf("%s\n", b.c_str());
      notice("Wrote engine inputs for %lld droplets to %s.*", (long long)row_bcs.size(), dumpPrefix.c_str());
      return 0;
    }

  }
  const int32_t nv = nv_total;
  const int64_t nrows = (int64_t)row_bcs.size();
  if (nv < 2) fatal("--vcf / --sm : at least 2 samples are needed to tell singlets from doublets (got %d)", nv);

  // ---- engine: one handle per GPU, contiguous row ranges balanced by entry count ----
  notice("Starting to identify best matching individual IDs");
  const std::vector<int> devs = parse_devices(gpuDevices);
  const int ndev = (int)devs.size();
  std::vector<ccu_demux_result> res(nrows);
  std::vector<double> sng_llk(writeSingle ? (size_t)nrows * nv : 0), llk00(writeSingle ? (size_t)nrows : 0);
  std::vector<int64_t> cut(ndev + 1, nrows);
  cut[0] = 0;
  for (int d = 1; d < ndev; ++d) {
    const int64_t target = cell_ptr[nrows] * d / ndev;
    cut[d] = std::lower_bound(cell_ptr.begin(), cell_ptr.end(), target) - cell_ptr.begin();
    cut[d] = std::min<int64_t>(std::max(cut[d], cut[d - 1]), nrows);
  }
  // The genotype matrix crosses PCIe once: the first device uploads and mixes it, the others copy the mixed matrix
  // from it over NVLink (ccu_demux_copy_genotypes); then every device scores its own range of rows.
  std::vector<std::string> errs(ndev);
  std::vector<ccu_handle*> hs(ndev, (ccu_handle*)NULL);
  for (int d = 0; d < ndev; ++d) {
    if (ccu_create(devs[d], &hs[d])) fatal("%s", ccu_last_error(NULL));
    if (ccu_demux_configure(hs[d], nv, nAlpha, gridAlpha.data(), doublet_prior)) fatal("%s", ccu_last_error(hs[d]));
  }
  if (ccu_demux_set_genotypes(hs[0], nsnps_total, gp_f32.data(), snp_err.data(), has_gp.data())) fatal("%s", ccu_last_error(hs[0]));
  // <out>.single reads the engine's singlet table after the scoring: the store must then be scored in one piece
  if (writeSingle) setenv("CCU_DEMUX_NO_PIPELINE", "1", 1);
  auto work = [&](int d) {
    ccu_handle* h = hs[d];
    const int64_t b = cut[d], e = cut[d + 1];
    std::vector<int64_t> cp(cell_ptr.begin() + b, cell_ptr.begin() + e + 1), rp;
    const int64_t e0 = cp.front(), e1 = cp.back();
    for (auto& x : cp) x -= e0;
    rp.assign(read_ptr.begin() + e0, read_ptr.begin() + e1 + 1);
    const int64_t r0 = rp.front();
    for (auto& x : rp) x -= r0;
    if ((d > 0 && ccu_demux_copy_genotypes(h, hs[0])) ||
        ccu_demux_score(h, e - b, cp.data(), snp_idx.data() + e0, rp.data(), al.data() + r0, bq.data() + r0, res.data() + b) ||
        (writeSingle && e > b && ccu_demux_get_singlets(h, sng_llk.data() + (size_t)b * nv, llk00.data() + b)))
      errs[d] = ccu_last_error(h);
  };
  if (ndev == 1) {
    work(0);
  } else {
    std::vector<std::thread> th;
    for (int d = 0; d < ndev; ++d) th.emplace_back(work, d);
    for (auto& t : th) t.join();
  }
  ccu_timings tm0;
  const bool have_tm = ccu_get_timings(hs[0], &tm0) == 0;
  for (int d = 0; d < ndev; ++d) ccu_destroy(hs[d]);
  for (int d = 0; d < ndev; ++d)
    if (!errs[d].empty()) fatal("%s", errs[d].c_str());
  // ---- the hypotheses a row names, re-scored in the reference's own operation order (cmd_cram_demuxlet.cpp:655-747):
  // orientation of alpha = 0.5 mirror pairs and every printed digit then equal the reference's (demux_refine.cpp)
  const auto t_ref0 = std::chrono::steady_clock::now();
  if (!gpuNoRefine &&
      ccu_demux_refine(nrows, cell_ptr.data(), snp_idx.data(), read_ptr.data(), al.data(), bq.data(), nv, nAlpha,
                       gridAlpha.data(), nsnps_total, gp_f32.data(), snp_err.data(), has_gp.data(), 0, res.data()))
    fatal("ccu_demux_refine: bad arguments");
  const double ms_refine = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_ref0).count();
  if (!timingsFile.empty() && have_tm) {  // machine-readable per-kernel timing (SURVEY section 5 build notes)
    Writer wt(timingsFile, false);
    wt.printf("{\"droplets\": %lld, \"entries\": %lld, \"samples\": %d, \"alphas\": %d, \"devices\": %d, "
              "\"ms_genotypes\": %.4f, \"ms_csr_h2d\": %.4f, \"ms_tile\": %.4f, \"ms_singlet\": %.4f, \"ms_score\": %.4f, "
              "\"ms_finalize\": %.4f, \"ms_kernels\": %.4f, \"ms_result_d2h\": %.4f, \"kernel_launches\": %d, "
              "\"score_ctas\": %d, \"score_items\": %lld, \"ms_refine_host\": %.3f}\n",
              (long long)nrows, (long long)cell_ptr[nrows], nv, nAlpha, ndev, tm0.ms_genotypes, tm0.ms_h2d, tm0.ms_tile,
              tm0.ms_singlet, tm0.ms_score, tm0.ms_finalize, tm0.ms_run, tm0.ms_d2h, tm0.launches, tm0.score_ctas,
              (long long)tm0.score_items, gpuNoRefine ? 0.0 : ms_refine);
    wt.close();
  }

  // ---- <out>.best: header :629, rows :993 ----
  Writer wbest(outPrefix + ".best", false);
  const char* hdr = ccu_demux_best_header();
  wbest.write(hdr, (int)strlen(hdr));
  std::vector<const char*> ids(nv);
  for (int32_t j = 0; j < nv; ++j) ids[j] = sample_ids[j].c_str();
  std::vector<char> row(8192);
  for (int64_t r = 0; r < nrows; ++r) {
    int n;
    while ((n = ccu_demux_best_row(row.data(), (int32_t)row.size(), row_int_id[r], row_bcs[r].c_str(),
                                   (uint32_t)(cell_ptr[r + 1] - cell_ptr[r]),
                                   (uint32_t)(read_ptr[cell_ptr[r + 1]] - read_ptr[cell_ptr[r]]), &res[r], nv, nAlpha,
                                   gridAlpha.data(), doublet_prior, ids.data())) == -2)
      row.resize(row.size() * 4);  // very long sample ids / barcodes
    if (n < 0) fatal("cannot format the row of droplet %s", row_bcs[r].c_str());
    wbest.write(row.data(), n);
  }
  wbest.close();
  if (writeSingle) {
    // <out>.single: header of cmd_cram_demuxlet.cpp:516, row format of :552-563 (commented out in the reference, so
    // no build of it writes this file: parity unpinned).  LLK1 = the engine's singlet log-likelihood llksAB[j][0][0],
    // LLK0 = llks00[0] (:762-773), POSTPRB = exp(LLK1 - log sum_j exp LLK1) accumulated as at :530-535.  RD.PASS is
    // not kept by this loader and repeats RD.UNIQ.
    Writer ws(outPrefix + ".single", false);
    ws.printf("BARCODE\tSM_ID\tRD.TOTL\tRD.PASS\tRD.UNIQ\tN.SNP\tLLK1\tLLK0\tPOSTPRB\n");
    for (int64_t r = 0; r < nrows; ++r) {
      const double* llk = &sng_llk[(size_t)r * nv];
      double sumLLK = -1e300;
      for (int32_t j = 0; j < nv; ++j)
        sumLLK = (sumLLK > llk[j]) ? sumLLK + log(1.0 + exp(llk[j] - sumLLK)) : llk[j] + log(1.0 + exp(sumLLK - llk[j]));
      const int32_t uniq = (int32_t)(read_ptr[cell_ptr[r + 1]] - read_ptr[cell_ptr[r]]);
      const int32_t totl = row_totl.empty() ? uniq : row_totl[r];
      for (int32_t j = 0; j < nv; ++j)
        ws.printf("%s\t%s\t%d\t%d\t%d\t%d\t%.5lf\t%.5lf\t%.3lg\n", row_bcs[r].c_str(), sample_ids[j].c_str(), totl,
                  row_uniq.empty() ? uniq : row_uniq[r], uniq, (int32_t)(cell_ptr[r + 1] - cell_ptr[r]), llk[j], llk00[r],
                  exp(llk[j] - sumLLK));
    }
    ws.close();
  }
  notice("Finished writing output files");
  return 0;
}
