// cmd_dsc_pileup.cpp -- `cramore dsc-pileup`: BAM x VCF -> digital pileup (<out>.cel.gz / .var.gz / .plp.gz / .umi.gz),
// the on-disk format `cramore demuxlet --plp` and `cramore freemuxlet --plp` read (cmd_cram_dsc_pileup.cpp).
// Host only: no GPU, no engine.  Flags, defaults, file headers and row formats are the reference's
// (cmd_cram_dsc_pileup.cpp:11-71, 438-523); the BAM is read by bam_io.h (no htslib: BAM only, no SAM text / CRAM,
// text VCF only).  Parity unpinned: the reference's command needs htslib and cannot be run here; the tests compare
// the files with an independent Python restatement on synthetic alignments and run them through demuxlet --plp.
#include "bam_io.h"
#include "cli_params.h"

using namespace hostio;

int cmdDscPileup(int argc, char** argv) {
  std::string samFile, tagGroup("CB"), tagUMI("UB"), vcfFile, smList, outPrefix, groupList;
  std::vector<std::string> smIDs;
  int32_t capBQ = 40, minBQ = 13, minTD = 0, minMQ = 20, exclFlag = 0x0f04, excludeFlag = 0x0704, samVerbose = 1000000,
          vcfVerbose = 10000, minTotalReads = 0, minUniqReads = 0, minCoveredSNPs = 0;
  bool skipUmiFlag = false;

  cli::Params pl;
  pl.group("Options for input SAM/BAM/CRAM");
  pl.add("sam", &samFile, "Input SAM/BAM/CRAM file. Must be sorted by coordinates and indexed");
  pl.add("tag-group", &tagGroup, "Tag representing readgroup or cell barcodes, in the case to partition the BAM file into multiple groups. For 10x genomics, use CB");
  pl.add("tag-UMI", &tagUMI, "Tag representing UMIs. For 10x genomiucs, use UB");
  pl.add("exclude-flag", &excludeFlag, "SAM/BAM flag to exclude");
  pl.group("Options for input VCF/BCF");
  pl.add("vcf", &vcfFile, "Input VCF/BCF file, containing the AC and AN field");
  pl.add("sm", &smIDs, "List of sample IDs to compare to (default: use all)");
  pl.add("sm-list", &smList, "File containing the list of sample IDs to compare");
  pl.group("Output Options");
  pl.add("out", &outPrefix, "Output file prefix");
  pl.add("sam-verbose", &samVerbose, "Verbose message frequency for SAM/BAM/CRAM");
  pl.add("vcf-verbose", &vcfVerbose, "Verbose message frequency for VCF/BCF");
  pl.add("skip-umi", &skipUmiFlag, "Do not generate [prefix].umi.gz file, which stores the regions covered by each barcode/UMI pair");
  pl.group("SNP-overlapping Read filtering Options");
  pl.add("cap-BQ", &capBQ, "Maximum base quality (higher BQ will be capped)");
  pl.add("min-BQ", &minBQ, "Minimum base quality to consider (lower BQ will be skipped)");
  pl.add("min-MQ", &minMQ, "Minimum mapping quality to consider (lower MQ will be ignored)");
  pl.add("min-TD", &minTD, "Minimum distance to the tail (lower will be ignored)");
  pl.add("excl-flag", &exclFlag, "SAM/BAM FLAGs to be excluded");
  pl.group("Cell/droplet filtering options");
  pl.add("group-list", &groupList, "List of tag readgroup/cell barcode to consider in this run. All other barcodes will be ignored. This is useful for parallelized run");
  pl.add("min-total", &minTotalReads, "Minimum number of total reads for a droplet/cell to be considered");
  pl.add("min-uniq", &minUniqReads, "Minimum number of unique reads (determined by UMI/SNP pair) for a droplet/cell to be considered");
  pl.add("min-snp", &minCoveredSNPs, "Minimum number of SNPs with coverage for a droplet/cell to be considered");
  if (!pl.read(argc, argv)) return 1;

  SamOptions so;
  if (!groupList.empty()) {
    LineReader lr(groupList);
    while (lr.read_fields() > 0) so.valid_bcs.insert(lr.fields[0]);
    notice("Finished loading %zu droplet/cell barcodes to consider", so.valid_bcs.size());
  }
  std::set<std::string> sm_ids(smIDs.begin(), smIDs.end());
  if (!smList.empty()) {
    LineReader lr(smList);
    while (lr.read_fields() > 0) sm_ids.insert(lr.fields[0]);
  }
  if (samFile.empty() || vcfFile.empty()) fatal("Missing required option(s) : --sam, --vcf");
  notice("Initializing BCF reader..");
  VcfReader vr(vcfFile, sm_ids, /*allow_no_samples=*/true);
  vr.minMAC = 0;  // no genotype-based threshold: every record passes (:87-88 set maxAlleles only, which passed_vfilter
  vr.minCallRate = 0;  // then never reaches)
  vr.maxAlleles = 2;
  if (outPrefix.empty()) fatal("--out parameter is missing");

  so.tagGroup = tagGroup;
  so.tagUMI = tagUMI;
  so.capBQ = capBQ;
  so.minBQ = minBQ;
  so.minMQ = minMQ;
  so.minTD = minTD;
  so.exclFlag = (uint32_t)exclFlag;
  so.pileup_mode = true;
  so.excludeFlag2 = (uint32_t)excludeFlag;
  so.collect_umi_loci = !skipUmiFlag;

  if (!vr.read()) fatal("Cannot read any single variant from %s", vcfFile.c_str());
  DropletStore st;
  SamJoinStats js;
  std::vector<std::string> refs, umis;
  std::vector<UmiLoci> loci;
  std::vector<int32_t> cell_numi;
  notice("Initializing SAM reader..");
  join_sam_vcf(samFile, so, vr, st, [&](int32_t snp, bool /*first*/) { st.snps[(size_t)snp].af = vr.calculate_af(true); }, js,
               &refs, &loci, &umis, &cell_numi);

  notice("Finished reading %d markers from the VCF file", st.nsnps());
  notice("Total number input reads : %lld", (long long)js.n_read);
  notice("Total number of unmapped or QC-failed reads : %lld", (long long)js.nReadsFilt);
  notice("Total number of read-QC-passed reads : %lld ", (long long)(js.n_read - js.n_skip));
  notice("Total number of skipped reads with ignored barcodes : %lld", (long long)js.nReadsSkipBCD);
  notice("Total number of non-skipped reads with considered barcodes : %lld", (long long)js.nReadsTMP);
  notice("Total number of gapped/noninformative reads : %lld", (long long)js.nReadsN);
  notice("Total number of base-QC-failed reads : %lld", (long long)js.nReadsLQ);
  notice("Total number of redundant reads : %lld", (long long)js.nReadsRedundant);
  notice("Total number of pass-filtered reads : %lld", (long long)js.nReadsPass);
  notice("Total number of pass-filtered reads overlapping with multiple SNPs : %lld", (long long)js.nReadsMultiSNPs);

  // pruning (:394-410): guarded by `minTotalReads + minUniqReads + minCoveredSNPs < 0`, which no sensible setting
  // satisfies -- the thresholds only act when their sum is negative, in which case no droplet is below them
  notice("Starting to prune out cells with too few reads...");
  const int32_t nbcs = st.nbcs();
  std::vector<uint8_t> pruned((size_t)nbcs, 0);
  int32_t nRemoved = 0;
  if (minTotalReads + minUniqReads + minCoveredSNPs < 0)
    for (int32_t i = 0; i < nbcs; ++i)
      if (st.totl_reads[i] < minTotalReads || st.uniq_reads[i] < minUniqReads ||
          (int32_t)(st.cell_ptr[i + 1] - st.cell_ptr[i]) < minCoveredSNPs) {
        pruned[(size_t)i] = 1;
        ++nRemoved;
      }
  notice("Finishing pruning out %d cells with too few reads...", nRemoved);

  // ---- <out>.cel.gz (:447-463) ----
  notice("Writing cell information");
  {
    Writer wC(outPrefix + ".cel.gz", true);
    wC.printf("#DROPLET_ID\tBARCODE\tNUM.READ\tNUM.UMI\tNUM.UMIwSNP\tNUM.SNP\n");
    for (int32_t i = 0; i < nbcs; ++i) {
      const unsigned nsnp = pruned[(size_t)i] ? 0u : (unsigned)(st.cell_ptr[i + 1] - st.cell_ptr[i]);
      if (!skipUmiFlag)
        wC.printf("%d\t%s\t%d\t%u\t%d\t%u\n", i, st.bcs[i].c_str(), st.totl_reads[i], (unsigned)cell_numi[(size_t)i], st.uniq_reads[i], nsnp);
      else
        wC.printf("%d\t%s\t%d\t.\t%d\t%u\n", i, st.bcs[i].c_str(), st.totl_reads[i], st.uniq_reads[i], nsnp);
    }
  }
  notice("Finished writing cell information");

  // ---- <out>.umi.gz (:470-491): droplets ascending, UMIs in string order, merged regions per strand ----
  notice("Build indices for sparse matrix");
  if (!skipUmiFlag) {
    std::vector<size_t> order(loci.size());
    for (size_t i = 0; i < order.size(); ++i) order[i] = i;
    std::sort(order.begin(), order.end(), [&](size_t a, size_t b) {
      if (loci[a].cell != loci[b].cell) return loci[a].cell < loci[b].cell;
      return umis[(size_t)loci[a].umi] < umis[(size_t)loci[b].umi];
    });
    Writer wU(outPrefix + ".umi.gz", true);
    std::string row;
    char buf[512];
    for (size_t oi : order) {
      const UmiLoci& u = loci[oi];
      if (u.strand[0].empty() && u.strand[1].empty()) continue;
      unsigned long len[2] = {0, 0};
      for (int s = 0; s < 2; ++s)
        for (const UmiLocus& l : u.strand[s]) len[s] += (unsigned long)(l.end0 - l.beg1 + 1);
      snprintf(buf, sizeof(buf), "%d\t%s\t%u\t%u", u.cell, umis[(size_t)u.umi].c_str(), (unsigned)len[0], (unsigned)len[1]);
      row = buf;
      for (int s = 0; s < 2; ++s)
        for (const UmiLocus& l : u.strand[s]) {
          snprintf(buf, sizeof(buf), "\t%s:%d:%d:%c", refs[(size_t)l.tid].c_str(), l.beg1, l.end0 - l.beg1 + 1, s ? '-' : '+');
          row += buf;
        }
      row += "\n";
      wU.write(row.c_str(), (int)row.size());
    }
  }

  // ---- <out>.var.gz and <out>.plp.gz (:494-517): SNP-major, droplets ascending inside a SNP ----
  {
    // transpose the droplet-major store: entries per SNP in ascending droplet id
    const int32_t nsnps = st.nsnps();
    std::vector<int64_t> sptr((size_t)nsnps + 1, 0);
    for (int32_t c = 0; c < nbcs; ++c)
      if (!pruned[(size_t)c])
        for (int64_t e = st.cell_ptr[c]; e < st.cell_ptr[c + 1]; ++e) ++sptr[(size_t)st.snp_idx[e] + 1];
    for (int32_t s = 0; s < nsnps; ++s) sptr[(size_t)s + 1] += sptr[(size_t)s];
    std::vector<int64_t> ent((size_t)sptr[(size_t)nsnps]);
    std::vector<int32_t> ent_cell(ent.size());
    {
      std::vector<int64_t> fill(sptr.begin(), sptr.end() - 1);
      for (int32_t c = 0; c < nbcs; ++c)
        if (!pruned[(size_t)c])
          for (int64_t e = st.cell_ptr[c]; e < st.cell_ptr[c + 1]; ++e) {
            const int64_t at = fill[(size_t)st.snp_idx[e]]++;
            ent[(size_t)at] = e;
            ent_cell[(size_t)at] = c;
          }
    }
    // rchroms (:148-165): the BAM's name of every VCF contig the BAM has too; a contig only the VCF knows prints empty
    std::vector<std::string> rchroms(vr.chroms.size());
    for (const std::string& name : refs)
      for (size_t i = 0; i < vr.chroms.size(); ++i)
        if (vr.chroms[i] == name) rchroms[i] = name;
    Writer wV(outPrefix + ".var.gz", true), wP(outPrefix + ".plp.gz", true);
    wV.printf("#SNP_ID\tCHROM\tPOS\tREF\tALT\tAF\n");
    wP.printf("#DROPLET_ID\tSNP_ID\tALLELES\tBASEQS\n");
    std::string row;
    char buf[64];
    for (int32_t s = 0; s < nsnps; ++s) {
      const SnpInfo& v = st.snps[(size_t)s];
      wV.printf("%d\t%s\t%d\t%c\t%c\t%.5lf\n", s, rchroms[(size_t)v.rid].c_str(), v.pos, v.ref, v.alt, v.af);
      for (int64_t k = sptr[(size_t)s]; k < sptr[(size_t)s + 1]; ++k) {
        const int64_t e = ent[(size_t)k];
        snprintf(buf, sizeof(buf), "%d\t%d\t", ent_cell[(size_t)k], s);
        row = buf;
        for (int64_t r = st.read_ptr[e]; r < st.read_ptr[e + 1]; ++r) row += (char)('0' + st.al[(size_t)r]);
        row += '\t';
        for (int64_t r = st.read_ptr[e]; r < st.read_ptr[e + 1]; ++r) row += (char)(33 + st.bq[(size_t)r]);
        row += '\n';
        wP.write(row.c_str(), (int)row.size());
      }
    }
  }
  notice("Finished builing indices for sparse matrix");
  return 0;
}
