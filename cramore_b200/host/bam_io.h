// bam_io.h -- a small BAM reader and the `--sam` ingestion of `cramore demuxlet` without htslib.
//
// The reference reads SAM/BAM/CRAM through htslib (sam_filtered_reader.cpp) and joins the coordinate-sorted reads
// with the VCF on the fly (cmd_cram_demuxlet.cpp:126-395).  htslib is not available to this build; BAM, however, is
// a plain format: BGZF blocks are gzip members (zlib reads a concatenation of them transparently) around
// little-endian records (SAM spec section 4).  This header reads BAM (not SAM text, not CRAM), applies the
// reference's read filter (--min-MQ, --excl-flag, sam_filtered_reader.cpp:298-310), finds the base a read shows at a
// SNP the way bam_get_base_and_qual_and_read_and_qual does (hts_utils.cpp:279-358 -- including its quirk that only
// CIGAR 'M' counts as aligned, '=' and 'X' are ignored), and builds the droplet store with the duplicate rule of
// sc_dropseq_lib_t::add_read (sc_drop_seq.cpp:45-91: the first read of a (droplet, SNP, UMI) wins).
// Parity unpinned: the reference's own --sam path needs htslib and cannot be run here; the tests check this
// reader against BAM files written by an independent Python encoder and against the --plp path on the same reads.
#pragma once
#include <unordered_map>

#include "host_io.h"

namespace hostio {

struct BamRecord {
  int32_t tid = -1, pos = 0, l_seq = 0;
  uint32_t mapq = 0, flag = 0;
  std::vector<uint32_t> cigar;  // op = v & 15 ("MIDNSHP=X"), len = v >> 4
  std::string qname, seq, qual;  // seq decoded ("=ACMGRSVTWYHKDBN"), qual = Phred + 33
  std::vector<uint8_t> aux;
  // value of a 'Z' tag, or NULL (bam_aux_get + bam_aux2Z)
  const char* tag_z(const char* tag) const {
    size_t p = 0;
    while (p + 3 <= aux.size()) {
      const bool hit = aux[p] == (uint8_t)tag[0] && aux[p + 1] == (uint8_t)tag[1];
      const char t = (char)aux[p + 2];
      p += 3;
      size_t len = 0;
      switch (t) {
        case 'A': case 'c': case 'C': len = 1; break;
        case 's': case 'S': len = 2; break;
        case 'i': case 'I': case 'f': len = 4; break;
        case 'Z': case 'H': {
          const size_t b = p;
          while (p < aux.size() && aux[p] != 0) ++p;
          if (p >= aux.size()) return NULL;
          ++p;
          if (hit) return (t == 'Z') ? reinterpret_cast<const char*>(&aux[b]) : NULL;
          continue;
        }
        case 'B': {
          if (p + 5 > aux.size()) return NULL;
          const char st = (char)aux[p];
          uint32_t n;
          memcpy(&n, &aux[p + 1], 4);
          const size_t es = (st == 'c' || st == 'C') ? 1 : (st == 's' || st == 'S') ? 2 : 4;
          len = 5 + es * (size_t)n;
          break;
        }
        default: return NULL;  // unknown type: cannot skip it
      }
      if (hit) return NULL;  // present, but not a string
      p += len;
    }
    return NULL;
  }
  // bam_endpos: one past the last reference base the alignment covers
  int32_t endpos() const {
    int32_t e = pos;
    for (uint32_t c : cigar) {
      const uint32_t op = c & 15;
      if (op == 0 || op == 2 || op == 3 || op == 7 || op == 8) e += (int32_t)(c >> 4);
    }
    return e;
  }
};

class BamReader {
 public:
  std::vector<std::string> refs;
  int64_t n_read = 0, n_skip = 0;
  int32_t minMQ = 20;
  uint32_t exclude_flag = 0x0f04;

  explicit BamReader(const std::string& fn) : fn_(fn) {
    gz_ = gzopen(fn.c_str(), "rb");
    if (!gz_) fatal("Cannot open file %s for reading", fn.c_str());
    gzbuffer(gz_, 1 << 20);
    char magic[4];
    if (!read_exact(magic, 4) || memcmp(magic, "BAM\1", 4) != 0)
      fatal("%s is not a BAM file (SAM text and CRAM need htslib, which this build does not include; convert with "
            "`samtools view -b`)", fn.c_str());
    const int32_t l_text = read_i32();
    std::string text((size_t)std::max(l_text, 0), '\0');
    if (l_text > 0 && !read_exact(&text[0], (size_t)l_text)) fatal("%s: truncated BAM header", fn.c_str());
    const int32_t n_ref = read_i32();
    for (int32_t i = 0; i < n_ref; ++i) {
      const int32_t l_name = read_i32();
      std::string name((size_t)std::max(l_name, 0), '\0');
      if (l_name > 0 && !read_exact(&name[0], (size_t)l_name)) fatal("%s: truncated BAM header", fn.c_str());
      while (!name.empty() && name.back() == '\0') name.pop_back();
      read_i32();  // l_ref
      refs.push_back(name);
    }
  }
  ~BamReader() {
    if (gz_) gzclose(gz_);
  }
  // SAMFilteredReader::read(): the next record passing the filter; false at EOF
  bool read(BamRecord& r) {
    while (next(r)) {
      ++n_read;
      if ((int32_t)r.mapq >= minMQ && !(exclude_flag & r.flag)) return true;  // passed_filter, :298-310
      ++n_skip;
    }
    return false;
  }

 private:
  bool read_exact(void* dst, size_t n) {
    char* p = static_cast<char*>(dst);
    while (n > 0) {
      const int got = gzread(gz_, p, (unsigned)std::min<size_t>(n, 1u << 30));
      if (got <= 0) return false;
      p += got;
      n -= (size_t)got;
    }
    return true;
  }
  int32_t read_i32() {
    int32_t v = 0;
    if (!read_exact(&v, 4)) fatal("%s: truncated BAM file", fn_.c_str());
    return v;
  }
  bool next(BamRecord& r) {
    int32_t block_size = 0;
    {
      char* p = reinterpret_cast<char*>(&block_size);
      const int got = gzread(gz_, p, 4);
      if (got == 0) return false;  // clean EOF
      if (got < 0 || (got < 4 && !read_exact(p + got, (size_t)(4 - got)))) fatal("%s: truncated BAM record", fn_.c_str());
    }
    if (block_size < 32) fatal("%s: malformed BAM record", fn_.c_str());
    buf_.resize((size_t)block_size);
    if (!read_exact(buf_.data(), buf_.size())) fatal("%s: truncated BAM record", fn_.c_str());
    const uint8_t* b = buf_.data();
    auto i32 = [&](size_t off) { int32_t v; memcpy(&v, b + off, 4); return v; };
    auto u16 = [&](size_t off) { uint16_t v; memcpy(&v, b + off, 2); return v; };
    r.tid = i32(0);
    r.pos = i32(4);
    const uint32_t l_read_name = b[8];
    r.mapq = b[9];
    const uint32_t n_cigar = u16(12);
    r.flag = u16(14);
    r.l_seq = i32(16);
    size_t off = 32;
    if (off + l_read_name + 4ull * n_cigar + (size_t)(r.l_seq + 1) / 2 + (size_t)r.l_seq > buf_.size() || r.l_seq < 0)
      fatal("%s: malformed BAM record", fn_.c_str());
    r.qname.assign(reinterpret_cast<const char*>(b + off), l_read_name ? l_read_name - 1 : 0);
    off += l_read_name;
    r.cigar.resize(n_cigar);
    if (n_cigar) memcpy(r.cigar.data(), b + off, 4ull * n_cigar);
    off += 4ull * n_cigar;
    static const char* kNt = "=ACMGRSVTWYHKDBN";
    r.seq.resize((size_t)r.l_seq);
    for (int32_t i = 0; i < r.l_seq; ++i) r.seq[i] = kNt[(b[off + (i >> 1)] >> ((~i & 1) << 2)) & 15];
    off += (size_t)(r.l_seq + 1) / 2;
    r.qual.resize((size_t)r.l_seq);
    for (int32_t i = 0; i < r.l_seq; ++i) r.qual[i] = (char)(b[off + i] + 33);  // bam_get_qual_string
    off += (size_t)r.l_seq;
    r.aux.assign(b + off, b + buf_.size());
    return true;
  }
  gzFile gz_ = NULL;
  std::string fn_;
  std::vector<uint8_t> buf_;
};

constexpr int32_t kBamReadIndexNA = -1;  // BAM_READ_INDEX_NA (hts_utils.h)

// hts_utils.cpp:279-358: base and quality (Phred + 33) the read shows at 0-based reference position `pos`;
// rpos = kBamReadIndexNA when the read does not show one
inline void bam_base_at(const BamRecord& r, uint32_t pos, char& base, char& qual, int32_t& rpos) {
  const int32_t rlen = r.l_seq;
  uint32_t cpos = (uint32_t)r.pos;
  rpos = 0;
  base = 'N';
  qual = 0;
  if (!r.cigar.empty()) {
    for (uint32_t c : r.cigar) {
      const char op = "MIDNSHP=XB??????"[c & 15];
      const uint32_t len = c >> 4;
      if (op == 'M') {
        if (pos >= cpos && pos <= cpos + len - 1) {
          rpos += (int32_t)(pos - cpos);
          break;
        }
        cpos += len;
        rpos += (int32_t)len;
      } else if (op == 'D' || op == 'N') {
        if (pos >= cpos && pos <= cpos + len - 1) {
          rpos = -1;
          break;
        }
        cpos += len;
      } else if (op == 'S' || op == 'I') {
        rpos += (int32_t)len;
      }
    }
    if (rpos >= 0 && rpos <= rlen) {
      base = (rpos < rlen) ? r.seq[(size_t)rpos] : '\0';  // the reference reads the string terminator at rpos == rlen
      qual = (rpos < rlen) ? r.qual[(size_t)rpos] : '\0';
    } else {
      rpos = kBamReadIndexNA;
    }
  }
  if (rpos >= rlen) {
    rpos = kBamReadIndexNA;
    base = '.';
  }
}

struct SamOptions {
  std::string tagGroup = "CB", tagUMI = "UB";
  int32_t capBQ = 20, minBQ = 13, minMQ = 20, minTD = 0;
  uint32_t exclFlag = 0x0f04;
  std::set<std::string> valid_bcs;  // --group-list
};

struct SamVariant {
  int32_t rid, pos1;  // VCF contig id, 1-based position
  char ref, alt;
};

// cmd_cram_demuxlet.cpp:222-395 on preloaded variants (`vars`, in VCF order, positions ascending inside a contig;
// `chroms` = the VCF's contig names): reads stream by, every SNP a read overlaps is looked at, and the droplet store
// comes out in the reference's iteration order (droplet id = order of first appearance, SNP id, UMI string).
inline void load_sam(const std::string& fn, const SamOptions& opt, const std::vector<std::string>& chroms,
                     const std::vector<SamVariant>& vars, DropletStore& st) {
  if (opt.tagUMI.empty()) fatal("--tag-UMI '' (random UMIs) is not supported by this build");
  BamReader br(fn);
  br.minMQ = opt.minMQ;
  br.exclude_flag = opt.exclFlag;
  std::unordered_map<std::string, int32_t> chrom_id;
  for (size_t i = 0; i < chroms.size(); ++i) chrom_id.emplace(chroms[i], (int32_t)i);
  std::vector<int32_t> tid2rid(br.refs.size(), -1);
  bool any = false;
  for (size_t t = 0; t < br.refs.size(); ++t) {
    auto it = chrom_id.find(br.refs[t]);
    if (it != chrom_id.end()) {
      tid2rid[t] = it->second;
      any = true;
    }
  }
  if (!any) fatal("Your VCF/BCF files and SAM/BAM/CRAM files does not have any matching chromosomes, or some chromosome names are duplicated");
  // per contig: [first, last) range of `vars`
  std::vector<std::pair<int64_t, int64_t>> vrange(chroms.size(), {0, 0});
  for (int64_t i = 0; i < (int64_t)vars.size(); ++i) {
    auto& pr = vrange[(size_t)vars[i].rid];
    if (pr.second == 0 && pr.first == 0) pr.first = i;
    else if (pr.second != i) fatal("the VCF is not sorted: contig %s appears in more than one block", chroms[(size_t)vars[i].rid].c_str());
    if (i > pr.first && vars[i].pos1 < vars[i - 1].pos1) fatal("the VCF is not sorted by position on contig %s", chroms[(size_t)vars[i].rid].c_str());
    pr.second = i + 1;
  }
  st.rid2chr = chroms;
  st.snps.clear();
  for (const SamVariant& v : vars) st.snps.push_back({v.rid, v.pos1, v.ref, v.alt, 0.0});

  struct Rec {
    int32_t cell, snp, umi;
    uint8_t al, bq;
    int64_t seq;
  };
  std::vector<Rec> recs;
  std::unordered_map<std::string, int32_t> cell_id, umi_id;
  std::vector<std::string> umis;
  const char* gtag = opt.tagGroup.c_str();
  const char* utag = opt.tagUMI.c_str();
  BamRecord r;
  int64_t n_no_gtag = 0, n_no_utag = 0, nReadsSkipBCD = 0, nReadsPass = 0;
  while (br.read(r)) {
    if (r.tid < 0 || r.tid >= (int32_t)tid2rid.size()) continue;
    const int32_t rid = tid2rid[(size_t)r.tid];
    if (rid < 0) continue;  // no matching VCF contig (:225-228)
    int32_t ibcd;
    {
      const char* sbcd = ".";
      if (!opt.tagGroup.empty()) {
        const char* z = (gtag[0] && gtag[1]) ? r.tag_z(gtag) : NULL;
        if (z) sbcd = z;
        else ++n_no_gtag;
        if (!opt.valid_bcs.empty() && !opt.valid_bcs.count(sbcd)) {
          ++nReadsSkipBCD;
          continue;
        }
      }
      auto ins = cell_id.emplace(sbcd, (int32_t)st.bcs.size());
      if (ins.second) {
        st.bcs.push_back(sbcd);
        st.totl_reads.push_back(0);
      }
      ibcd = ins.first->second;
    }
    int32_t iumi;
    {
      const char* z = (utag[0] && utag[1]) ? r.tag_z(utag) : NULL;
      if (!z) ++n_no_utag;
      auto ins = umi_id.emplace(z ? z : ".", (int32_t)umis.size());
      if (ins.second) umis.push_back(ins.first->first);
      iumi = ins.first->second;
    }
    ++st.totl_reads[(size_t)ibcd];  // :325
    // SNPs from the read's start to its end (:340-367)
    const auto& pr = vrange[(size_t)rid];
    const int32_t endpos = r.endpos();
    auto lo = std::lower_bound(vars.begin() + pr.first, vars.begin() + pr.second, r.pos + 1,
                               [](const SamVariant& v, int32_t p1) { return v.pos1 < p1; });
    bool passed = false;
    for (auto it = lo; it != vars.begin() + pr.second && it->pos1 - 1 < endpos; ++it) {
      char base, qual;
      int32_t rpos;
      bam_base_at(r, (uint32_t)(it->pos1 - 1), base, qual, rpos);
      if (rpos == kBamReadIndexNA) continue;
      if (base == 'N') continue;
      if (qual - 33 < opt.minBQ) continue;
      if (rpos < opt.minTD - 1) continue;
      if (rpos + opt.minTD > r.l_seq) continue;
      const uint8_t allele = (base == it->ref) ? 0 : ((base == it->alt) ? 1 : 2);
      const int bq = qual - 33 > opt.capBQ ? opt.capBQ : qual - 33;
      recs.push_back({ibcd, (int32_t)(it - vars.begin()), iumi, allele, (uint8_t)bq, (int64_t)recs.size()});
      passed = true;
    }
    if (passed) ++nReadsPass;
  }
  if (n_no_gtag) notice("WARNING: %lld reads carry no droplet/cell tag %s and were treated as a single group", (long long)n_no_gtag, gtag);
  if (n_no_utag) notice("WARNING: %lld reads carry no UMI tag %s and were treated as a single UMI", (long long)n_no_utag, utag);
  notice("Finished reading %lld reads from %s (%lld skipped by the read filter, %lld outside --group-list, %lld overlapping SNPs)",
         (long long)br.n_read, fn.c_str(), (long long)br.n_skip, (long long)nReadsSkipBCD, (long long)nReadsPass);

  // reference iteration order: droplet id, SNP id, UMI string; of equal (droplet, SNP, UMI) the first read wins
  std::sort(recs.begin(), recs.end(), [&](const Rec& a, const Rec& b) {
    if (a.cell != b.cell) return a.cell < b.cell;
    if (a.snp != b.snp) return a.snp < b.snp;
    if (a.umi != b.umi) return umis[(size_t)a.umi] < umis[(size_t)b.umi];
    return a.seq < b.seq;
  });
  const int32_t nb = st.nbcs();
  st.cell_ptr.assign((size_t)nb + 1, 0);
  st.uniq_reads.assign((size_t)nb, 0);
  st.read_ptr.assign(1, 0);
  st.snp_idx.clear();
  st.al.clear();
  st.bq.clear();
  bool open = false;
  for (size_t i = 0; i < recs.size(); ++i) {
    const Rec& x = recs[i];
    if (i > 0 && x.cell == recs[i - 1].cell && x.snp == recs[i - 1].snp && x.umi == recs[i - 1].umi) continue;  // duplicate
    if (i == 0 || x.cell != recs[i - 1].cell || x.snp != recs[i - 1].snp) {
      if (open) st.read_ptr.push_back((int64_t)st.al.size());
      st.snp_idx.push_back(x.snp);
      ++st.cell_ptr[(size_t)x.cell + 1];
      open = true;
    }
    st.al.push_back(x.al);
    st.bq.push_back(x.bq);
    ++st.uniq_reads[(size_t)x.cell];
  }
  if (open) st.read_ptr.push_back((int64_t)st.al.size());
  for (int32_t c = 0; c < nb; ++c) st.cell_ptr[(size_t)c + 1] += st.cell_ptr[(size_t)c];
  st.numi = (int64_t)st.al.size();
}

}  // namespace hostio
