// bam_io.h -- a small BAM reader and the `--sam` ingestion of `cramore demuxlet` without htslib.
//
// The reference reads SAM/BAM/CRAM through htslib (sam_filtered_reader.cpp) and joins the coordinate-sorted reads
// with the VCF on the fly (cmd_cram_demuxlet.cpp:126-395).  htslib is not available to this build; BAM, however, is
// a plain format: BGZF blocks are gzip members (zlib reads a concatenation of them transparently) around
// little-endian records (SAM spec section 4).  This header reads BAM and SAM text (not CRAM), applies the
// reference's read filter (--min-MQ, --excl-flag, sam_filtered_reader.cpp:298-310), finds the base a read shows at a
// SNP the way bam_get_base_and_qual_and_read_and_qual does (hts_utils.cpp:279-358 -- including its quirk that only
// CIGAR 'M' counts as aligned, '=' and 'X' are ignored), and builds the droplet store with the duplicate rule of
// sc_dropseq_lib_t::add_read (sc_drop_seq.cpp:45-91: the first read of a (droplet, SNP, UMI) wins).
// Parity unpinned: the reference's own --sam path needs htslib and cannot be run here; the tests check this
// reader against BAM files written by an independent Python encoder and against the --plp path on the same reads.
#pragma once
#include <memory>
#include <unordered_map>
#include <unordered_set>

#include "host_io.h"

namespace hostio {

struct BamRecord {
  int32_t tid = -1, pos = 0, l_seq = 0;
  uint32_t mapq = 0, flag = 0, n_cigar = 0;
  // The raw record (SAM spec 4.2) and offsets into it: nothing is decoded until somebody asks -- most reads of a
  // single-cell BAM overlap no SNP at all, and a read that does shows one base.
  std::vector<uint8_t> data;
  size_t off_cigar = 0, off_seq = 0, off_qual = 0, off_aux = 0;
  uint32_t cigar_at(uint32_t i) const {  // op = v & 15 ("MIDNSHP=X"), len = v >> 4
    uint32_t v;
    memcpy(&v, data.data() + off_cigar + 4ull * i, 4);
    return v;
  }
  char base_at(int32_t i) const { return "=ACMGRSVTWYHKDBN"[(data[off_seq + ((size_t)i >> 1)] >> ((~i & 1) << 2)) & 15]; }
  char qual_at(int32_t i) const { return (char)(data[off_qual + (size_t)i] + 33); }  // bam_get_qual_string
  std::string qname() const { return std::string(reinterpret_cast<const char*>(data.data() + 32)); }
  // value of a 'Z' tag, or NULL (bam_aux_get + bam_aux2Z)
  const char* tag_z(const char* tag) const {
    const uint8_t* aux = data.data() + off_aux;
    const size_t naux = data.size() - off_aux;
    size_t p = 0;
    while (p + 3 <= naux) {
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
          while (p < naux && aux[p] != 0) ++p;
          if (p >= naux) return NULL;
          ++p;
          if (hit) return (t == 'Z') ? reinterpret_cast<const char*>(&aux[b]) : NULL;
          continue;
        }
        case 'B': {
          if (p + 5 > naux) return NULL;
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
    for (uint32_t i = 0; i < n_cigar; ++i) {
      const uint32_t c = cigar_at(i);
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

  explicit BamReader(const std::string& fn) : fn_(fn), in_(fn) {
    char magic[4] = {0, 0, 0, 0};
    const bool got4 = read_exact(magic, 4);
    if (got4 && memcmp(magic, "CRAM", 4) == 0)
      fatal("%s is a CRAM file: CRAM needs htslib (reference-based codecs), which this build does not include; convert "
            "with `samtools view -b`", fn.c_str());
    if (!got4 || memcmp(magic, "BAM\1", 4) != 0) {
      // SAM text (plain or gzip): header lines give the contigs, alignment lines are re-encoded as BAM records
      if (!got4 || magic[0] != '@') fatal("%s is neither BAM nor SAM text with a header", fn.c_str());
      sam_.reset(new LineReader(fn));
      while (sam_->next_line()) {
        const std::string& l = sam_->line;
        if (l.empty() || l[0] != '@') {
          sam_pending_ = true;
          break;
        }
        if (l.compare(0, 4, "@SQ\t") == 0) {
          const size_t a = l.find("\tSN:");
          if (a == std::string::npos) fatal("%s: @SQ line without SN", fn.c_str());
          const size_t e = l.find('\t', a + 4);
          refs.push_back(l.substr(a + 4, e == std::string::npos ? std::string::npos : e - a - 4));
        }
      }
      for (size_t i = 0; i < refs.size(); ++i) ref_id_.emplace(refs[i], (int32_t)i);
      return;
    }
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
  bool read_exact(void* dst, size_t n) { return in_.read(dst, n) == n; }
  int32_t read_i32() {
    int32_t v = 0;
    if (!read_exact(&v, 4)) fatal("%s: truncated BAM file", fn_.c_str());
    return v;
  }
  // one SAM text line -> the BAM layout of the same alignment (SAM spec 1.4 / 4.2); tags other than A, i, f, Z, H
  // are dropped (nothing downstream reads them)
  bool next_sam(BamRecord& r) {
    if (!sam_pending_ && !sam_->next_line()) return false;
    sam_pending_ = false;
    std::string& l = sam_->line;
    std::vector<char*> f;
    char* p = &l[0];
    f.push_back(p);
    for (; *p; ++p)
      if (*p == '\t') {
        *p = '\0';
        f.push_back(p + 1);
      }
    if (f.size() < 11) fatal("%s: SAM line %lld has %zu fields", fn_.c_str(), (long long)sam_->nlines, f.size());
    const size_t l_name = strlen(f[0]) + 1;
    if (l_name > 255) fatal("%s: read name too long", fn_.c_str());
    int32_t tid = -1;
    if (strcmp(f[2], "*") != 0) {
      auto it = ref_id_.find(f[2]);
      if (it == ref_id_.end()) fatal("%s: contig %s is not in the header", fn_.c_str(), f[2]);
      tid = it->second;
    }
    std::vector<uint32_t> cig;
    if (strcmp(f[5], "*") != 0)
      for (const char* c = f[5]; *c;) {
        char* e;
        const unsigned long n = strtoul(c, &e, 10);
        const char* op = strchr("MIDNSHP=X", *e);
        if (e == c || !*e || !op) fatal("%s: malformed CIGAR %s", fn_.c_str(), f[5]);
        cig.push_back((uint32_t)(n << 4) | (uint32_t)(op - "MIDNSHP=X"));
        c = e + 1;
      }
    const bool has_seq = strcmp(f[9], "*") != 0;
    const int32_t l_seq = has_seq ? (int32_t)strlen(f[9]) : 0;
    std::vector<uint8_t>& d = r.data;
    d.assign(32, 0);
    auto put32 = [&](size_t off, int32_t v) { memcpy(d.data() + off, &v, 4); };
    put32(0, tid);
    put32(4, atoi(f[3]) - 1);
    d[8] = (uint8_t)l_name;
    d[9] = (uint8_t)atoi(f[4]);
    const uint16_t ncig = (uint16_t)cig.size(), flag = (uint16_t)strtoul(f[1], NULL, 0);
    memcpy(d.data() + 12, &ncig, 2);
    memcpy(d.data() + 14, &flag, 2);
    put32(16, l_seq);
    put32(20, -1);
    put32(24, -1);
    d.insert(d.end(), f[0], f[0] + l_name);
    for (uint32_t c : cig) {
      uint8_t b[4];
      memcpy(b, &c, 4);
      d.insert(d.end(), b, b + 4);
    }
    static const char* kNt = "=ACMGRSVTWYHKDBN";
    for (int32_t i = 0; i < l_seq; i += 2) {
      auto code = [&](char ch) {
        const char* q = strchr(kNt, toupper((unsigned char)ch));
        return (uint8_t)(q && ch ? q - kNt : 15);
      };
      d.push_back((uint8_t)((code(f[9][i]) << 4) | (i + 1 < l_seq ? code(f[9][i + 1]) : 0)));
    }
    const bool has_qual = strcmp(f[10], "*") != 0;
    const int32_t l_qual = has_qual ? (int32_t)strlen(f[10]) : 0;
    for (int32_t i = 0; i < l_seq; ++i) d.push_back(i < l_qual ? (uint8_t)(f[10][i] - 33) : 0xff);
    for (size_t t = 11; t < f.size(); ++t) {
      const char* a = f[t];
      if (strlen(a) < 5 || a[2] != ':' || a[4] != ':') continue;
      const char ty = a[3];
      const char* v = a + 5;
      if (ty == 'Z' || ty == 'H') {
        d.push_back((uint8_t)a[0]);
        d.push_back((uint8_t)a[1]);
        d.push_back((uint8_t)ty);
        d.insert(d.end(), v, v + strlen(v) + 1);
      } else if (ty == 'A') {
        d.push_back((uint8_t)a[0]);
        d.push_back((uint8_t)a[1]);
        d.push_back('A');
        d.push_back((uint8_t)v[0]);
      } else if (ty == 'i' || ty == 'f') {
        d.push_back((uint8_t)a[0]);
        d.push_back((uint8_t)a[1]);
        d.push_back((uint8_t)ty);
        uint8_t b[4];
        if (ty == 'i') {
          const int32_t x = (int32_t)strtol(v, NULL, 10);
          memcpy(b, &x, 4);
        } else {
          const float x = strtof(v, NULL);
          memcpy(b, &x, 4);
        }
        d.insert(d.end(), b, b + 4);
      }
    }
    r.tid = tid;
    r.pos = atoi(f[3]) - 1;
    r.mapq = d[9];
    r.n_cigar = ncig;
    r.flag = flag;
    r.l_seq = l_seq;
    r.off_cigar = 32 + l_name;
    r.off_seq = r.off_cigar + 4ull * ncig;
    r.off_qual = r.off_seq + (size_t)(l_seq + 1) / 2;
    r.off_aux = r.off_qual + (size_t)l_seq;
    return true;
  }
  bool next(BamRecord& r) {
    if (sam_) return next_sam(r);
    int32_t block_size = 0;
    {
      const size_t got = in_.read(&block_size, 4);
      if (got == 0) return false;  // clean EOF
      if (got < 4) fatal("%s: truncated BAM record", fn_.c_str());
    }
    if (block_size < 32) fatal("%s: malformed BAM record", fn_.c_str());
    r.data.resize((size_t)block_size);
    if (!read_exact(r.data.data(), r.data.size())) fatal("%s: truncated BAM record", fn_.c_str());
    const uint8_t* b = r.data.data();
    auto i32 = [&](size_t off) { int32_t v; memcpy(&v, b + off, 4); return v; };
    auto u16 = [&](size_t off) { uint16_t v; memcpy(&v, b + off, 2); return v; };
    r.tid = i32(0);
    r.pos = i32(4);
    const uint32_t l_read_name = b[8];
    r.mapq = b[9];
    r.n_cigar = u16(12);
    r.flag = u16(14);
    r.l_seq = i32(16);
    r.off_cigar = 32 + (size_t)l_read_name;
    r.off_seq = r.off_cigar + 4ull * r.n_cigar;
    r.off_qual = r.off_seq + (size_t)(r.l_seq < 0 ? 0 : r.l_seq + 1) / 2;
    r.off_aux = r.off_qual + (size_t)(r.l_seq < 0 ? 0 : r.l_seq);
    if (r.l_seq < 0 || l_read_name == 0 || r.off_aux > r.data.size() || b[32 + l_read_name - 1] != 0)
      fatal("%s: malformed BAM record", fn_.c_str());
    return true;
  }
  std::string fn_;
  BgzfStream in_;
  std::unique_ptr<LineReader> sam_;  // set when the input is SAM text
  bool sam_pending_ = false;         // sam_->line already holds the first alignment line
  std::unordered_map<std::string, int32_t> ref_id_;
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
  if (r.n_cigar) {
    for (uint32_t ci = 0; ci < r.n_cigar; ++ci) {
      const uint32_t c = r.cigar_at(ci);
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
      base = (rpos < rlen) ? r.base_at(rpos) : '\0';  // the reference reads the string terminator at rpos == rlen
      qual = (rpos < rlen) ? r.qual_at(rpos) : '\0';
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
  // dsc-pileup only (cmd_cram_dsc_pileup.cpp): a second flag mask applied in the loop, reads on contigs the VCF does
  // not have still count towards the droplet's totals, and the regions covered per (droplet, UMI) are collected
  bool pileup_mode = false;
  uint32_t excludeFlag2 = 0;
  bool collect_umi_loci = false;
};

// one aligned block of a read: genomeLocus (genomeLoci.h:17-75), positions 1-based and closed
struct UmiLocus {
  int32_t tid, beg1, end0;
};
// (droplet, UMI) -> covered regions, forward and reverse strand (umiLoci of cmd_cram_dsc_pileup.cpp:195)
struct UmiLoci {
  int32_t cell, umi;
  std::vector<UmiLocus> strand[2];
};

// ordering of genomeLocus::operator< (genomeLoci.h:53-75): contigs by their leading number when both have one,
// numbered contigs before the others, otherwise by name.  Two different contigs with the same leading number are
// EQUIVALENT under it -- a std::set then refuses the second locus; insert_locus below keeps that.
struct UmiLocusLess {
  const std::vector<std::string>* names;
  bool operator()(const UmiLocus& a, const UmiLocus& b) const {
    if (a.tid == b.tid) return (a.beg1 == b.beg1) ? (a.end0 < b.end0) : (a.beg1 < b.beg1);
    const std::string& ca = (*names)[(size_t)a.tid];
    const std::string& cb = (*names)[(size_t)b.tid];
    const int32_t n1 = atoi(ca.c_str()), n2 = atoi(cb.c_str());
    if (n1 == 0 && n2 == 0) return ca < cb;
    if (n1 > 0 && n2 > 0) return n1 < n2;
    return n1 > 0;
  }
};

// genomeLoci::add (genomeLoci.h:227-237) on a sorted vector standing in for the std::set
inline bool insert_locus(std::vector<UmiLocus>& v, const UmiLocus& x, const UmiLocusLess& less) {
  auto it = std::lower_bound(v.begin(), v.end(), x, less);
  if (it != v.end() && !less(x, *it)) return false;  // an equivalent element is there already
  v.insert(it, x);
  return true;
}
// genomeLoci::resolveOverlaps (genomeLoci.h:255-287): neighbours sharing a base are merged (touching ones are not)
inline void resolve_overlaps(std::vector<UmiLocus>& v, const UmiLocusLess& less) {
  size_t i = 1;
  while (i < v.size()) {
    UmiLocus& p = v[i - 1];
    const UmiLocus& c = v[i];
    if (p.tid == c.tid && p.beg1 <= c.end0 && c.beg1 <= p.end0) {
      UmiLocus m = p;
      if (c.beg1 < m.beg1) m.beg1 = c.beg1;
      if (c.end0 > m.end0) m.end0 = c.end0;
      v.erase(v.begin() + (i - 1), v.begin() + (i + 1));
      auto it = std::lower_bound(v.begin(), v.end(), m, less);
      if (it != v.end() && !less(m, *it)) {
        i = (size_t)(it - v.begin()) + 1;  // std::set::insert returned the equivalent element
      } else {
        i = (size_t)(v.insert(it, m) - v.begin()) + 1;
      }
    } else {
      ++i;
    }
  }
}

// Open-addressing indices for the hot look-ups of the join (droplet tag, UMI tag, (droplet, UMI) pair): one probe
// in a flat table instead of a node chase -- std::unordered_map<std::string, int> costs ~0.5 us per read here.
class FlatStrIndex {
 public:
  // id of `s` (ids count up from 0 in order of first appearance); *is_new tells whether it was just added
  int32_t get(const char* s, size_t n, bool* is_new) {
    if ((names_.size() + 1) * 2 > slots_.size()) grow();
    const uint64_t h = hash(s, n);
    const size_t mask = slots_.size() - 1;
    for (size_t i = (size_t)h & mask;; i = (i + 1) & mask) {
      Slot& sl = slots_[i];
      if (sl.id < 0) {
        sl.h = h;
        sl.id = (int32_t)names_.size();
        names_.emplace_back(s, n);
        *is_new = true;
        return sl.id;
      }
      if (sl.h == h) {
        const std::string& nm = names_[(size_t)sl.id];
        if (nm.size() == n && memcmp(nm.data(), s, n) == 0) {
          *is_new = false;
          return sl.id;
        }
      }
    }
  }
  const std::vector<std::string>& names() const { return names_; }

 private:
  struct Slot {
    uint64_t h;
    int32_t id;
  };
  static uint64_t hash(const char* s, size_t n) {
    uint64_t h = 0x9e3779b97f4a7c15ull ^ (n * 0xff51afd7ed558ccdull);
    while (n >= 8) {
      uint64_t w;
      memcpy(&w, s, 8);
      h = (h ^ w) * 0xc4ceb9fe1a85ec53ull;
      h ^= h >> 29;
      s += 8;
      n -= 8;
    }
    uint64_t w = 0;
    memcpy(&w, s, n);
    h = (h ^ w) * 0xc4ceb9fe1a85ec53ull;
    return h ^ (h >> 32);
  }
  void grow() {
    std::vector<Slot> old;
    old.swap(slots_);
    slots_.assign(old.empty() ? 1024 : old.size() * 2, Slot{0, -1});
    const size_t mask = slots_.size() - 1;
    for (const Slot& sl : old)
      if (sl.id >= 0) {
        size_t i = (size_t)sl.h & mask;
        while (slots_[i].id >= 0) i = (i + 1) & mask;
        slots_[i] = sl;
      }
  }
  std::vector<Slot> slots_;
  std::vector<std::string> names_;
};

class FlatU64Index {
 public:
  // value stored for `key`, or -1; put() adds a key that is not there yet
  int64_t find(uint64_t key) const {
    if (slots_.empty()) return -1;
    const size_t mask = slots_.size() - 1;
    for (size_t i = (size_t)mix(key) & mask;; i = (i + 1) & mask) {
      if (slots_[i].val < 0) return -1;
      if (slots_[i].key == key) return slots_[i].val;
    }
  }
  void put(uint64_t key, int64_t val) {
    if ((n_ + 1) * 2 > slots_.size()) grow();
    const size_t mask = slots_.size() - 1;
    size_t i = (size_t)mix(key) & mask;
    while (slots_[i].val >= 0) i = (i + 1) & mask;
    slots_[i] = {key, val};
    ++n_;
  }

 private:
  struct Slot {
    uint64_t key;
    int64_t val;
  };
  static uint64_t mix(uint64_t x) {
    x ^= x >> 33;
    x *= 0xff51afd7ed558ccdull;
    x ^= x >> 33;
    return x;
  }
  void grow() {
    std::vector<Slot> old;
    old.swap(slots_);
    slots_.assign(old.empty() ? 1024 : old.size() * 2, Slot{0, -1});
    const size_t mask = slots_.size() - 1;
    for (const Slot& sl : old)
      if (sl.val >= 0) {
        size_t i = (size_t)mix(sl.key) & mask;
        while (slots_[i].val >= 0) i = (i + 1) & mask;
        slots_[i] = sl;
      }
  }
  std::vector<Slot> slots_;
  size_t n_ = 0;
};

struct SamVariant {
  int32_t rid, pos1;  // VCF contig id, 1-based position
  char ref, alt;
  int32_t rlen;  // length of the REF allele (bcf1_t::rlen)
};

struct SamJoinStats {
  int64_t nReadsMultiSNPs = 0, nReadsSkipBCD = 0, nReadsPass = 0, nReadsRedundant = 0, nReadsN = 0, nReadsLQ = 0,
          nReadsTMP = 0, nReadsFilt = 0, n_read = 0, n_skip = 0;
};

// The streaming BAM x VCF join of cmd_cram_demuxlet.cpp:222-395 and cmd_cram_dsc_pileup.cpp:199-372, restated on
// this build's readers.  The VCF cursor runs ahead of the coordinate-sorted reads; a window [ibeg, nsnps) of the
// SNPs added so far is examined per read.  Kept as the reference has it, quirks included:
//   * a SNP becomes known only when a read's end passes the previous one, so the SNP list ends one variant past the
//     last read;
//   * the window is trimmed against the contig of the VCF CURSOR and the start of the read
//     (clear_buffer_before, bcf_filtered_reader.cpp:722-742): once the cursor has moved on to the next contig, the
//     remaining reads of the current one lose the SNPs still in the window;
//   * a SNP in the window is matched to the read by position alone.
// `vr` must stand on its first variant (read() already called once); on_snp(id, first) is called for every SNP
// right after it is appended to st.snps -- demuxlet parses genotypes there, dsc-pileup the allele frequency.
// The store comes out in the reference's iteration order (droplet id = order of first appearance, SNP id, UMI string).
template <class OnSnp>
inline void join_sam_vcf(const std::string& fn, const SamOptions& opt, VcfReader& vr, DropletStore& st, OnSnp on_snp,
                         SamJoinStats& js, std::vector<std::string>* bam_refs = NULL,
                         std::vector<UmiLoci>* umi_loci = NULL, std::vector<std::string>* umi_names = NULL,
                         std::vector<int32_t>* cell_numi = NULL) {
  if (!opt.tagGroup.empty() && opt.tagGroup.size() != 2)
    fatal("Cannot recognize group tag %s. It is suppose to be a length 2 string", opt.tagGroup.c_str());
  if (!opt.tagUMI.empty() && opt.tagUMI.size() != 2)
    fatal("Cannot recognize UMI tag %s. It is suppose to be a length 2 string", opt.tagUMI.c_str());
  BamReader br(fn);
  br.minMQ = opt.minMQ;
  br.exclude_flag = opt.exclFlag;
  if (bam_refs) *bam_refs = br.refs;
  auto name2rid = [&](const std::string& name) {  // bcf_hdr_name2id
    for (size_t i = 0; i < vr.chroms.size(); ++i)
      if (vr.chroms[i] == name) return (int32_t)i;
    return (int32_t)-1;
  };
  // contig order check (:181-199 / dsc-pileup :145-169)
  int32_t prevrid = -1, nmatch = 0, rchroms_size = 0;
  {
    std::set<int32_t> rids;
    for (size_t t = 0; t < br.refs.size(); ++t) {
      const int32_t rid = name2rid(br.refs[t]);
      if (rid >= 0) {
        if (prevrid >= rid)
          fatal("Your VCF/BCF files and SAM/BAM/CRAM files have different ordering of chromosomes. SAM/BAM/CRAM file has %s before %s, but VCF/BCF file has %s after %s",
                vr.chroms[(size_t)prevrid].c_str(), br.refs[t].c_str(), vr.chroms[(size_t)prevrid].c_str(), br.refs[t].c_str());
        rids.insert(rid);
        ++nmatch;
        rchroms_size = rid + 1;
        prevrid = rid;
      }
    }
    if (nmatch == 0 || (int32_t)rids.size() != nmatch)
      fatal("Your VCF/BCF files and SAM/BAM/CRAM files does not have any matching chromosomes, or some chromosome names are duplicated");
  }
  if (opt.pileup_mode) notice("Identified %d overlapping chromosomes between VCF and BAM", nmatch);

  std::vector<SamVariant> vars;  // parallel to st.snps
  auto add_cursor = [&](bool first) {
    vars.push_back({vr.cur.rid, vr.cur.pos1, vr.cur.ref[0], vr.cur.alt[0], (int32_t)vr.cur.ref.size()});
    st.snps.push_back({vr.cur.rid, vr.cur.pos1, vr.cur.ref[0], vr.cur.alt[0], 0.0});
    on_snp((int32_t)vars.size() - 1, first);
  };
  st.snps.clear();
  add_cursor(true);
  bool vcf_eof = false;
  int64_t ibeg = 0;  // window [ibeg, vars.size())

  struct Rec {
    int32_t cell, snp, umi;
    uint8_t al, bq;
    int64_t seq;
  };
  struct Key {
    int32_t cell, snp, umi;
    bool operator==(const Key& o) const { return cell == o.cell && snp == o.snp && umi == o.umi; }
  };
  struct KeyHash {
    size_t operator()(const Key& k) const {
      uint64_t h = (uint64_t)(uint32_t)k.cell * 0x9e3779b97f4a7c15ull;
      h ^= ((uint64_t)(uint32_t)k.snp + 0x7f4a7c15ull) * 0xbf58476d1ce4e5b9ull;
      h ^= ((uint64_t)(uint32_t)k.umi + 0x1ce4e5b9ull) * 0x94d049bb133111ebull;
      return (size_t)(h ^ (h >> 29));
    }
  };
  std::vector<Rec> recs;
  std::unordered_set<Key, KeyHash> seen;
  FlatStrIndex cell_id, umi_id;
  const std::vector<std::string>& umis = umi_id.names();
  FlatU64Index loci_at;  // (cell, umi) -> index into *umi_loci
  const UmiLocusLess less{&br.refs};
  const char* gtag = opt.tagGroup.c_str();
  const char* utag = opt.tagUMI.c_str();
  BamRecord r;
  int64_t n_no_gtag = 0, n_no_utag = 0;
  std::vector<int32_t> tid_rid;  // cache of name2rid per BAM contig, refreshed when the VCF meets a new contig
  size_t tid_rid_for = 0;
  char numbuf[32];
  while (br.read(r)) {
    const int32_t endpos = r.endpos();
    if (tid_rid_for != vr.chroms.size()) {
      tid_rid.assign(br.refs.size(), -1);
      for (size_t t = 0; t < br.refs.size(); ++t) tid_rid[t] = name2rid(br.refs[t]);
      tid_rid_for = vr.chroms.size();
    }
    const int32_t tid2rid = (r.tid >= 0 && r.tid < (int32_t)tid_rid.size()) ? tid_rid[(size_t)r.tid] : -1;
    bool noBCF = false;
    if (opt.pileup_mode) {
      if (r.flag & opt.excludeFlag2) {  // dsc-pileup :208-212
        ++js.nReadsFilt;
        continue;
      }
      if (tid2rid < 0) noBCF = true;
      if (vars.back().rid >= rchroms_size)
        fatal("The contig %s was absent in the VCF header. This happens when contigs in VCF header do not match to actual records. Run 'bcftools reheader -f [ref.fasta.fai] [in.vcf.gz] > [out.vcf.gz] to resolve this problem",
              (r.tid >= 0 && r.tid < (int32_t)br.refs.size()) ? br.refs[(size_t)r.tid].c_str() : "*");
    } else if (tid2rid < 0) {
      continue;  // demuxlet :225-228
    }
    {  // clear_buffer_before(contig of the cursor, start of the read)
      const int32_t crid = vars.back().rid;
      while (ibeg < (int64_t)vars.size()) {
        const SamVariant& v = vars[(size_t)ibeg];
        if (v.rid < crid) ++ibeg;
        else if (v.rid == crid && (v.pos1 - 1) + v.rlen < r.pos) ++ibeg;
        else break;
      }
    }
    if (!noBCF) {
      while (!vcf_eof && (vars.back().rid < tid2rid || (vars.back().rid == tid2rid && vars.back().pos1 - 1 < endpos))) {
        if (vr.read()) add_cursor(false);
        else vcf_eof = true;
      }
    }
    // droplet
    int32_t ibcd;
    {
      const char* sbcd = ".";
      if (!opt.tagGroup.empty()) {
        const char* z = r.tag_z(gtag);
        if (z) sbcd = z;
        else ++n_no_gtag;
        if (!opt.valid_bcs.empty() && !opt.valid_bcs.count(sbcd)) {
          ++js.nReadsSkipBCD;
          continue;
        }
      }
      bool fresh = false;
      ibcd = cell_id.get(sbcd, strlen(sbcd), &fresh);
      if (fresh) {
        st.bcs.push_back(sbcd);
        st.totl_reads.push_back(0);
        if (cell_numi) cell_numi->push_back(0);
      }
    }
    ++js.nReadsTMP;
    // UMI: looked up when it is needed -- for the covered regions of dsc-pileup, or once the read shows a base at a
    // SNP (most reads of a single-cell BAM overlap none).  A random UMI (--tag-UMI '') is drawn for every read,
    // as the reference does, so that the rand() sequence stays the same (:298-301).
    int32_t iumi = -1;
    auto resolve_umi = [&]() {
      const char* z = ".";
      if (opt.tagUMI.empty()) {
        snprintf(numbuf, sizeof(numbuf), "%x", rand());
        z = numbuf;
      } else {
        const char* t = r.tag_z(utag);
        if (t) z = t;
        else ++n_no_utag;
      }
      bool fresh = false;
      iumi = umi_id.get(z, strlen(z), &fresh);
    };
    if (opt.tagUMI.empty() || (opt.collect_umi_loci && umi_loci)) resolve_umi();
    ++st.totl_reads[(size_t)ibcd];
    if (opt.collect_umi_loci && umi_loci) {  // dsc-pileup :296-331
      const uint64_t lkey = ((uint64_t)(uint32_t)ibcd << 32) | (uint32_t)iumi;
      int64_t at = loci_at.find(lkey);
      if (at < 0) {
        at = (int64_t)umi_loci->size();
        loci_at.put(lkey, at);
        umi_loci->push_back({ibcd, iumi, {}});
        if (cell_numi) ++(*cell_numi)[(size_t)ibcd];
      }
      std::vector<UmiLocus>& loci = (*umi_loci)[(size_t)at].strand[(r.flag & 0x10) ? 1 : 0];
      bool grew = false;
      // (an alignment without a valid contig -- unmapped but let through by the flag masks, or a damaged record --
      // has no region to report; the reference would index its contig table with that id)
      if (r.n_cigar && r.tid >= 0 && r.tid < (int32_t)br.refs.size()) {
        int32_t cpos = r.pos;
        for (uint32_t ci = 0; ci < r.n_cigar; ++ci) {
          const uint32_t c = r.cigar_at(ci);
          const char op = "MIDNSHP=XB??????"[c & 15];
          const int32_t len = (int32_t)(c >> 4);
          if (op == 'M') {
            grew |= insert_locus(loci, {r.tid, cpos + 1, cpos + len}, less);
            cpos += len;
          } else if (op == 'D' || op == 'N') {
            cpos += len;
          }
        }
        if (grew) resolve_overlaps(loci, less);  // an unchanged, already resolved set stays as it is
      }
    }
    if (noBCF) continue;
    // SNPs of the window (:340-367)
    int32_t nv_pass = 0, nv_redundant = 0, nv_valid = 0;
    for (int64_t i = ibeg; i < (int64_t)vars.size(); ++i) {
      const SamVariant& v = vars[(size_t)i];
      char base, qual;
      int32_t rpos;
      bam_base_at(r, (uint32_t)(v.pos1 - 1), base, qual, rpos);
      if (rpos == kBamReadIndexNA) continue;
      if (base == 'N') continue;
      ++nv_valid;
      if (qual - 33 < opt.minBQ) continue;
      if (rpos < opt.minTD - 1) continue;
      if (rpos + opt.minTD > r.l_seq) continue;
      const uint8_t allele = (base == v.ref) ? 0 : ((base == v.alt) ? 1 : 2);
      const int bq = qual - 33 > opt.capBQ ? opt.capBQ : qual - 33;
      if (iumi < 0) resolve_umi();
      if (seen.insert({ibcd, (int32_t)i, iumi}).second) {  // sc_dropseq_lib_t::add_read: the first read wins
        recs.push_back({ibcd, (int32_t)i, iumi, allele, (uint8_t)bq, (int64_t)recs.size()});
        ++nv_pass;
      } else {
        ++nv_redundant;
      }
    }
    if (nv_pass > 1) ++js.nReadsMultiSNPs;
    if (nv_pass > 0) ++js.nReadsPass;
    else if (nv_redundant > 0) ++js.nReadsRedundant;
    else if (nv_valid > 0) ++js.nReadsLQ;
    else ++js.nReadsN;
  }
  js.n_read = br.n_read;
  js.n_skip = br.n_skip;
  if (n_no_utag) notice("WARNING: %lld reads carry no UMI tag %s and were treated as a single UMI", (long long)n_no_utag, utag);
  if (n_no_gtag) notice("WARNING: %lld reads carry no droplet/cell tag %s and were treated as a single group", (long long)n_no_gtag, gtag);
  if (umi_names) *umi_names = umis;

  // reference iteration order: droplet id, SNP id, UMI string
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
