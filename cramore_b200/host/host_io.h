// host_io.h -- host-side ingestion of the `--plp` path, flat and allocation-light: the digital pileup
// (.cel.gz/.var.gz/.plp.gz, format of cmd_cram_dsc_pileup.cpp:438-523) and a text VCF are turned directly
// into the CSR droplet store + float genotype matrix the C ABI takes, with the semantics of
// sc_dropseq_lib_t::load_from_plp (sc_drop_seq.cpp:103-384) and BCFFilteredReader
// (bcf_filtered_reader.cpp:406-500, 571-647, 744-870).  No map-of-maps: reads are collected as flat records
// and sorted once into the reference's iteration order.
#pragma once
#include <stdint.h>
#include <zlib.h>

#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <future>
#include <map>
#include <set>
#include <stdexcept>
#include <string>
#include <thread>
#include <vector>

namespace hostio {

[[noreturn]] inline void fatal(const char* fmt, ...) {  // Error.cpp:29-43: "FATAL ERROR -" then abort
  char buf[2048];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  fprintf(stderr, "\nFATAL ERROR - \n%s\n\n", buf);
  throw std::runtime_error(buf);
}
inline void notice(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  fprintf(stderr, "NOTICE - ");
  vfprintf(stderr, fmt, ap);
  fprintf(stderr, "\n");
  va_end(ap);
}

// BGZF (SAM spec 4.1) is a series of independent gzip members of at most 64 KiB, each announcing its compressed
// size in a 'BC' extra field: the blocks of a batch are inflated on all host cores at once and handed out as one
// byte stream.  zlib's gzread on the same file is a single sequential inflate -- for a real 10-50 GB BAM the read
// loop then IS the run time of dsc-pileup / demuxlet --sam.  A file that is not BGZF (plain gzip, or not compressed)
// goes through gzread as before.
class BgzfStream {
 public:
  explicit BgzfStream(const std::string& fn) : fn_(fn) {
    fp_ = fopen(fn.c_str(), "rb");
    if (!fp_) fatal("Cannot open file %s for reading", fn.c_str());
    unsigned char h[18];
    const size_t got = fread(h, 1, sizeof(h), fp_);
    bgzf_ = got == sizeof(h) && h[0] == 0x1f && h[1] == 0x8b && h[2] == 8 && (h[3] & 4) && h[12] == 'B' && h[13] == 'C';
    if (bgzf_) {
      fseek(fp_, 0, SEEK_SET);
      nthreads_ = (int)std::min<unsigned>(std::max(1u, std::thread::hardware_concurrency()), 32u);
      if (const char* e = getenv("CRAMORE_BGZF_THREADS")) nthreads_ = std::max(1, atoi(e));
    } else {
      fclose(fp_);
      fp_ = NULL;
      gz_ = gzopen(fn.c_str(), "rb");
      if (!gz_) fatal("Cannot open file %s for reading", fn.c_str());
      gzbuffer(gz_, 1 << 20);
    }
  }
  ~BgzfStream() {
    if (ahead_.valid()) {
      try {
        ahead_.get();
      } catch (...) {
      }
    }
    if (fp_) fclose(fp_);
    if (gz_) gzclose(gz_);
  }
  BgzfStream(const BgzfStream&) = delete;
  BgzfStream& operator=(const BgzfStream&) = delete;
  // up to n bytes; fewer only at the end of the file
  size_t read(void* dst, size_t n) {
    char* p = static_cast<char*>(dst);
    size_t done = 0;
    if (!bgzf_) {
      while (done < n) {
        const int got = gzread(gz_, p + done, (unsigned)std::min<size_t>(n - done, 1u << 30));
        if (got < 0) fatal("%s: read error", fn_.c_str());
        if (got == 0) break;
        done += (size_t)got;
      }
      return done;
    }
    while (done < n) {
      if (pos_ == out_.size() && !fill()) break;
      const size_t take = std::min(n - done, out_.size() - pos_);
      memcpy(p + done, out_.data() + pos_, take);
      pos_ += take;
      done += take;
    }
    return done;
  }

 private:
  struct Block {
    size_t in_off, in_len, out_off, out_len;  // deflate payload inside in_, inflated bytes inside out_
    uint32_t crc;
  };
  // The next batch is read and inflated by a helper thread while the caller works through the current one.
  bool fill() {
    if (!ahead_.valid()) ahead_ = std::async(std::launch::async, [this]() { return produce(next_); });
    const bool ok = ahead_.get();  // rethrows what fatal() threw in the helper
    if (!ok) return false;
    out_.swap(next_);
    pos_ = 0;
    ahead_ = std::async(std::launch::async, [this]() { return produce(next_); });
    return true;
  }
  // reads the next batch of blocks and inflates it into `out_`; false at the end of the file
  bool produce(std::vector<unsigned char>& out_) {
    constexpr size_t kBatchBlocks = 512;  // <= 32 MiB inflated
    in_.clear();
    std::vector<Block> blocks;
    size_t out_total = 0;
    unsigned char h[18];
    while (blocks.size() < kBatchBlocks) {
      const size_t got = fread(h, 1, 12, fp_);
      if (got == 0) break;
      if (got != 12 || h[0] != 0x1f || h[1] != 0x8b || h[2] != 8 || !(h[3] & 4)) fatal("%s: malformed BGZF block header", fn_.c_str());
      const size_t xlen = h[10] | ((size_t)h[11] << 8);
      std::vector<unsigned char> extra(xlen);
      if (fread(extra.data(), 1, xlen, fp_) != xlen) fatal("%s: truncated BGZF block", fn_.c_str());
      size_t bsize = 0;
      for (size_t i = 0; i + 4 <= xlen;) {
        const size_t slen = extra[i + 2] | ((size_t)extra[i + 3] << 8);
        if (extra[i] == 'B' && extra[i + 1] == 'C' && slen == 2 && i + 6 <= xlen) bsize = (extra[i + 4] | ((size_t)extra[i + 5] << 8)) + 1;
        i += 4 + slen;
      }
      if (bsize < 12 + xlen + 8) fatal("%s: BGZF block without a size field", fn_.c_str());
      const size_t payload = bsize - 12 - xlen - 8;
      const size_t off = in_.size();
      in_.resize(off + payload + 8);
      if (fread(in_.data() + off, 1, payload + 8, fp_) != payload + 8) fatal("%s: truncated BGZF block", fn_.c_str());
      uint32_t crc, isize;
      memcpy(&crc, in_.data() + off + payload, 4);
      memcpy(&isize, in_.data() + off + payload + 4, 4);
      if (isize > 65536) fatal("%s: BGZF block larger than 64 KiB", fn_.c_str());
      if (isize == 0) {  // the empty end-of-file marker (or padding): nothing to inflate
        in_.resize(off);
        continue;
      }
      blocks.push_back({off, payload, out_total, isize, crc});
      out_total += isize;
    }
    if (blocks.empty()) return false;
    out_.resize(out_total);
    std::atomic<size_t> next{0};
    std::atomic<int> bad{0};
    auto work = [&]() {
      z_stream zs;
      memset(&zs, 0, sizeof(zs));
      if (inflateInit2(&zs, -15) != Z_OK) {
        bad = 1;
        return;
      }
      for (size_t i = next++; i < blocks.size(); i = next++) {
        const Block& b = blocks[i];
        inflateReset(&zs);
        zs.next_in = in_.data() + b.in_off;
        zs.avail_in = (uInt)b.in_len;
        zs.next_out = out_.data() + b.out_off;
        zs.avail_out = (uInt)b.out_len;
        const int rc = inflate(&zs, Z_FINISH);
        if (rc != Z_STREAM_END || zs.avail_out != 0 ||
            (uint32_t)crc32(crc32(0L, Z_NULL, 0), out_.data() + b.out_off, (uInt)b.out_len) != b.crc)
          bad = 1;
      }
      inflateEnd(&zs);
    };
    const int nt = (int)std::min<size_t>((size_t)nthreads_, (blocks.size() + 15) / 16);
    if (nt <= 1) {
      work();
    } else {
      std::vector<std::thread> th;
      for (int t = 0; t < nt; ++t) th.emplace_back(work);
      for (auto& t : th) t.join();
    }
    if (bad) fatal("%s: corrupt BGZF block (inflate or CRC failed)", fn_.c_str());
    return true;
  }
  std::string fn_;
  FILE* fp_ = NULL;
  gzFile gz_ = NULL;
  bool bgzf_ = false;
  int nthreads_ = 1;
  std::vector<unsigned char> in_, out_, next_;
  std::future<bool> ahead_;
  size_t pos_ = 0;
};

// ---- line reader over plain, gzip or BGZF text, fields split on white space (tsv_reader.cpp:27-55) ----
// A BCF2 file (BGZF around the binary layout of the VCF 4.2 specification, section 6) is served as the lines of the
// equivalent text VCF: header text first, then one line per record, numbers printed so that they parse back to the
// stored value exactly (%.9g for float32).  Everything downstream -- the variant filter, parse_posteriors, INFO
// look-ups -- therefore sees a BCF exactly as it sees a VCF, and gets the values htslib's typed accessors would
// hand out.  Parity unpinned (no htslib here): checked against an independent BCF encoder in the tests.
class LineReader {
 public:
  explicit LineReader(const std::string& fn) : nlines(0), fn_(fn), in_(fn), buf_(1 << 20), bpos_(0), blen_(0) {
    blen_ = in_.read(buf_.data(), buf_.size());
    if (blen_ >= 9 && memcmp(buf_.data(), "BCF\2\2", 5) == 0) open_bcf();
  }
  bool next_line() {
    line.clear();
    if (bcf_) {
      if (hdr_next_ < hdr_lines_.size()) line = hdr_lines_[hdr_next_++];
      else if (!bcf_record()) return false;
      ++nlines;
      return true;
    }
    for (;;) {
      if (bpos_ == blen_) {
        blen_ = in_.read(buf_.data(), buf_.size());
        bpos_ = 0;
        if (blen_ == 0) break;
      }
      const char* p = buf_.data() + bpos_;
      const char* nl = static_cast<const char*>(memchr(p, '\n', blen_ - bpos_));
      if (nl) {
        line.append(p, (size_t)(nl - p) + 1);
        bpos_ += (size_t)(nl - p) + 1;
        break;
      }
      line.append(p, blen_ - bpos_);
      bpos_ = blen_;
    }
    if (line.empty()) return false;
    while (!line.empty() && (line.back() == '\n' || line.back() == '\r')) line.pop_back();
    ++nlines;
    return true;
  }
  // number of white-space separated fields of the next line, 0 at EOF
  int read_fields() {
    fields.clear();
    if (!next_line()) return 0;
    char* p = &line[0];
    char* end = p + line.size();
    while (p < end) {
      while (p < end && isspace((unsigned char)*p)) ++p;
      if (p >= end) break;
      fields.push_back(p);
      while (p < end && !isspace((unsigned char)*p)) ++p;
      if (p < end) *p++ = '\0';
    }
    return (int)fields.size();
  }
  std::string line;
  std::vector<char*> fields;
  int64_t nlines;

 private:
  // bytes of the underlying stream, continuing after whatever next_line() has consumed
  size_t raw(void* dst, size_t n) {
    char* p = static_cast<char*>(dst);
    size_t done = 0;
    while (done < n) {
      if (bpos_ == blen_) {
        blen_ = in_.read(buf_.data(), buf_.size());
        bpos_ = 0;
        if (blen_ == 0) break;
      }
      const size_t take = std::min(n - done, blen_ - bpos_);
      memcpy(p + done, buf_.data() + bpos_, take);
      bpos_ += take;
      done += take;
    }
    return done;
  }
  // ---- BCF2 ----
  void open_bcf() {
    bcf_ = true;
    bpos_ = 5;
    uint32_t l_text = 0;
    if (raw(&l_text, 4) != 4) fatal("%s: truncated BCF header", fn_.c_str());
    std::string text(l_text, '\0');
    if (raw(&text[0], l_text) != l_text) fatal("%s: truncated BCF header", fn_.c_str());
    while (!text.empty() && text.back() == '\0') text.pop_back();
    size_t p = 0;
    while (p < text.size()) {
      size_t e = text.find('\n', p);
      if (e == std::string::npos) e = text.size();
      if (e > p) hdr_lines_.push_back(text.substr(p, e - p));
      p = e + 1;
    }
    // dictionaries (VCF spec 6.2.1): FILTER / INFO / FORMAT ids share one, PASS is entry 0; contigs have their own.
    // An IDX= field fixes the position, otherwise ids count up in order of first appearance.
    auto put = [](std::vector<std::string>& dict, size_t idx, const std::string& id) {
      if (dict.size() <= idx) dict.resize(idx + 1);
      dict[idx] = id;
    };
    str_dict_.assign(1, "PASS");
    for (const std::string& l : hdr_lines_) {
      const bool is_str = l.compare(0, 10, "##FILTER=<") == 0 || l.compare(0, 8, "##INFO=<") == 0 || l.compare(0, 10, "##FORMAT=<") == 0;
      const bool is_ctg = l.compare(0, 10, "##contig=<") == 0;
      if (!is_str && !is_ctg) continue;
      const size_t a = l.find("ID=");
      if (a == std::string::npos) continue;
      const std::string id = l.substr(a + 3, l.find_first_of(",>", a + 3) - a - 3);
      const size_t x = l.find("IDX=");
      std::vector<std::string>& dict = is_ctg ? ctg_dict_ : str_dict_;
      if (x != std::string::npos) {
        put(dict, (size_t)atol(l.c_str() + x + 4), id);
      } else if (std::find(dict.begin(), dict.end(), id) == dict.end()) {
        dict.push_back(id);
      }
    }
  }
  struct Typed {
    int type;     // 0 missing-typed, 1 int8, 2 int16, 3 int32, 5 float, 7 char
    size_t n;     // elements
  };
  static size_t type_size(int t) { return t == 1 || t == 7 ? 1 : t == 2 ? 2 : (t == 3 || t == 5) ? 4 : 0; }
  Typed read_descriptor(const uint8_t*& p, const uint8_t* end) {
    if (p >= end) fatal("%s: malformed BCF record", fn_.c_str());
    const uint8_t b = *p++;
    Typed t{b & 15, (size_t)(b >> 4)};
    if (t.n == 15) {  // the real length follows as a typed integer
      const Typed lt = read_descriptor(p, end);
      int64_t v = 0;
      if (!read_int(p, end, lt.type, &v) || v < 0) fatal("%s: malformed BCF record", fn_.c_str());
      t.n = (size_t)v;
    }
    return t;
  }
  // one integer of the given type; false for "missing" and "end of vector" (flag says which)
  bool read_int(const uint8_t*& p, const uint8_t* end, int type, int64_t* out, bool* eov = NULL) {
    const size_t sz = type_size(type);
    if (sz == 0 || type == 5 || type == 7 || p + sz > end) fatal("%s: malformed BCF record", fn_.c_str());
    int64_t v = 0;
    int64_t miss, stop;
    if (type == 1) { int8_t x; memcpy(&x, p, 1); v = x; miss = INT8_MIN; stop = INT8_MIN + 1; }
    else if (type == 2) { int16_t x; memcpy(&x, p, 2); v = x; miss = INT16_MIN; stop = INT16_MIN + 1; }
    else { int32_t x; memcpy(&x, p, 4); v = x; miss = INT32_MIN; stop = (int64_t)INT32_MIN + 1; }
    p += sz;
    if (eov) *eov = (v == stop);
    if (v == miss || v == stop) return false;
    *out = v;
    return true;
  }
  // text of a typed vector (comma separated; "." for missing); stops at an end-of-vector value
  void append_values(std::string& out, const uint8_t*& p, const uint8_t* end, Typed t) {
    char num[48];
    if (t.type == 7) {
      if (p + t.n > end) fatal("%s: malformed BCF record", fn_.c_str());
      size_t n = 0;
      while (n < t.n && p[n] != 0) ++n;
      if (n == 0) out += '.';
      else out.append(reinterpret_cast<const char*>(p), n);
      p += t.n;
      return;
    }
    bool any = false, stopped = false;
    for (size_t i = 0; i < t.n; ++i) {
      if (t.type == 5) {
        if (p + 4 > end) fatal("%s: malformed BCF record", fn_.c_str());
        uint32_t bits;
        memcpy(&bits, p, 4);
        p += 4;
        if (stopped || bits == 0x7F800002u) { stopped = true; continue; }
        if (any) out += ',';
        if (bits == 0x7F800001u) out += '.';
        else {
          float f;
          memcpy(&f, &bits, 4);
          snprintf(num, sizeof(num), "%.9g", (double)f);
          out += num;
        }
        any = true;
      } else {
        int64_t v = 0;
        bool eov = false;
        const bool ok = read_int(p, end, t.type, &v, &eov);
        if (stopped || eov) { stopped = true; continue; }
        if (any) out += ',';
        if (!ok) out += '.';
        else {
          snprintf(num, sizeof(num), "%lld", (long long)v);
          out += num;
        }
        any = true;
      }
    }
    if (!any) out += '.';
  }
  const std::string& dict_name(const std::vector<std::string>& d, int64_t i) {
    if (i < 0 || (size_t)i >= d.size() || d[(size_t)i].empty()) fatal("%s: BCF record refers to an id the header does not define", fn_.c_str());
    return d[(size_t)i];
  }
  bool bcf_record() {
    uint32_t len[2];
    const size_t got = raw(len, 8);
    if (got == 0) return false;
    if (got != 8 || len[0] < 24) fatal("%s: truncated BCF record", fn_.c_str());
    rec_.resize((size_t)len[0] + len[1]);
    if (raw(rec_.data(), rec_.size()) != rec_.size()) fatal("%s: truncated BCF record", fn_.c_str());
    const uint8_t* p = rec_.data();
    const uint8_t* end = p + len[0];
    int32_t chrom, pos, rlen;
    uint32_t qbits, n_allele_info, n_fmt_sample;
    memcpy(&chrom, p, 4); memcpy(&pos, p + 4, 4); memcpy(&rlen, p + 8, 4); memcpy(&qbits, p + 12, 4);
    memcpy(&n_allele_info, p + 16, 4); memcpy(&n_fmt_sample, p + 20, 4);
    p += 24;
    const uint32_t n_allele = n_allele_info >> 16, n_info = n_allele_info & 0xffff;
    const uint32_t n_fmt = n_fmt_sample >> 24, n_sample = n_fmt_sample & 0xffffff;
    char num[64];
    line = dict_name(ctg_dict_, chrom);
    snprintf(num, sizeof(num), "\t%d\t", pos + 1);
    line += num;
    append_values(line, p, end, read_descriptor(p, end));  // ID
    for (uint32_t a = 0; a < n_allele; ++a) {                // REF, then ALT comma separated
      line += (a == 0) ? '\t' : (a == 1 ? '\t' : ',');
      append_values(line, p, end, read_descriptor(p, end));
    }
    if (n_allele < 2) line += "\t.";
    if (qbits == 0x7F800001u) line += "\t.";
    else {
      float q;
      memcpy(&q, &qbits, 4);
      snprintf(num, sizeof(num), "\t%.9g", (double)q);
      line += num;
    }
    {  // FILTER
      const Typed t = read_descriptor(p, end);
      line += '\t';
      if (t.n == 0) line += '.';
      for (size_t i = 0; i < t.n; ++i) {
        int64_t v = 0;
        if (!read_int(p, end, t.type, &v)) continue;
        if (i) line += ';';
        line += dict_name(str_dict_, v);
      }
    }
    line += '\t';
    if (n_info == 0) line += '.';
    for (uint32_t i = 0; i < n_info; ++i) {
      const Typed kt = read_descriptor(p, end);
      int64_t key = 0;
      if (kt.n != 1 || !read_int(p, end, kt.type, &key)) fatal("%s: malformed BCF INFO key", fn_.c_str());
      if (i) line += ';';
      line += dict_name(str_dict_, key);
      const Typed vt = read_descriptor(p, end);
      if (vt.n == 0) continue;  // a flag
      line += '=';
      append_values(line, p, end, vt);
    }
    if (n_fmt == 0 || n_sample == 0) return true;
    // FORMAT: one typed matrix [n_sample][n] per key; columns are gathered per sample afterwards
    p = rec_.data() + len[0];
    end = p + len[1];
    std::vector<std::string> keys(n_fmt);
    cols_.assign((size_t)n_fmt * n_sample, std::string());
    for (uint32_t f = 0; f < n_fmt; ++f) {
      const Typed kt = read_descriptor(p, end);
      int64_t key = 0;
      if (kt.n != 1 || !read_int(p, end, kt.type, &key)) fatal("%s: malformed BCF FORMAT key", fn_.c_str());
      keys[f] = dict_name(str_dict_, key);
      const Typed vt = read_descriptor(p, end);
      const bool is_gt = keys[f] == "GT";
      for (uint32_t sm = 0; sm < n_sample; ++sm) {
        std::string& out = cols_[(size_t)f * n_sample + sm];
        if (!is_gt) {
          append_values(out, p, end, vt);
          continue;
        }
        bool any = false, stopped = false;  // (allele + 1) << 1 | phased; 0 = missing allele
        for (size_t i = 0; i < vt.n; ++i) {
          int64_t v = 0;
          bool eov = false;
          const bool ok = read_int(p, end, vt.type, &v, &eov);
          if (stopped || eov) { stopped = true; continue; }
          if (!ok) v = 0;
          if (any) out += (v & 1) ? '|' : '/';
          if ((v >> 1) == 0) out += '.';
          else {
            snprintf(num, sizeof(num), "%lld", (long long)((v >> 1) - 1));
            out += num;
          }
          any = true;
        }
        if (!any) out += '.';
      }
    }
    line += '\t';
    for (uint32_t f = 0; f < n_fmt; ++f) {
      if (f) line += ':';
      line += keys[f];
    }
    for (uint32_t sm = 0; sm < n_sample; ++sm) {
      line += '\t';
      for (uint32_t f = 0; f < n_fmt; ++f) {
        if (f) line += ':';
        line += cols_[(size_t)f * n_sample + sm];
      }
    }
    return true;
  }
  std::string fn_;
  BgzfStream in_;
  std::vector<char> buf_;
  size_t bpos_, blen_;
  bool bcf_ = false;
  std::vector<std::string> hdr_lines_, str_dict_, ctg_dict_, cols_;
  size_t hdr_next_ = 0;
  std::vector<uint8_t> rec_;
};

// ---- gz / plain writer with printf (hts_utils.cpp:1018-1038) ----
class Writer {
 public:
  Writer(const std::string& fn, bool gz) : fp_(NULL), gz_(NULL), fn_(fn) {
    if (gz) gz_ = gzopen(fn.c_str(), "wb");
    else fp_ = fopen(fn.c_str(), "w");
    if (!fp_ && !gz_) fatal("Cannot open file %s for writing", fn.c_str());
  }
  ~Writer() { close(); }
  // A short write or a failing close (disk full, quota) is fatal: a truncated output file must not exit 0.
  void close() {
    if (fp_ && fclose(fp_) != 0) {
      fp_ = NULL;
      fatal("Cannot finish writing %s", fn_.c_str());
    }
    if (gz_ && gzclose(gz_) != Z_OK) {
      gz_ = NULL;
      fatal("Cannot finish writing %s", fn_.c_str());
    }
    fp_ = NULL;
    gz_ = NULL;
  }
  void printf(const char* fmt, ...) {
    char buf[8192];
    va_list ap;
    va_start(ap, fmt);
    int n = vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    if (n < 0) fatal("Cannot format a line of %s", fn_.c_str());
    if (n < (int)sizeof(buf)) {
      write(buf, n);
      return;
    }
    std::vector<char> big((size_t)n + 1);  // a line longer than the stack buffer: format it again, in full
    va_start(ap, fmt);
    vsnprintf(big.data(), big.size(), fmt, ap);
    va_end(ap);
    write(big.data(), n);
  }
  void write(const char* s, int n) {
    if (n <= 0) return;
    const bool ok = fp_ ? fwrite(s, 1, (size_t)n, fp_) == (size_t)n : gzwrite(gz_, s, (unsigned)n) == n;
    if (!ok) fatal("Cannot write to %s", fn_.c_str());
  }

 private:
  FILE* fp_;
  gzFile gz_;
  std::string fn_;
};

// ---- text VCF reader: sequential cursor with the variant filter of cmdCramDemuxlet ----
// strtof for the plain decimals a VCF holds ("0.97", "1", "3.3e-05"), bit-identical to strtof: up to 15 digits
// and a small power of ten go through Clinger's exact double path (double(m) * or / 10^k is the correctly rounded
// double) and one cast; a value whose double sits within 2 ulp of a float rounding boundary, or anything else
// (long mantissas, inf/nan, hex, large exponents), is handed to strtof itself.
inline float fast_strtof(const char* s, const char** end) {
  static const double p10[] = {1e0, 1e1, 1e2, 1e3, 1e4, 1e5, 1e6, 1e7, 1e8, 1e9, 1e10, 1e11,
                               1e12, 1e13, 1e14, 1e15, 1e16, 1e17, 1e18, 1e19, 1e20, 1e21, 1e22};
  const char* p = s;
  bool neg = false;
  if (*p == '-' || *p == '+') neg = (*p++ == '-');
  uint64_t m = 0;
  int nd = 0, dexp = 0;
  bool any = false;
  while (*p >= '0' && *p <= '9') {
    any = true;
    if (nd < 15) {
      m = m * 10 + (uint64_t)(*p - '0');
      if (m) ++nd;
    } else {
      ++dexp;
    }
    ++p;
  }
  bool lossy = dexp > 0;
  if (*p == '.') {
    ++p;
    while (*p >= '0' && *p <= '9') {
      any = true;
      if (nd < 15) {
        m = m * 10 + (uint64_t)(*p - '0');
        if (m) ++nd;
        --dexp;
      } else if (*p != '0') {
        lossy = true;
      }
      ++p;
    }
  }
  if (any && (*p == 'e' || *p == 'E')) {
    const char* q = p + 1;
    bool eneg = false;
    if (*q == '-' || *q == '+') eneg = (*q++ == '-');
    if (*q >= '0' && *q <= '9') {
      int ev = 0;
      while (*q >= '0' && *q <= '9' && ev < 10000) ev = ev * 10 + (*q++ - '0');
      dexp += eneg ? -ev : ev;
      p = q;
    }
  }
  if (any && !lossy && dexp >= -22 && dexp <= 22) {
    const double v = (dexp < 0) ? (double)m / p10[-dexp] : (double)m * p10[dexp];
    uint64_t bits;
    memcpy(&bits, &v, 8);
    const int64_t low = (int64_t)(bits & 0x1FFFFFFFull) - 0x10000000ll;  // distance to the float rounding boundary
    const int ex = (int)((bits >> 52) & 0x7ff);
    if ((low > 2 || low < -2) && ex > 1023 - 120 && ex < 1023 + 120) {  // normal float range, unambiguous cast
      *end = p;
      const float f = (float)v;
      return neg ? -f : f;
    }
    if (m == 0) {
      *end = p;
      return neg ? -0.0f : 0.0f;
    }
  }
  char* e;
  const float f = strtof(s, &e);
  *end = e;
  return f;
}

struct VcfRecord {
  int32_t rid = -1, pos1 = 0;  // 1-based position
  std::string chrom, ref, alt;
  int32_t n_allele = 0;
  std::string info;
  std::vector<std::string> fmt;
  std::vector<uint32_t> col_off;  // sample columns: offsets into line (NUL-terminated fields)
  std::string line;
  // filled by the parsers (per record, so that a batch of records can be parsed on several threads)
  std::vector<double> acs;
  int32_t an = 0;
  std::vector<int32_t> gts;
  std::vector<float> gp;   // posteriors of the pre-fetch field, [nv*3]
  bool gp_ok = false, pass = false;
  std::string error;       // text of a fatal() raised while this record was parsed ahead
};

// Sequential cursor over the records that pass the variant filter.  Records are read in batches and tokenised,
// filtered and (when a pre-fetch field is set) turned into float posteriors on all host cores at once; the cursor
// then serves them in file order, so callers see exactly the sequence of the one-record-at-a-time reader this
// replaced -- the parse of a 64-sample line costs ~10 us, 200,000 of them were most of a C2 run.
class VcfReader {
 public:
  std::vector<std::string> chroms, samples;
  std::vector<int32_t> sm_icols;  // selected sample columns, in VCF order
  int32_t minMAC = 1, maxAlleles = 2;
  double minCallRate = 0.5;
  int64_t nRead = 0, nSkip = 0;
  bool eof = false;
  VcfRecord cur;
  std::vector<double> acs;  // allele counts / number of the current record (after read(), parse_posteriors("GT"/"PL"))
  int32_t an = 0;

  VcfReader(const std::string& fn, const std::set<std::string>& sm_ids, bool allow_no_samples = false) : lr_(fn) {
    while (lr_.next_line()) {
      const std::string& l = lr_.line;
      if (l.compare(0, 2, "##") == 0) {
        if (l.compare(0, 13, "##contig=<ID=") == 0) chroms.push_back(l.substr(13, l.find_first_of(",>", 13) - 13));
        continue;
      }
      if (l.compare(0, 6, "#CHROM") == 0) {
        size_t p = 0;
        int col = 0;
        while (p <= l.size()) {
          size_t e = l.find('\t', p);
          if (e == std::string::npos) e = l.size();
          if (col >= 9) samples.push_back(l.substr(p, e - p));
          ++col;
          p = e + 1;
        }
        break;
      }
      fatal("Malformed VCF header in %s", fn.c_str());
    }
    for (int32_t i = 0; i < (int32_t)samples.size(); ++i)
      if (sm_ids.empty() || sm_ids.count(samples[i])) sm_icols.push_back(i);
    if (sm_icols.empty() && !allow_no_samples) fatal("No sample to read from %s", fn.c_str());
    nthreads_ = (int)std::min<unsigned>(std::max(1u, std::thread::hardware_concurrency()), 32u);
    if (const char* e = getenv("CRAMORE_VCF_THREADS")) nthreads_ = std::max(1, atoi(e));
  }

  int32_t nsamples() const { return (int32_t)sm_icols.size(); }
  const char* sample_id(int32_t i) const { return samples[sm_icols[i]].c_str(); }

  int chrom_id(const std::string& name) {
    for (size_t i = 0; i < chroms.size(); ++i)
      if (chroms[i] == name) return (int)i;
    chroms.push_back(name);
    return (int)chroms.size() - 1;
  }

  // The posteriors most records will be asked for: computed ahead, in parallel, for every record that passes the
  // filter.  parse_posteriors() with other arguments still works (it computes on the spot).
  void prefetch(const char* field, double gt_error) {
    prefetch_field_ = field;
    prefetch_gt_error_ = gt_error;
  }

  // BCFFilteredReader::read(): advance to the next variant passing the filter; false at EOF
  bool read() {
    for (;;) {
      if (next_ == batch_n_ && !fill_batch()) {
        eof = true;
        return false;
      }
      VcfRecord& r = batch_[next_++];
      if (!r.error.empty()) throw std::runtime_error(r.error);  // fatal() already printed it
      r.rid = chrom_id(r.chrom);
      ++nRead;
      if (!r.pass) {
        ++nSkip;
        continue;
      }
      std::swap(cur, r);
      acs = cur.acs;
      an = cur.an;
      return true;
    }
  }

  // posteriors of the selected samples for the current record: out[nv*3] (float, bcf_filtered_reader.h:168-170)
  bool parse_posteriors(const char* field, double gt_error, float* out) {
    if (cur.gp_ok && prefetch_field_ == field && prefetch_gt_error_ == gt_error) {
      memcpy(out, cur.gp.data(), sizeof(float) * cur.gp.size());
      return true;
    }
    if (cur.n_allele != 2) fatal("parse_posteriors: only biallelic sites are supported (%s:%d)", chroms[cur.rid].c_str(), cur.pos1);
    const bool ok = compute_posteriors(cur, field, gt_error, out);
    acs = cur.acs;
    an = cur.an;
    return ok;
  }

  bool info_float(const char* tag, float* v) const {
    const char* s = info_value(tag);
    if (s) *v = strtof(s, NULL);
    return s != NULL;
  }
  // start of the value of an INFO key, or NULL
  const char* info_value(const char* tag) const {
    const std::string key = std::string(tag) + "=";
    size_t p = 0;
    while (p < cur.info.size()) {
      size_t e = cur.info.find(';', p);
      if (e == std::string::npos) e = cur.info.size();
      if (cur.info.compare(p, key.size(), key) == 0) return cur.info.c_str() + p + key.size();
      p = e + 1;
    }
    return NULL;
  }

  // BCFFilteredReader::calculate_af (bcf_filtered_reader.cpp:845-879): INFO/AF, else INFO/AC / INFO/AN, when all
  // samples of the file are in use; otherwise the alternate allele's share among the selected samples' genotypes
  double calculate_af(bool use_info_field) {
    if (use_info_field && sm_icols.size() == samples.size()) {
      float af;
      if (info_float("AF", &af)) return af;
      const char* ac = info_value("AC");
      const char* an_ = info_value("AN");
      if (!ac) fatal("Cannot find AC field from INFO field at %s:%d", chroms[cur.rid].c_str(), cur.pos1);
      if (!an_) fatal("Cannot find AN field from INFO field at %s:%d", chroms[cur.rid].c_str(), cur.pos1);
      return (double)strtol(ac, NULL, 10) / (double)strtol(an_, NULL, 10);
    }
    parse_genotypes(cur);
    acs = cur.acs;
    an = cur.an;
    return (double)acs[1] / (double)an;
  }

 private:
  // ---- batch: read B lines, parse them on nthreads_ threads ----
  bool fill_batch() {
    constexpr size_t kBatch = 2048;
    if (batch_.size() < kBatch) batch_.resize(kBatch);
    batch_n_ = 0;
    next_ = 0;
    while (batch_n_ < kBatch && lr_.next_line()) {
      if (lr_.line.empty() || lr_.line[0] == '#') continue;
      VcfRecord& r = batch_[batch_n_++];
      r.line.swap(lr_.line);
      r.error.clear();
      r.gp_ok = false;
      r.pass = false;
    }
    if (batch_n_ == 0) return false;
    std::atomic<size_t> next{0};
    auto work = [&]() {
      for (size_t i = next++; i < batch_n_; i = next++) {
        VcfRecord& r = batch_[i];
        try {
          parse_record(r);
        } catch (const std::exception& e) {
          r.error = e.what();
          if (r.error.empty()) r.error = "malformed VCF record";
        }
      }
    };
    const int nt = (int)std::min<size_t>((size_t)nthreads_, (batch_n_ + 63) / 64);
    if (nt <= 1) {
      work();
    } else {
      std::vector<std::thread> th;
      for (int t = 0; t < nt; ++t) th.emplace_back(work);
      for (auto& t : th) t.join();
    }
    return true;
  }
  // everything about one record that does not depend on the records before it
  void parse_record(VcfRecord& r) const {
    tokenize(r);
    // passed_vfilter (bcf_filtered_reader.cpp:571-645): with no genotype-based threshold set (require_GT false,
    // bcf_filter_arg.h:109-113) it returns true before it ever looks at the allele count
    const bool require_GT = minMAC > 0 || minCallRate > 0;
    bool pass = !require_GT || r.n_allele <= maxAlleles;
    if (pass && require_GT) {  // vfilt.require_GT, bcf_filtered_reader.cpp:613-645
      if (!parse_genotypes(r)) fatal("Cannot find the field GT from the VCF file at position %s:%d", r.chrom.c_str(), r.pos1);
      if (minCallRate > (double)r.an / (2.0 * (double)sm_icols.size())) pass = false;
      const int32_t ac = r.an - (int32_t)r.acs[0];
      if (ac < minMAC || r.an - ac < minMAC) pass = false;
    }
    r.pass = pass;
    if (pass && !prefetch_field_.empty() && r.n_allele == 2 && !sm_icols.empty()) {
      r.gp.resize((size_t)nsamples() * 3);
      // the filter's allele counts are what read() hands out; the pre-fetch (whose PL branch re-estimates them)
      // must not disturb them
      const std::vector<double> acs_keep = r.acs;
      const int32_t an_keep = r.an;
      r.gp_ok = compute_posteriors(r, prefetch_field_.c_str(), prefetch_gt_error_, r.gp.data());
      r.acs = acs_keep;
      r.an = an_keep;
    }
  }
  void tokenize(VcfRecord& r) const {
    std::vector<uint32_t> f;
    char* base = &r.line[0];
    f.push_back(0);
    for (char* p = base; *p; ++p)
      if (*p == '\t') {
        *p = '\0';
        f.push_back((uint32_t)(p + 1 - base));
      }
    if (f.size() < 8) fatal("Malformed VCF record (%zu columns)", f.size());
    r.chrom = base + f[0];
    r.pos1 = atoi(base + f[1]);
    r.ref = base + f[3];
    r.alt = base + f[4];
    r.n_allele = (r.alt == ".") ? 1 : 2 + (int32_t)std::count(r.alt.begin(), r.alt.end(), ',');
    r.info = base + f[7];
    r.fmt.clear();
    r.col_off.clear();
    if (f.size() > 9) {
      const char* q = base + f[8];
      while (true) {
        const char* e = strchr(q, ':');
        if (!e) {
          r.fmt.push_back(q);
          break;
        }
        r.fmt.push_back(std::string(q, e - q));
        q = e + 1;
      }
      r.col_off.assign(f.begin() + 9, f.end());
    }
  }
  static int fmt_index(const VcfRecord& r, const char* key) {
    for (size_t i = 0; i < r.fmt.size(); ++i)
      if (r.fmt[i] == key) return (int)i;
    return -1;
  }
  static const char* sample_field(const VcfRecord& r, int col, int idx) {
    const char* p = r.line.data() + r.col_off[(size_t)col];
    for (int i = 0; i < idx; ++i) {
      p = strchr(p, ':');
      if (!p) return ".";
      ++p;
    }
    return p;
  }
  // bcf_filtered_reader.cpp:198-254: allele counts over the selected samples
  bool parse_genotypes(VcfRecord& r) const {
    const int gi = fmt_index(r, "GT");
    if (gi < 0) return false;
    const int32_t nv = nsamples();
    r.gts.resize(2 * (size_t)nv);
    r.acs.assign((size_t)r.n_allele, 0.0);
    r.an = 0;
    for (int32_t i = 0; i < nv; ++i) {
      const char* p = sample_field(r, sm_icols[i], gi);
      int a[2] = {-1, -1};
      int k = 0;
      while (*p && *p != ':' && k < 2) {
        if (*p == '.') {
          ++p;
        } else if (isdigit((unsigned char)*p)) {
          a[k] = (int)strtol(p, (char**)&p, 10);
        } else {
          break;
        }
        ++k;
        if (*p == '/' || *p == '|') ++p;
      }
      r.gts[2 * (size_t)i] = a[0];
      r.gts[2 * (size_t)i + 1] = a[1];
      for (int j = 0; j < 2; ++j)
        if (a[j] >= 0 && a[j] < r.n_allele) {
          ++r.an;
          ++r.acs[(size_t)a[j]];
        }
    }
    return true;
  }
  // BCFFilteredReader::parse_posteriors (bcf_filtered_reader.cpp:406-500) for a biallelic record
  bool compute_posteriors(VcfRecord& r, const char* field, double gt_error, float* out) const {
    const int32_t nv = nsamples();
    if (strcmp(field, "GT") == 0) {  // :414-452
      if (!parse_genotypes(r)) return false;
      for (int32_t i = 0; i < nv; ++i) {
        const int32_t a1 = r.gts[2 * (size_t)i], a2 = r.gts[2 * (size_t)i + 1];
        const int32_t g = (a1 < 0 || a2 < 0) ? -1 : (a1 > a2 ? a1 * (a1 + 1) / 2 + a2 : a2 * (a2 + 1) / 2 + a1);
        if (g < 0) {
          for (int32_t j = 0, l = 0; j < 2; ++j)
            for (int32_t k = 0; k <= j; ++k, ++l)
              out[i * 3 + l] = (float)((j == k ? 1.0 : 2.0) * (r.acs[j] + 1.0 / 2) / (r.an + 1.0) * (r.acs[k] + 1.0 / 2) / (r.an + 1.0));
        } else {
          for (int32_t j = 0; j < 3; ++j) out[i * 3 + j] = (g == j) ? 1.0 - gt_error : gt_error / (3 - 1.0);
        }
      }
      return true;
    }
    if (strcmp(field, "PL") == 0) {  // parse_likelihoods, bcf_filtered_reader.cpp:289-365 (diploid samples)
      // ten EM iterations for the allele frequency under HWE; the posteriors of the last one are the output.
      // toProb(pl) = pow(0.1, pl * 0.1), capped at pl = 255 (PhredHelper.cpp:31, PhredHelper.h:40).
      // Parity unpinned: the reference's own reader is htslib-bound and not buildable here.
      const int fi = fmt_index(r, "PL");
      if (fi < 0) return false;
      std::vector<uint32_t> pls((size_t)nv * 3);
      for (int32_t i = 0; i < nv; ++i) {
        const char* p = sample_field(r, sm_icols[i], fi);
        for (int j = 0; j < 3; ++j) {
          char* e;
          long v = (*p == '.') ? 255 : strtol(p, &e, 10);  // a missing value carries no information either way
          if (*p == '.') e = const_cast<char*>(p) + 1;
          pls[(size_t)i * 3 + j] = (uint32_t)(v < 0 ? 0 : v);
          p = (*e == ',') ? e + 1 : e;
        }
      }
      r.acs.assign(2, 0.5);
      double gp[3];
      for (int it = 0; it < 10; ++it) {
        double newacs[2] = {0, 0};
        r.an = 0;
        for (int32_t i = 0; i < nv; ++i) {
          double sumgp = 0;
          for (int j = 0, l = 0; j < 2; ++j)
            for (int k = 0; k <= j; ++k, ++l) {
              const uint32_t pl = pls[(size_t)i * 3 + l];
              sumgp += (gp[l] = (j == k ? 1 : 2) * r.acs[j] * r.acs[k] * pow(0.1, (pl > 255 ? 255 : pl) * 0.1));
            }
          for (int j = 0, l = 0; j < 2; ++j)
            for (int k = 0; k <= j; ++k, ++l) {
              gp[l] /= sumgp;
              newacs[j] += gp[l];
              newacs[k] += gp[l];
            }
          r.an += 2;
          if (it + 1 == 10)
            for (int l = 0; l < 3; ++l) out[i * 3 + l] = (float)gp[l];
        }
        for (int j = 0; j < 2; ++j) r.acs[j] = newacs[j] / r.an;
      }
      for (int j = 0; j < 2; ++j) r.acs[j] *= r.an;
      return true;
    }
    const int fi = fmt_index(r, field);
    if (fi < 0) return false;
    float gpSums[3] = {1.0f / 4.0f, 2.0f / 4.0f, 1.0f / 4.0f};  // :462-468 (nalleles = 2)
    for (int32_t i = 0; i < nv; ++i) {  // :474-485
      const char* p = sample_field(r, sm_icols[i], fi);
      float g[3];
      for (int j = 0; j < 3; ++j) {
        const char* e;
        g[j] = fast_strtof(p, &e);
        p = (*e == ',') ? e + 1 : e;
      }
      float sumgp = 0;
      for (int j = 0; j < 3; ++j) sumgp += g[j];
      for (int j = 0; j < 3; ++j) {
        g[j] /= sumgp;
        gpSums[j] += g[j];
        out[i * 3 + j] = g[j];
      }
    }
    for (int j = 0; j < 3; ++j) gpSums[j] /= (int32_t)(nv + 1.0);
    for (int32_t i = 0; i < nv; ++i)  // :490-496
      for (int j = 0; j < 3; ++j) out[i * 3 + j] = ((1.0 - gt_error) * out[i * 3 + j] + gt_error * gpSums[j]);
    return true;
  }
  LineReader lr_;
  std::vector<VcfRecord> batch_;
  size_t batch_n_ = 0, next_ = 0;
  int nthreads_ = 1;
  std::string prefetch_field_;
  double prefetch_gt_error_ = 0;
};

// ---- droplet store flattened to CSR ----
struct SnpInfo {
  int32_t rid, pos;
  char ref, alt;
  double af;
};

struct DropletStore {
  // droplets in id order (= .cel.gz row order minus skipped rows, sc_drop_seq.cpp:182-185)
  std::vector<std::string> bcs;
  std::vector<int32_t> totl_reads, uniq_reads;  // cell_totl_reads / cell_uniq_reads
  std::vector<std::string> rid2chr;
  std::vector<SnpInfo> snps;
  // CSR
  std::vector<int64_t> cell_ptr, read_ptr;
  std::vector<int32_t> snp_idx;
  std::vector<uint8_t> al, bq;
  int64_t numi = 0;

  int32_t nbcs() const { return (int32_t)bcs.size(); }
  int32_t nsnps() const { return (int32_t)snps.size(); }
};

// hexadecimal-string order of the global read counter: the reference keys reads by sprintf("%x", numi++) in a
// std::map<std::string,...> (sc_drop_seq.cpp:365, sc_drop_seq.h:26), so "100" < "ff"
inline uint64_t hex_lex_key(uint32_t c, int* ndigits) {
  int nd = 1;
  for (uint32_t t = c >> 4; t; t >>= 4) ++nd;
  *ndigits = nd;
  return (uint64_t)c << (4 * (16 - nd));
}

struct ReadRec {
  int32_t cell, snp;
  uint64_t key;  // left-aligned hex digits
  uint8_t nd, al, bq;
};

struct PlpFilters {
  int32_t minRead = 0, minUMI = 0, minSNP = 0, minBQ = 1, capBQ = 60;
  std::set<std::string> valid_bcs;  // --group-list
};

// Loads <prefix>.cel.gz and <prefix>.var.gz; `on_snp(idx, chrom, SnpInfo)` is called per .var row in order
// (demuxlet joins the VCF there).  Then <prefix>.plp.gz -> CSR.
template <class OnSnp>
inline void load_plp(const std::string& prefix, const PlpFilters& flt, DropletStore& st, OnSnp on_snp) {
  notice("Loading pileup information with prefix %s", prefix.c_str());
  std::vector<int32_t> index_bcs;
  std::vector<int32_t> tmp_totl, tmp_uniq, tmp_nsnp;
  {
    LineReader lr(prefix + ".cel.gz");
    if (lr.read_fields() != 6 || strcmp(lr.fields[0], "#DROPLET_ID") || strcmp(lr.fields[1], "BARCODE") ||
        strcmp(lr.fields[2], "NUM.READ") || strcmp(lr.fields[3], "NUM.UMI") || strcmp(lr.fields[4], "NUM.UMIwSNP") ||
        strcmp(lr.fields[5], "NUM.SNP"))
      fatal("The header line of %s.cel.gz is malformed or outdated. Expecting #DROPLET_ID BARCODE NUM.READ NUM.UMI NUM.UMIwSNP NUM.SNP", prefix.c_str());
    int32_t nskip = 0;
    std::map<std::string, int32_t> bc_map;
    while (lr.read_fields() > 0) {
      if (lr.fields.size() < 6) fatal("Malformed row in %s.cel.gz", prefix.c_str());
      if (!flt.valid_bcs.empty() && !flt.valid_bcs.count(lr.fields[1])) {
        ++nskip;
        index_bcs.push_back(-1);
        continue;
      }
      const int32_t n_reads = atoi(lr.fields[2]), n_umis = atoi(lr.fields[3]), n_umi_w = atoi(lr.fields[4]),
                    n_snps = atoi(lr.fields[5]);
      if (n_reads < flt.minRead || n_umis < flt.minUMI || n_snps < flt.minSNP) {
        index_bcs.push_back(-1);
        ++nskip;
        continue;
      }
      auto it = bc_map.find(lr.fields[1]);
      int32_t id;
      if (it == bc_map.end()) {
        id = (int32_t)st.bcs.size();
        bc_map[lr.fields[1]] = id;
        st.bcs.push_back(lr.fields[1]);
      } else {
        id = it->second;
      }
      index_bcs.push_back(id);
      if (id + nskip != atoi(lr.fields[0]))
        fatal("Observed DROPLET_ID %d is different from expected DROPLET_ID. Did you modify the digital pileup files by yourself?", atoi(lr.fields[0]));
      tmp_totl.push_back(n_reads);
      tmp_uniq.push_back(n_umi_w);
      tmp_nsnp.push_back(n_snps);
    }
    notice("Finished loading %d droplets, skipping %d", st.nbcs(), nskip);
  }
  {
    LineReader lr(prefix + ".var.gz");
    if (lr.read_fields() != 6 || strcmp(lr.fields[0], "#SNP_ID") || strcmp(lr.fields[1], "CHROM") ||
        strcmp(lr.fields[2], "POS") || strcmp(lr.fields[3], "REF") || strcmp(lr.fields[4], "ALT") || strcmp(lr.fields[5], "AF"))
      fatal("The header line of %s.var.gz is malformed or outdated. Expecting #SNP_ID CHROM POS REF ALT AF", prefix.c_str());
    std::map<std::string, int32_t> chr2rid;
    while (lr.read_fields() > 0) {
      if (lr.fields.size() < 6) fatal("Malformed row in %s.var.gz", prefix.c_str());
      const char* chr = lr.fields[1];
      auto it = chr2rid.find(chr);
      if (it == chr2rid.end()) {
        it = chr2rid.emplace(chr, (int32_t)chr2rid.size()).first;
        st.rid2chr.push_back(chr);
      }
      SnpInfo s;
      s.rid = it->second;
      s.pos = atoi(lr.fields[2]);
      s.ref = lr.fields[3][0];
      s.alt = lr.fields[4][0];
      s.af = atof(lr.fields[5]);
      if ((int64_t)st.snps.size() + 2 != lr.nlines) fatal("Expected SNP nID = %lld but observed %d", (long long)lr.nlines - 2, st.nsnps());
      on_snp(st.nsnps(), chr, s);
      st.snps.push_back(s);
    }
    notice("Finished loading %d variants..", st.nsnps());
  }
  std::vector<ReadRec> recs;
  std::vector<int32_t> pass_totl(st.nbcs(), 0);
  {
    LineReader lr(prefix + ".plp.gz");
    if (lr.read_fields() != 4 || strcmp(lr.fields[0], "#DROPLET_ID") || strcmp(lr.fields[1], "SNP_ID") ||
        strcmp(lr.fields[2], "ALLELES") || strcmp(lr.fields[3], "BASEQS"))
      fatal("THe header line of %s.cel.gz is malformed or outdated. Expecting #DROPLET_ID SNP_ID ALLELES BASEQS", prefix.c_str());
    uint32_t numi = 0;
    while (lr.read_fields() > 0) {
      if (lr.fields.size() < 4) fatal("Malformed row in %s.plp.gz", prefix.c_str());
      const int32_t did = atoi(lr.fields[0]);
      if (did < 0 || did >= (int32_t)index_bcs.size()) fatal("DROPLET_ID %d out of range in %s.plp.gz", did, prefix.c_str());
      const int32_t ibc = index_bcs[did];
      if (ibc < 0) continue;
      const int32_t snp = atoi(lr.fields[1]);
      if (snp < 0 || snp >= st.nsnps()) fatal("SNP_ID %d out of range in %s.plp.gz", snp, prefix.c_str());
      const char* pa = lr.fields[2];
      const char* pq = lr.fields[3];
      const int32_t l = (int32_t)strlen(pq);
      if ((int32_t)strlen(pa) < l) fatal("Length are different between %s and %s", pa, pq);
      for (int32_t i = 0; i < l; ++i) {
        char q = (char)(pq[i] - (char)33);
        if (q >= flt.minBQ) {  // sc_drop_seq.cpp:362-364
          if (q > flt.capBQ) q = (char)flt.capBQ;
          ReadRec r;
          r.cell = ibc;
          r.snp = snp;
          int nd;
          r.key = hex_lex_key(numi++, &nd);
          r.nd = (uint8_t)nd;
          r.al = (uint8_t)(pa[i] - '0');
          r.bq = (uint8_t)q;
          recs.push_back(r);
          ++pass_totl[ibc];
        }
      }
    }
    st.numi = numi;
    notice("Finished loading %u UMIs in total..", numi);
  }
  // reference iteration order: droplet id, SNP id (std::map<int32_t,...>), UMI string (std::map<std::string,...>)
  // (one counting pass by droplet, then every droplet's few hundred reads sorted on their own, droplets spread over
  // the host cores: the same order as one big sort by (droplet, SNP, key, digits), at a fraction of its cost)
  const int32_t nb = st.nbcs();
  {
    std::vector<size_t> start((size_t)nb + 1, 0);
    for (const ReadRec& r : recs) ++start[(size_t)r.cell + 1];
    for (int32_t c = 0; c < nb; ++c) start[(size_t)c + 1] += start[(size_t)c];
    std::vector<ReadRec> by_cell(recs.size());
    {
      std::vector<size_t> fill(start.begin(), start.end() - 1);
      for (const ReadRec& r : recs) by_cell[fill[(size_t)r.cell]++] = r;
    }
    recs.swap(by_cell);
    std::atomic<int32_t> next{0};
    auto work = [&]() {
      for (int32_t c = next++; c < nb; c = next++)
        std::sort(recs.begin() + (ptrdiff_t)start[(size_t)c], recs.begin() + (ptrdiff_t)start[(size_t)c + 1],
                  [](const ReadRec& a, const ReadRec& b) {
                    if (a.snp != b.snp) return a.snp < b.snp;
                    if (a.key != b.key) return a.key < b.key;
                    return a.nd < b.nd;
                  });
    };
    const int nt = (int)std::min<unsigned>(std::max(1u, std::thread::hardware_concurrency()), 16u);
    if (nt <= 1 || recs.size() < 100000) {
      work();
    } else {
      std::vector<std::thread> th;
      for (int t = 0; t < nt; ++t) th.emplace_back(work);
      for (auto& t : th) t.join();
    }
  }
  st.cell_ptr.assign(nb + 1, 0);
  st.uniq_reads.assign(nb, 0);
  st.read_ptr.assign(1, 0);
  st.al.reserve(recs.size());
  st.bq.reserve(recs.size());
  for (size_t i = 0; i < recs.size(); ++i) {
    const ReadRec& r = recs[i];
    if (i == 0 || r.cell != recs[i - 1].cell || r.snp != recs[i - 1].snp) {
      if (i > 0) st.read_ptr.push_back((int64_t)st.al.size());
      st.snp_idx.push_back(r.snp);
      ++st.cell_ptr[r.cell + 1];
    }
    st.al.push_back(r.al);
    st.bq.push_back(r.bq);
    ++st.uniq_reads[r.cell];  // every plp read has its own UMI string (sc_drop_seq.cpp:365)
  }
  if (!recs.empty()) st.read_ptr.push_back((int64_t)st.al.size());
  for (int32_t c = 0; c < nb; ++c) st.cell_ptr[c + 1] += st.cell_ptr[c];
  // cell_totl_reads: counted reads, overwritten by the .cel.gz value when the other counts agree (:375-380)
  st.totl_reads.assign(nb, 0);
  for (int32_t c = 0; c < nb; ++c) {
    st.totl_reads[c] = pass_totl[c];
    if (st.uniq_reads[c] == tmp_uniq[c] && tmp_nsnp[c] == (int32_t)(st.cell_ptr[c + 1] - st.cell_ptr[c])) st.totl_reads[c] = tmp_totl[c];
  }
}

}  // namespace hostio
