/*
 * ref_shim.cpp -- TEST INFRASTRUCTURE: zlib-backed implementations of the stand-ins declared in
 * shim_decls.h, and the `ref_cramore` command dispatcher.  See shim_decls.h for what is real
 * reference code (the engine) and what is restated here (text I/O only).
 */
#include "shim_decls.h"

#include <stdarg.h>
#include <zlib.h>

#include <cmath>

int32_t cmdCramDemuxlet(int32_t argc, char** argv);
int32_t cmdCramFreemuxlet(int32_t argc, char** argv);
int32_t cmdCramFreemux2(int32_t argc, char** argv);

/* ------------------------------------------------------------------ output (hts_utils.cpp:1018-1038) */
struct htsFile {
  FILE* fp;
  gzFile gz;
  FILE* raw;
};

htsFile* hts_open(const char* fn, const char* mode) {
  htsFile* h = new htsFile();
  h->fp = NULL;
  h->gz = NULL;
  h->raw = NULL;
  if (strchr(mode, 'w') && getenv("REF_SHIM_RAW")) h->raw = fopen((std::string(fn) + ".raw").c_str(), "w");
  if (strchr(mode, 'z'))
    h->gz = gzopen(fn, "wb");
  else
    h->fp = fopen(fn, mode);
  if (!h->fp && !h->gz) {
    delete h;
    return NULL;
  }
  return h;
}

int hts_close(htsFile* h) {
  if (!h) return -1;
  if (h->fp) fclose(h->fp);
  if (h->gz) gzclose(h->gz);
  if (h->raw) fclose(h->raw);
  delete h;
  return 0;
}

/* Besides formatting like the reference's hprintf, the shim can record every argument at full
 * precision: with REF_SHIM_RAW=1 in the environment each hprintf call to <file> also appends one
 * line to <file>.raw holding its arguments tab-separated, doubles as %.17g.  The reference prints
 * log-likelihoods with two decimals; the raw side file is what pins the oracle to 1e-12. */
static void raw_dump(FILE* raw, const char* fmt, va_list ap) {
  bool first = true;
  for (const char* p = fmt; *p; ++p) {
    if (*p != '%') continue;
    ++p;
    if (*p == '%') continue;
    while (*p && strchr("-+ #0123456789.", *p)) ++p;
    bool is_long = false;
    while (*p == 'l' || *p == 'h' || *p == 'z') {
      is_long = is_long || (*p == 'l');
      ++p;
    }
    if (!first) fputc('\t', raw);
    first = false;
    switch (*p) {
      case 'd': case 'i': case 'u': case 'x': case 'c':
        if (is_long) fprintf(raw, "%ld", va_arg(ap, long));
        else fprintf(raw, "%d", va_arg(ap, int));
        break;
      case 'f': case 'g': case 'e': case 'G': case 'E':
        fprintf(raw, "%.17g", va_arg(ap, double));
        break;
      case 's': fprintf(raw, "%s", va_arg(ap, const char*)); break;
      default: fprintf(raw, "?"); break;
    }
  }
  fputc('\n', raw);
}

void hprintf(htsFile* h, const char* msg, ...) {
  va_list ap;
  va_start(ap, msg);
  if (h->raw) {
    va_list ap3;
    va_copy(ap3, ap);
    raw_dump(h->raw, msg, ap3);
    va_end(ap3);
  }
  char stackbuf[8192];
  va_list ap2;
  va_copy(ap2, ap);
  int n = vsnprintf(stackbuf, sizeof(stackbuf), msg, ap);
  va_end(ap);
  std::vector<char> big;
  const char* out = stackbuf;
  if (n >= (int)sizeof(stackbuf)) {
    big.resize(n + 1);
    vsnprintf(big.data(), big.size(), msg, ap2);
    out = big.data();
  }
  va_end(ap2);
  if (h->fp) fwrite(out, 1, n, h->fp);
  else gzwrite(h->gz, out, n);
}

/* ------------------------------------------------------------------ tsv_reader (tsv_reader.cpp) */
static bool gz_getline(gzFile gz, std::string& line) {
  line.clear();
  char buf[65536];
  while (gzgets(gz, buf, sizeof(buf)) != NULL) {
    line += buf;
    if (!line.empty() && line[line.size() - 1] == '\n') {
      line.erase(line.size() - 1);
      if (!line.empty() && line[line.size() - 1] == '\r') line.erase(line.size() - 1);
      return true;
    }
  }
  return !line.empty();
}

bool tsv_reader::open(const char* fn) {
  filename = fn;
  gz = gzopen(fn, "rb");
  return gz != NULL;
}
bool tsv_reader::close() {
  if (gz) gzclose((gzFile)gz);
  gz = NULL;
  return true;
}
/* tsv_reader.cpp:27-55: delimiter 0 = split on any white space (ksplit semantics) */
int32_t tsv_reader::read_line() {
  if (!gz_getline((gzFile)gz, line)) {
    nfields = 0;
    return 0;
  }
  ++nlines;
  starts.clear();
  size_t i = 0, n = line.size();
  while (i < n) {
    while (i < n && (delimiter ? false : isspace((unsigned char)line[i]))) ++i;
    if (i >= n) break;
    starts.push_back((int32_t)i);
    while (i < n && (delimiter ? line[i] != (char)delimiter : !isspace((unsigned char)line[i]))) ++i;
    if (i < n) line[i++] = '\0';
  }
  nfields = (int32_t)starts.size();
  return nfields;
}
const char* tsv_reader::str_field_at(int32_t idx) {
  if (idx >= nfields) error("[E:%s] Cannot access field at %d >= %d", __FUNCTION__, idx, nfields);
  return line.c_str() + starts[idx];
}
int32_t tsv_reader::int_field_at(int32_t idx) { return atoi(str_field_at(idx)); }
double tsv_reader::double_field_at(int32_t idx) { return atof(str_field_at(idx)); }

/* ------------------------------------------------------------------ text VCF reader */
const char* bcf_hdr_id2name(const bcf_hdr_t* hdr, int rid) {
  return (rid >= 0 && rid < (int)hdr->chroms.size()) ? hdr->chroms[rid].c_str() : NULL;
}
int bcf_hdr_name2id(const bcf_hdr_t* hdr, const char* name) {
  for (size_t i = 0; i < hdr->chroms.size(); ++i)
    if (hdr->chroms[i] == name) return (int)i;
  return -1;
}
int bcf_get_info_float(const bcf_hdr_t*, bcf1_t* v, const char* tag, float** dst, int* ndst) {
  std::string key = std::string(tag) + "=";
  size_t p = 0;
  while (p < v->info.size()) {
    size_t e = v->info.find(';', p);
    if (e == std::string::npos) e = v->info.size();
    if (v->info.compare(p, key.size(), key) == 0) {
      *dst = (float*)realloc(*dst, sizeof(float));
      (*dst)[0] = strtof(v->info.c_str() + p + key.size(), NULL);
      *ndst = 1;
      return 1;
    }
    p = e + 1;
  }
  return -1;
}

static void split(const std::string& s, char d, std::vector<std::string>& out) {
  out.clear();
  size_t p = 0;
  while (true) {
    size_t e = s.find(d, p);
    if (e == std::string::npos) {
      out.push_back(s.substr(p));
      return;
    }
    out.push_back(s.substr(p, e - p));
    p = e + 1;
  }
}

/* bcf_filtered_reader.cpp:6-165 reduced to: open, header, sample subset (--sm / --sm-list) */
void BCFFilteredReader::init_params() {
  gz = gzopen(bcf_file_name.c_str(), "rb");
  if (!gz) error("[E:%s] Cannot open %s", __PRETTY_FUNCTION__, bcf_file_name.c_str());
  cdr.hdr = new bcf_hdr_t();
  std::string line;
  while (gz_getline((gzFile)gz, line)) {
    if (line.compare(0, 2, "##") == 0) {
      if (line.compare(0, 13, "##contig=<ID=") == 0) {
        size_t e = line.find_first_of(",>", 13);
        cdr.hdr->chroms.push_back(line.substr(13, e - 13));
      }
      continue;
    }
    if (line.compare(0, 6, "#CHROM") == 0) {
      std::vector<std::string> f;
      split(line, '\t', f);
      for (size_t i = 9; i < f.size(); ++i) cdr.hdr->samples.push_back(f[i]);
      break;
    }
    error("[E:%s] Malformed VCF header in %s", __PRETTY_FUNCTION__, bcf_file_name.c_str());
  }
  if (!sample_id_list.empty()) {
    tsv_reader tr(sample_id_list.c_str());
    while (tr.read_line() > 0) sm_ids.insert(tr.str_field_at(0));
  }
  for (int32_t i = 0; i < (int32_t)cdr.hdr->samples.size(); ++i)
    if (sm_ids.empty() || sm_ids.find(cdr.hdr->samples[i]) != sm_ids.end()) sm_icols.push_back(i);
  if (sm_icols.empty()) error("[E:%s] No sample to read from %s", __PRETTY_FUNCTION__, bcf_file_name.c_str());
  vbufs.push_back(new bcf1_t());
  vidx = 0;
}

static bool parse_vcf_line(bcf_hdr_t* hdr, const std::string& line, bcf1_t* v) {
  std::vector<std::string> f;
  split(line, '\t', f);
  if (f.size() < 8) return false;
  int rid = bcf_hdr_name2id(hdr, f[0].c_str());
  if (rid < 0) {
    hdr->chroms.push_back(f[0]);
    rid = (int)hdr->chroms.size() - 1;
  }
  v->rid = rid;
  v->pos = atoi(f[1].c_str()) - 1;
  v->alleles_store.clear();
  v->alleles_store.push_back(f[3]);
  std::vector<std::string> alts;
  split(f[4], ',', alts);
  for (size_t i = 0; i < alts.size(); ++i)
    if (alts[i] != ".") v->alleles_store.push_back(alts[i]);
  v->n_allele = (int32_t)v->alleles_store.size();
  v->allele_ptrs.clear();
  for (size_t i = 0; i < v->alleles_store.size(); ++i) v->allele_ptrs.push_back(&v->alleles_store[i][0]);
  v->d.allele = v->allele_ptrs.data();
  v->qual = (f[5] == ".") ? NAN : strtof(f[5].c_str(), NULL);
  v->info = f[7];
  v->fmt_keys.clear();
  v->sample_cols.clear();
  if (f.size() > 9) {
    split(f[8], ':', v->fmt_keys);
    v->sample_cols.assign(f.begin() + 9, f.end());
  }
  return true;
}

static int fmt_index(const bcf1_t* v, const char* key) {
  for (size_t i = 0; i < v->fmt_keys.size(); ++i)
    if (v->fmt_keys[i] == key) return (int)i;
  return -1;
}
static std::string sample_field(const bcf1_t* v, int sample, int idx) {
  const std::string& col = v->sample_cols[sample];
  size_t p = 0;
  for (int i = 0; i < idx; ++i) {
    p = col.find(':', p);
    if (p == std::string::npos) return ".";
    ++p;
  }
  size_t e = col.find(':', p);
  return col.substr(p, e == std::string::npos ? std::string::npos : e - p);
}

/* bcf_filtered_reader.cpp:198-254 (allele counts over the selected samples, diploid) */
bool BCFFilteredReader::parse_genotypes(bcf_hdr_t* hdr, bcf1_t* v) {
  if (hdr == NULL) hdr = cdr.hdr;
  if (v == NULL) v = cursor();
  const int gi = fmt_index(v, "GT");
  if (gi < 0) return false;
  const int32_t nsamples = (int32_t)hdr->samples.size();
  gts = (int32_t*)realloc(gts, sizeof(int32_t) * 2 * nsamples);
  for (int32_t i = 0; i < nsamples; ++i) {
    std::string g = sample_field(v, i, gi);
    int a[2] = {-1, -1};
    size_t sep = g.find_first_of("/|");
    std::string s0 = g.substr(0, sep), s1 = (sep == std::string::npos) ? "" : g.substr(sep + 1);
    if (!s0.empty() && s0 != ".") a[0] = atoi(s0.c_str());
    if (!s1.empty() && s1 != ".") a[1] = atoi(s1.c_str());
    gts[2 * i] = a[0];
    gts[2 * i + 1] = a[1];
  }
  acs.resize(v->n_allele);
  std::fill(acs.begin(), acs.end(), 0);
  an = 0;
  for (int32_t i = 0; i < (int32_t)sm_icols.size(); ++i)
    for (int32_t j = 0; j < 2; ++j) {
      int32_t al = gts[sm_icols[i] * 2 + j];
      if (al >= 0) {
        ++an;
        ++acs[al];
      }
    }
  return true;
}

/* bcf_filtered_reader.cpp:744-870 (sequential mode) + passed_vfilter :571-647 */
bcf1_t* BCFFilteredReader::read() {
  if (unlimited_buffer && nbuf == (int32_t)vbufs.size()) {
    vbufs.push_back(new bcf1_t());
    vidx = (int32_t)vbufs.size() - 1;
  } else {
    vidx = (vidx + 1) % (int32_t)vbufs.size();
  }
  ++nbuf;
  std::string line;
  while (gz_getline((gzFile)gz, line)) {
    if (line.empty() || line[0] == '#') continue;
    bcf1_t* v = vbufs[vidx];
    if (!parse_vcf_line(cdr.hdr, line, v)) error("[E:%s] Malformed VCF record in %s", __PRETTY_FUNCTION__, bcf_file_name.c_str());
    ++nRead;
    bool pass = v->n_allele <= vfilt.maxAlleles;
    if (pass && (vfilt.minMAC > 0 || vfilt.minCallRate > 0)) {  /* require_GT */
      if (!parse_genotypes(cdr.hdr, v))
        error("[E:%s] Cannot find the field GT from the VCF file at position %s:%d", __PRETTY_FUNCTION__,
              bcf_hdr_id2name(cdr.hdr, v->rid), v->pos + 1);
      if (vfilt.minCallRate > (double)an / (2.0 * (double)sm_icols.size())) pass = false;
      const int32_t ac = an - (int32_t)acs[0];
      if (ac < vfilt.minMAC || an - ac < vfilt.minMAC) pass = false;
    }
    if (pass) return v;
    ++nSkip;
  }
  vidx = (vidx + (int32_t)vbufs.size() - 1) % (int32_t)vbufs.size();
  --nbuf;
  eof = true;
  return NULL;
}

int32_t BCFFilteredReader::clear_buffer_before(const char*, int32_t) { return 0; }

/* bcf_filtered_reader.cpp:406-500, GT and GP branches for biallelic diploid sites */
bool BCFFilteredReader::parse_posteriors(bcf_hdr_t* hdr, bcf1_t* v, const char* name, double gt_error) {
  if (hdr == NULL) hdr = cdr.hdr;
  if (v == NULL) v = cursor();
  const int32_t nsamples = (int32_t)hdr->samples.size();
  const int32_t nalleles = v->n_allele;
  const int32_t ngenos = (nalleles + 1) * nalleles / 2;
  if (strcmp(name, "GT") == 0) {
    if (!parse_genotypes(hdr, v)) return false;
    gps = (float*)realloc(gps, sizeof(float) * ngenos * nsamples);
    n_gps = nsamples * 3;
    for (int32_t i = 0; i < (int32_t)sm_icols.size(); ++i) {
      const int32_t a1 = gts[sm_icols[i] * 2], a2 = gts[sm_icols[i] * 2 + 1];
      const int32_t g = (a1 < 0 || a2 < 0) ? -1 : (a1 > a2 ? a1 * (a1 + 1) / 2 + a2 : a2 * (a2 + 1) / 2 + a1);
      const int32_t icol = sm_icols[i] * ngenos;
      if (g < 0) {
        for (int32_t j = 0, l = 0; j < nalleles; ++j)
          for (int32_t k = 0; k <= j; ++k, ++l)
            gps[icol + l] = (float)((j == k ? 1.0 : 2.0) * (acs[j] + 1.0 / nalleles) / (an + 1.0) * (acs[k] + 1.0 / nalleles) / (an + 1.0));
      } else {
        for (int32_t j = 0; j < ngenos; ++j) gps[icol + j] = (g == j) ? 1.0 - gt_error : gt_error / (ngenos - 1.0);
      }
    }
    return true;
  }
  if (strcmp(name, "PL") == 0) error("[E:%s] the I/O shim does not read PL", __PRETTY_FUNCTION__);
  const int fi = fmt_index(v, name);
  if (fi < 0) return false;
  gps = (float*)realloc(gps, sizeof(float) * ngenos * nsamples);
  n_gps = ngenos * nsamples;
  for (int32_t i = 0; i < nsamples; ++i) {
    std::string g = sample_field(v, i, fi);
    const char* p = g.c_str();
    for (int32_t j = 0; j < ngenos; ++j) {
      char* e;
      gps[i * ngenos + j] = strtof(p, &e);
      p = (*e == ',') ? e + 1 : e;
    }
  }
  std::vector<float> gpSums(ngenos, 0);
  for (int32_t i = 0; i < nalleles; ++i)
    for (int32_t j = 0; j <= i; ++j) gpSums[(i + 1) * i / 2 + j] = ((i == j) ? 1.0 : 2.0) / (float)(nalleles * nalleles);
  for (int32_t i = 0; i < (int32_t)sm_icols.size(); ++i) {  /* :474-485 */
    const int32_t icol = sm_icols[i] * ngenos;
    float sumgp = 0;
    for (int32_t j = 0; j < ngenos; ++j) sumgp += gps[icol + j];
    for (int32_t j = 0; j < ngenos; ++j) {
      gps[icol + j] /= sumgp;
      gpSums[j] += gps[icol + j];
    }
  }
  for (int32_t j = 0; j < ngenos; ++j) gpSums[j] /= (int32_t)(sm_icols.size() + 1.0);
  for (int32_t i = 0; i < (int32_t)sm_icols.size(); ++i) {  /* :490-496 */
    const int32_t icol = sm_icols[i] * ngenos;
    for (int32_t j = 0; j < ngenos; ++j) gps[icol + j] = ((1.0 - gt_error) * gps[icol + j] + gt_error * gpSums[j]);
  }
  return true;
}

/* ------------------------------------------------------------------ BAM side: not available */
static void no_bam() { error("the reference I/O shim has no BAM/CRAM reader: use --plp"); }
int32_t bam_hdr_get_n_targets(bam_hdr_t*) { no_bam(); return 0; }
const char* bam_get_chromi(bam_hdr_t*, int32_t) { no_bam(); return NULL; }
const char* bam_get_chrom(bam_hdr_t*, bam1_t*) { no_bam(); return NULL; }
int32_t bam_endpos(const bam1_t*) { no_bam(); return 0; }
uint8_t* bam_aux_get(const bam1_t*, const char*) { no_bam(); return NULL; }
char* bam_aux2Z(const uint8_t*) { no_bam(); return NULL; }
const char* bam_get_qname(const bam1_t*) { no_bam(); return NULL; }
void bam_get_base_and_qual_and_read_and_qual(bam1_t*, uint32_t, char&, char&, int32_t&, kstring_t*, kstring_t*) { no_bam(); }
void SAMFilteredReader::set_buffer_size(int32_t) {}
void SAMFilteredReader::init_params() { no_bam(); }
bam1_t* SAMFilteredReader::read() { no_bam(); return NULL; }
bam1_t* SAMFilteredReader::cursor() { no_bam(); return NULL; }
void SAMFilteredReader::close() {}

/* ------------------------------------------------------------------ command dispatch (cramore.cpp:89-182) */
int main(int argc, char** argv) {
  if (argc < 2) {
    fprintf(stderr, "usage: ref_cramore {demuxlet|freemuxlet|freemux2} [options]\n");
    return 1;
  }
  if (strcmp(argv[1], "demuxlet") == 0) return cmdCramDemuxlet(argc - 1, argv + 1);
  if (strcmp(argv[1], "freemuxlet") == 0) return cmdCramFreemuxlet(argc - 1, argv + 1);
  if (strcmp(argv[1], "freemux2") == 0) return cmdCramFreemux2(argc - 1, argv + 1);
  fprintf(stderr, "ref_cramore: only demuxlet, freemuxlet and freemux2 are built\n");
  return 1;
}
