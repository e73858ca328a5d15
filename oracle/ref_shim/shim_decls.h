/*
 * shim_decls.h -- TEST INFRASTRUCTURE.  Lets the reference's own likelihood code
 * (/root/reference/cmd_cram_demuxlet.cpp, cmd_cram_freemuxlet.cpp, sc_drop_seq.cpp, params.cpp,
 * Error.cpp, PhredHelper.cpp) compile WHERE IT LIES, unmodified, without htslib.
 *
 * htslib is only on the I/O side of the path.  This header pre-defines the include
 * guards of the reference's four htslib-facing headers (hts_utils.h, bcf_filtered_reader.h,
 * sam_filtered_reader.h, tsv_reader.h) and declares stand-ins with the same class and function
 * names, exposing exactly the members the three command/library files touch.  The stand-ins are
 * implemented in ref_shim.cpp on top of zlib: a tab/space line reader for the digital pileup
 * files, a text-VCF reader restating BCFFilteredReader::read/passed_vfilter/parse_posteriors
 * (bcf_filtered_reader.cpp:406-500, 571-647) for GT / GP fields, and hprintf writers.  The BAM
 * side (`--sam`) is declared so the file compiles, and aborts if reached.
 *
 * Everything between "ingestion finished" and "row printed" -- the engine this repository
 * replaces -- is therefore the reference's own compiled code.
 */
#ifndef REF_SHIM_DECLS_H
#define REF_SHIM_DECLS_H

#define HTS_UTILS_H
#define __BCF_FILTERED_READER_H
#define __SAM_FILTERED_READER_H
#define __TSV_READER_H

#include <stdint.h>

#include <climits>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <set>
#include <string>
#include <vector>

#include "Error.h"

/* ---- the handful of htslib types the engine files name ------------------------------------ */
struct kstring_t {
  size_t l, m;
  char* s;
};
struct bcf_hdr_t {
  std::vector<std::string> chroms;   /* contig names in order of first appearance */
  std::vector<std::string> samples;  /* all sample ids of the #CHROM line */
};
struct bcf_dec_t {
  char** allele;
};
struct bcf1_t {
  int32_t rid, pos; /* pos is 0-based like htslib */
  int32_t n_allele;
  float qual;
  bcf_dec_t d;
  std::vector<std::string> alleles_store;
  std::vector<char*> allele_ptrs;
  std::string info;
  std::vector<std::string> fmt_keys;
  std::vector<std::string> sample_cols;
};
struct bam_hdr_t {};
struct bam1_core_t {
  int32_t tid, pos, l_qseq;
};
#define BAM_READ_INDEX_NA -1
struct bam1_t {
  bam1_core_t core;
};
struct htsFile;

htsFile* hts_open(const char* fn, const char* mode);
int hts_close(htsFile* fp);
void hprintf(htsFile* fp, const char* msg, ...);

const char* bcf_hdr_id2name(const bcf_hdr_t* hdr, int rid);
int bcf_hdr_name2id(const bcf_hdr_t* hdr, const char* name);
int bcf_get_info_float(const bcf_hdr_t* hdr, bcf1_t* v, const char* tag, float** dst, int* ndst);

/* BAM side: declared only (the --sam branch of cmdCramDemuxlet must compile) */
int32_t bam_hdr_get_n_targets(bam_hdr_t* h);
const char* bam_get_chromi(bam_hdr_t* h, int32_t i);
const char* bam_get_chrom(bam_hdr_t* h, bam1_t* b);
int32_t bam_endpos(const bam1_t* b);
uint8_t* bam_aux_get(const bam1_t* b, const char tag[2]);
char* bam_aux2Z(const uint8_t* s);
const char* bam_get_qname(const bam1_t* b);
#define bam_get_pos1(b) ((b)->core.pos + 1)
void bam_get_base_and_qual_and_read_and_qual(bam1_t* b, uint32_t pos, char& base, char& qual, int32_t& rpos,
                                             kstring_t* readseq, kstring_t* readqual);

/* ---- tsv_reader (tsv_reader.h) -------------------------------------------------------------- */
class tsv_reader {
 public:
  std::string filename;
  void* gz;
  std::string line;
  std::vector<int32_t> starts;
  int32_t nfields;
  int32_t nlines;
  int32_t delimiter;
  bool open(const char* filename);
  bool close();
  int32_t read_line();
  const char* str_field_at(int32_t idx);
  int32_t int_field_at(int32_t idx);
  double double_field_at(int32_t idx);
  tsv_reader() : gz(NULL), nfields(0), nlines(0), delimiter(0) {}
  tsv_reader(const char* fn) : gz(NULL), nfields(0), nlines(0), delimiter(0) {
    if (!open(fn)) error("[E:%s] Cannot open file %s for reading", __FUNCTION__, fn);
  }
  ~tsv_reader() { close(); }
};

/* ---- BCFFilteredReader (bcf_filtered_reader.h), members used by the engine files ---------- */
struct bcf_vfilter_arg {
  int32_t minMAC, maxAlleles;
  double minCallRate;
  bcf_vfilter_arg() : minMAC(0), maxAlleles(INT_MAX), minCallRate(0) {}
};
struct BCFChunkedReader {
  bcf_hdr_t* hdr;
  BCFChunkedReader() : hdr(NULL) {}
};

class BCFFilteredReader {
 public:
  std::string bcf_file_name;
  std::string sample_id_list;
  bcf_vfilter_arg vfilt;
  int32_t verbose;
  bool unlimited_buffer;
  int32_t nbuf;
  bool eof;
  BCFChunkedReader cdr;
  float* gps;
  int32_t n_gps;
  int32_t* gts; /* 2 per sample, -1 = missing */
  std::vector<double> acs;
  int32_t an;
  std::set<std::string> sm_ids;
  std::vector<int32_t> sm_icols;
  int32_t nRead, nSkip;
  std::vector<bcf1_t*> vbufs;
  int32_t vidx;
  void* gz;

  BCFFilteredReader()
      : verbose(10000), unlimited_buffer(false), nbuf(0), eof(false), gps(NULL), n_gps(0), gts(NULL), an(0), nRead(0),
        nSkip(0), vidx(-1), gz(NULL) {}
  void init_params();
  inline bcf1_t* cursor() { return vbufs[vidx]; }
  bcf1_t* read();
  int32_t clear_buffer_before(const char* chr = NULL, int32_t pos1 = INT_MAX);
  bool parse_genotypes(bcf_hdr_t* hdr = NULL, bcf1_t* v = NULL);
  bool parse_posteriors(bcf_hdr_t* hdr = NULL, bcf1_t* v = NULL, const char* name = "GP", double gt_error_offset = 1e-4);
  inline float get_posterior_at(int32_t i, int32_t ngenotypes = 3) {
    return gps[sm_icols[i / ngenotypes] * ngenotypes + (i % ngenotypes)];
  }
  inline const char* get_sample_id_at(int32_t i) { return cdr.hdr->samples[sm_icols[i]].c_str(); }
  inline int32_t get_nsamples() { return (int32_t)sm_icols.size(); }
  inline bool add_specified_sample(const char* id) { return sm_ids.insert(id).second; }
  inline double get_af(int32_t allele, double pseudocount = 1.0) {
    return (acs[allele] + pseudocount / acs.size()) / (an + pseudocount);
  }
};

/* ---- SAMFilteredReader (sam_filtered_reader.h): --sam is not available through the shim ---- */
struct sam_filter_arg {
  int32_t exclude_flag, minMQ;
};
class SAMFilteredReader {
 public:
  std::string sam_file_name;
  sam_filter_arg filt;
  int32_t verbose;
  int32_t n_read, n_skip;
  bam_hdr_t* hdr;
  SAMFilteredReader() : verbose(1000000), n_read(0), n_skip(0), hdr(NULL) { filt.exclude_flag = 0; filt.minMQ = 0; }
  void set_buffer_size(int32_t n);
  void init_params();
  bam1_t* read();
  bam1_t* cursor();
  void close();
};

#endif
