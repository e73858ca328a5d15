// main.cpp -- `cramore <command>` dispatcher (cramore.cpp:89-182) for the commands on the GPU path.
#include <cstdio>
#include <cstring>
#include <stdexcept>

#include "host_io.h"

int cmdDemuxlet(int argc, char** argv);
int cmdFreemuxlet(int argc, char** argv);
int cmdFreemux2(int argc, char** argv);
int cmdDscPileup(int argc, char** argv);

// `cramore vcf-text FILE`: the records of a VCF / gzip VCF / bgzipped VCF / BCF as the text lines the readers of this
// build see (a debugging aid and the hook the BCF decoder is tested through; not a reference command)
static int cmdVcfText(int argc, char** argv) {
  if (argc != 2) {
    fprintf(stderr, "usage: cramore vcf-text FILE\n");
    return 1;
  }
  hostio::LineReader lr(argv[1]);
  while (lr.next_line()) {
    fwrite(lr.line.data(), 1, lr.line.size(), stdout);
    fputc('\n', stdout);
  }
  return 0;
}

int main(int argc, char** argv) {
  if (argc < 2 || strcmp(argv[1], "--help") == 0) {
    fprintf(stderr, "cramore (B200 engine build) -- available commands:\n"
                    "  demuxlet    Deconvolute sample identify of droplet-based sc-RNAseq\n"
                    "  freemuxlet  Genotype-free deconvolution of scRNA-seq\n"
                    "  freemux2    Genotype-free deconvolution of scRNA-seq (v2)\n"
                    "  dsc-pileup  Produce pileup of dsc-RNAseq\n");
    return argc < 2 ? 1 : 0;
  }
  try {
    if (strcmp(argv[1], "demuxlet") == 0) return cmdDemuxlet(argc - 1, argv + 1);
    if (strcmp(argv[1], "freemuxlet") == 0) return cmdFreemuxlet(argc - 1, argv + 1);
    if (strcmp(argv[1], "freemux2") == 0) return cmdFreemux2(argc - 1, argv + 1);
    if (strcmp(argv[1], "dsc-pileup") == 0) return cmdDscPileup(argc - 1, argv + 1);
    if (strcmp(argv[1], "vcf-text") == 0) return cmdVcfText(argc - 1, argv + 1);
  } catch (const std::exception&) {
    return 134;  // the reference aborts on error() (uncaught pexception, Error.cpp:29-43)
  }
  fprintf(stderr, "cramore: unknown command %s (this build provides demuxlet, freemuxlet, freemux2 and dsc-pileup)\n", argv[1]);
  return 1;
}
