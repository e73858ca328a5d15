/*
 * ref_phred_shim.cpp -- C accessor around the REFERENCE's own PhredHelper.cpp.
 * Built only into oracle/_ref/libphred_ref.so (git-ignored) from the sources where they lie
 * under /root/reference (see Makefile target `ref`); nothing of the reference is copied here.
 * Used by tests to pin oracle/demux_oracle.cpp's Phred tables and by nothing else.
 */
#include "PhredHelper.h"  // -I/root/reference

extern "C" void ref_phred_tables(double* err256, double* mat256) {
  for (int q = 0; q < 256; ++q) {
    err256[q] = phredConv.phred2Err[q];
    mat256[q] = phredConv.phred2Mat[q];
  }
}
