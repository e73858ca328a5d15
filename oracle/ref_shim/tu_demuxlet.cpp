// The reference's demuxlet driver, compiled from where it lies against the I/O shim.
#include "shim_decls.h"
#include "cmd_cram_demuxlet.cpp"
