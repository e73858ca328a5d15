// The reference's freemux2 driver, compiled from where it lies against the I/O shim.
#include "shim_decls.h"
#include "cmd_cram_freemux2.cpp"
