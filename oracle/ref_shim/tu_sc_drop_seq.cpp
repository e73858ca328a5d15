// The reference's droplet store + GL helpers, compiled from where they lie against the I/O shim.
#include "shim_decls.h"
#include "sc_drop_seq.cpp"
