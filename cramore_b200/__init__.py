"""cramore_b200 -- B200-native (sm_100a) droplet-likelihood engine behind `cramore demuxlet` /
`cramore freemuxlet` (reference: hyunminkang/cramore, cmd_cram_demuxlet.cpp / cmd_cram_freemuxlet.cpp).

The compute path is libcramore_cuda.so (hand-written CUDA, C ABI in include/cramore_cuda.h);
this package is its ctypes binding plus the synthetic-input generator and shard helpers.
"""
from .engine import (ABI_SYMBOLS, CALL_DTYPE, FMX_RESULT_DTYPE, LIB_PATH, PILEUP_DTYPE, RESULT_DTYPE, DemuxEngine,
                     EngineError, FreemuxEngine, best_header, best_row, classify, load_library)

__all__ = ["ABI_SYMBOLS", "CALL_DTYPE", "FMX_RESULT_DTYPE", "LIB_PATH", "PILEUP_DTYPE", "RESULT_DTYPE", "DemuxEngine",
           "EngineError", "FreemuxEngine",
           "best_header", "best_row", "classify", "load_library"]
