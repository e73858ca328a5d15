"""ctypes loader for the CPU oracle (TEST INFRASTRUCTURE -- see oracle/oracle.h).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
may import this module.  The product package (cramore_b200) never does.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None
_REF = None


class DemuxResult(C.Structure):
    _fields_ = [
        ("sBest", C.c_int32), ("sNext", C.c_int32),
        ("dBest1", C.c_int32), ("dBest2", C.c_int32), ("dBestA", C.c_int32),
        ("dNext1", C.c_int32), ("dNext2", C.c_int32), ("dNextA", C.c_int32),
        ("sngBestLLK", C.c_double), ("sngNextLLK", C.c_double),
        ("dblBestLLK", C.c_double), ("dblNextLLK", C.c_double),
        ("sumLLK", C.c_double), ("sngLLK", C.c_double),
    ]


DEMUX_RESULT_DTYPE = np.dtype([
    ("sBest", "<i4"), ("sNext", "<i4"), ("dBest1", "<i4"), ("dBest2", "<i4"), ("dBestA", "<i4"),
    ("dNext1", "<i4"), ("dNext2", "<i4"), ("dNextA", "<i4"),
    ("sngBestLLK", "<f8"), ("sngNextLLK", "<f8"), ("dblBestLLK", "<f8"), ("dblNextLLK", "<f8"),
    ("sumLLK", "<f8"), ("sngLLK", "<f8")])

DEMUX_CALL_DTYPE = np.dtype([
    ("type", "<i4"), ("jBest", "<i4"), ("kBest", "<i4"), ("alphaBest", "<i4"),
    ("jNext", "<i4"), ("kNext", "<i4"), ("alphaNext", "<i4"), ("_pad", "<i4"),
    ("bestLLK", "<f8"), ("nextLLK", "<f8"), ("bestPP", "<f8"), ("sngPP", "<f8"), ("sngOnlyPP", "<f8")])

PILEUP_DTYPE = np.dtype([
    ("nreads", "<i4"), ("nref", "<i4"), ("nalt", "<i4"), ("_pad", "<i4"),
    ("gls", "<f8", (9,)), ("logdenom", "<f8")])

FMX_RESULT_DTYPE = np.dtype([
    ("sBest", "<i4"), ("sNext", "<i4"), ("dBest1", "<i4"), ("dBest2", "<i4"), ("dNext1", "<i4"),
    ("dNext2", "<i4"), ("type", "<i4"), ("jBest", "<i4"), ("kBest", "<i4"), ("jNext", "<i4"),
    ("kNext", "<i4"), ("_pad", "<i4"),
    ("sngBestLLK", "<f8"), ("sngNextLLK", "<f8"), ("dblBestLLK", "<f8"), ("dblNextLLK", "<f8"),
    ("sumLLK", "<f8"), ("sngLLK", "<f8"), ("bestLLK", "<f8"), ("nextLLK", "<f8"),
    ("bestPP", "<f8"), ("sngPP", "<f8"), ("sngOnlyPP", "<f8")])


def build(force=False):
    """Compile oracle/liboracle.so (and oracle/_ref when /root/reference is present)."""
    so = os.path.join(_HERE, "liboracle.so")
    srcs = [os.path.join(_HERE, f) for f in ("demux_oracle.cpp", "fmx_oracle.cpp", "oracle.h", "Makefile")]
    stale = force or not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs)
    if stale:
        subprocess.check_call(["make", "-C", _HERE, "liboracle.so"], stdout=subprocess.DEVNULL)
    if os.path.isdir("/root/reference") and (force or not os.path.exists(os.path.join(_HERE, "_ref", "libphred_ref.so"))):
        subprocess.check_call(["make", "-C", _HERE, "ref"], stdout=subprocess.DEVNULL)
    return so


def _p(a, t):
    return a.ctypes.data_as(C.POINTER(t)) if a is not None else None


def lib():
    global _LIB
    if _LIB is None:
        _LIB = C.CDLL(build())
        _LIB.orc_log_add.restype = C.c_double
        _LIB.orc_log_add.argtypes = [C.c_double, C.c_double]
        _LIB.orc_demux_score.restype = C.c_int
        _LIB.orc_demux_best_row.restype = C.c_int
    return _LIB


def ref_lib():
    """The reference's own PhredHelper.cpp, compiled into oracle/_ref (None if not built)."""
    global _REF
    if _REF is None:
        p = os.path.join(_HERE, "_ref", "libphred_ref.so")
        if not os.path.exists(p):
            return None
        _REF = C.CDLL(p)
    return _REF


def phred_tables():
    err = np.empty(256, np.float64)
    mat = np.empty(256, np.float64)
    lib().orc_phred_tables(_p(err, C.c_double), _p(mat, C.c_double))
    return err, mat


def ref_phred_tables():
    r = ref_lib()
    if r is None:
        return None
    err = np.empty(256, np.float64)
    mat = np.empty(256, np.float64)
    r.ref_phred_tables(_p(err, C.c_double), _p(mat, C.c_double))
    return err, mat


def log_add(a, b):
    return lib().orc_log_add(a, b)


def mix_genotypes(gp_f32, snp_err, has_gp=None):
    gp = np.ascontiguousarray(gp_f32, np.float32)
    nsnps, nv, _ = gp.shape
    err = np.ascontiguousarray(np.broadcast_to(np.asarray(snp_err, np.float64), (nsnps,)))
    hg = None if has_gp is None else np.ascontiguousarray(has_gp, np.uint8)
    out = np.zeros((nsnps, nv, 3), np.float64)
    lib().orc_mix_genotypes(C.c_int64(nsnps), C.c_int32(nv), _p(gp, C.c_float), _p(err, C.c_double),
                            _p(hg, C.c_uint8), _p(out, C.c_double))
    return out


def demux_tile(al, bq, alpha):
    al = np.ascontiguousarray(al, np.uint8)
    bq = np.ascontiguousarray(bq, np.uint8)
    alpha = np.ascontiguousarray(alpha, np.float64)
    out = np.empty(len(alpha) * 9, np.float64)
    lib().orc_demux_tile(C.c_int64(len(al)), _p(al, C.c_uint8), _p(bq, C.c_uint8), C.c_int32(len(alpha)),
                         _p(alpha, C.c_double), _p(out, C.c_double))
    return out


def demux_score(cell_ptr, snp_idx, read_ptr, al, bq, gps, alpha, doublet_prior=0.5, has_gp=None,
                cell_begin=0, cell_end=None, want_llk=False, want_tiles=False):
    """Returns (results[cells] structured array, llk [cells,nv,nv,nA] or None, tiles or None)."""
    cell_ptr = np.ascontiguousarray(cell_ptr, np.int64)
    snp_idx = np.ascontiguousarray(snp_idx, np.int32)
    read_ptr = np.ascontiguousarray(read_ptr, np.int64)
    al = np.ascontiguousarray(al, np.uint8)
    bq = np.ascontiguousarray(bq, np.uint8)
    gps = np.ascontiguousarray(gps, np.float64)
    alpha = np.ascontiguousarray(alpha, np.float64)
    nsnps, nv, _ = gps.shape
    nbcs = len(cell_ptr) - 1
    nA = len(alpha)
    if cell_end is None:
        cell_end = nbcs
    n = cell_end - cell_begin
    out = np.zeros(n, DEMUX_RESULT_DTYPE)
    llk = np.zeros((n, nv, nv, nA), np.float64) if want_llk else None
    nent = int(cell_ptr[cell_end] - cell_ptr[cell_begin])
    tiles = np.zeros((nent, nA * 9), np.float64) if want_tiles else None
    hg = None if has_gp is None else np.ascontiguousarray(has_gp, np.uint8)
    rc = lib().orc_demux_score(
        C.c_int64(nbcs), _p(cell_ptr, C.c_int64), _p(snp_idx, C.c_int32), _p(read_ptr, C.c_int64),
        _p(al, C.c_uint8), _p(bq, C.c_uint8), C.c_int64(nsnps), C.c_int32(nv), _p(gps, C.c_double),
        _p(hg, C.c_uint8), C.c_int32(nA), _p(alpha, C.c_double), C.c_double(doublet_prior),
        C.c_int64(cell_begin), C.c_int64(cell_end), out.ctypes.data_as(C.c_void_p),
        _p(llk, C.c_double), _p(tiles, C.c_double))
    if rc != 0:
        raise ValueError("orc_demux_score: bad arguments (nv>=2, nA>=2, alpha[0]==0 required)")
    return out, llk, tiles


def demux_classify(results, nv, alpha, doublet_prior=0.5):
    alpha = np.ascontiguousarray(alpha, np.float64)
    res = np.ascontiguousarray(results)
    calls = np.zeros(len(res), DEMUX_CALL_DTYPE)
    L = lib()
    for i in range(len(res)):
        L.orc_demux_classify(C.c_void_p(res[i:i + 1].ctypes.data), C.c_int32(nv), C.c_int32(len(alpha)),
                             _p(alpha, C.c_double), C.c_double(doublet_prior),
                             C.c_void_p(calls[i:i + 1].ctypes.data))
    return calls


def demux_best_row(int_id, barcode, nsnps_cell, nreads_cell, result, nv, alpha, doublet_prior, sample_ids):
    alpha = np.ascontiguousarray(alpha, np.float64)
    buf = C.create_string_buffer(4096)
    ids = (C.c_char_p * nv)(*[s.encode() for s in sample_ids])
    res = np.ascontiguousarray(result).reshape(1)
    n = lib().orc_demux_best_row(buf, C.c_int(4096), C.c_int32(int_id), barcode.encode(),
                                 C.c_uint32(nsnps_cell), C.c_int32(nreads_cell),
                                 C.c_void_p(res.ctypes.data), C.c_int32(nv), C.c_int32(len(alpha)),
                                 _p(alpha, C.c_double), C.c_double(doublet_prior), ids)
    return buf.raw[:n].decode()


# ------------------------------------------------------------------ freemuxlet

def fmx_pileup(read_ptr, al, bq, alpha=0.5):
    read_ptr = np.ascontiguousarray(read_ptr, np.int64)
    al = np.ascontiguousarray(al, np.uint8)
    bq = np.ascontiguousarray(bq, np.uint8)
    nnz = len(read_ptr) - 1
    out = np.zeros(nnz, PILEUP_DTYPE)
    lib().orc_fmx_pileup(C.c_int64(nnz), _p(read_ptr, C.c_int64), _p(al, C.c_uint8), _p(bq, C.c_uint8),
                         C.c_double(alpha), C.c_void_p(out.ctypes.data))
    return out


def fmx_lmix(cell_ptr, snp_idx, plp, af):
    cell_ptr = np.ascontiguousarray(cell_ptr, np.int64)
    snp_idx = np.ascontiguousarray(snp_idx, np.int32)
    af = np.ascontiguousarray(af, np.float64)
    nbcs = len(cell_ptr) - 1
    llk0 = np.zeros(nbcs)
    llk2 = np.zeros(nbcs)
    lib().orc_fmx_lmix(C.c_int64(nbcs), _p(cell_ptr, C.c_int64), _p(snp_idx, C.c_int32),
                       C.c_void_p(plp.ctypes.data), _p(af, C.c_double), _p(llk0, C.c_double),
                       _p(llk2, C.c_double))
    return llk0, llk2


def fmx_mstep(cell_ptr, snp_idx, plp, assign, K, nsnps):
    cell_ptr = np.ascontiguousarray(cell_ptr, np.int64)
    snp_idx = np.ascontiguousarray(snp_idx, np.int32)
    assign = np.ascontiguousarray(assign, np.int32)
    clust = np.zeros((K, nsnps), PILEUP_DTYPE)
    lib().orc_fmx_mstep(C.c_int64(len(cell_ptr) - 1), _p(cell_ptr, C.c_int64), _p(snp_idx, C.c_int32),
                        C.c_void_p(plp.ctypes.data), _p(assign, C.c_int32), C.c_int32(K),
                        C.c_int64(nsnps), C.c_void_p(clust.ctypes.data))
    return clust


def fmx_estep(cell_ptr, snp_idx, plp, af, clust, doublet_prior=0.5, geno_error=0.0, apply_geno_error=False,
              want_llk=False):
    cell_ptr = np.ascontiguousarray(cell_ptr, np.int64)
    snp_idx = np.ascontiguousarray(snp_idx, np.int32)
    af = np.ascontiguousarray(af, np.float64)
    clust = np.ascontiguousarray(clust)
    K, nsnps = clust.shape
    nbcs = len(cell_ptr) - 1
    out = np.zeros(nbcs, FMX_RESULT_DTYPE)
    llk = np.zeros((nbcs, K * (K + 1) // 2)) if want_llk else None
    lib().orc_fmx_estep(C.c_int64(nbcs), _p(cell_ptr, C.c_int64), _p(snp_idx, C.c_int32),
                        C.c_void_p(plp.ctypes.data), _p(af, C.c_double), C.c_int32(K), C.c_int64(nsnps),
                        C.c_void_p(clust.ctypes.data), C.c_double(doublet_prior), C.c_double(geno_error),
                        C.c_int32(1 if apply_geno_error else 0), C.c_void_p(out.ctypes.data),
                        _p(llk, C.c_double))
    return out, llk


DROPD_DTYPE = np.dtype([("nsnps", "<i4"), ("nread1", "<i4"), ("nread2", "<i4"), ("_pad", "<i4"),
                        ("llk0", "<f8"), ("llk2", "<f8")])


def fmx_pair_dist(cell_ptr, snp_idx, plp, af, nsnps):
    """Dense dropDs matrix [nbcs][nbcs] (rows = the larger droplet id), cmd_cram_freemuxlet.cpp:172-221."""
    cell_ptr = np.ascontiguousarray(cell_ptr, np.int64)
    snp_idx = np.ascontiguousarray(snp_idx, np.int32)
    af = np.ascontiguousarray(af, np.float64)
    nbcs = len(cell_ptr) - 1
    out = np.zeros((nbcs, nbcs), DROPD_DTYPE)
    lib().orc_fmx_pair_dist(C.c_int64(nbcs), C.c_int64(nsnps), _p(cell_ptr, C.c_int64), _p(snp_idx, C.c_int32),
                            C.c_void_p(plp.ctypes.data), _p(af, C.c_double), C.c_void_p(out.ctypes.data))
    return out
