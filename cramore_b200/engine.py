"""ctypes binding of libcramore_cuda.so -- the C ABI in include/cramore_cuda.h.

This is the only compute path of the package: there is no Python or CPU implementation behind
it.  If the shared library is missing or no sm_100 GPU is usable, construction raises.
"""
import ctypes as C
import os

import numpy as np

_PKG = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("CCU_LIB") or os.path.join(_PKG, "libcramore_cuda.so")  # CCU_LIB: A/B builds of the same ABI
_lib = None

RESULT_DTYPE = np.dtype([
    ("sBest", "<i4"), ("sNext", "<i4"), ("dBest1", "<i4"), ("dBest2", "<i4"), ("dBestA", "<i4"),
    ("dNext1", "<i4"), ("dNext2", "<i4"), ("dNextA", "<i4"),
    ("sngBestLLK", "<f8"), ("sngNextLLK", "<f8"), ("dblBestLLK", "<f8"), ("dblNextLLK", "<f8"),
    ("sumLLK", "<f8"), ("sngLLK", "<f8")])

CALL_DTYPE = np.dtype([
    ("type", "<i4"), ("jBest", "<i4"), ("kBest", "<i4"), ("alphaBest", "<i4"),
    ("jNext", "<i4"), ("kNext", "<i4"), ("alphaNext", "<i4"), ("_pad", "<i4"),
    ("bestLLK", "<f8"), ("nextLLK", "<f8"), ("bestPP", "<f8"), ("sngPP", "<f8"), ("sngOnlyPP", "<f8")])


class Timings(C.Structure):
    _fields_ = [("ms_genotypes", C.c_float), ("ms_tile", C.c_float), ("ms_singlet", C.c_float),
                ("ms_score", C.c_float), ("ms_finalize", C.c_float), ("ms_run", C.c_float),
                ("ms_h2d", C.c_float), ("ms_d2h", C.c_float), ("launches", C.c_int32),
                ("score_ctas", C.c_int32), ("score_items", C.c_int64)]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_}


# every symbol include/cramore_cuda.h declares (tests check the library exports all of them)
ABI_SYMBOLS = [
    "ccu_create", "ccu_destroy", "ccu_last_error", "ccu_set_stream", "ccu_synchronize", "ccu_get_timings",
    "ccu_demux_configure", "ccu_demux_set_genotypes", "ccu_demux_score", "ccu_demux_upload",
    "ccu_demux_run", "ccu_demux_wait", "ccu_demux_download", "ccu_demux_classify", "ccu_demux_best_row",
    "ccu_demux_best_header", "ccu_demux_get_tiles", "ccu_demux_get_genotypes", "ccu_demux_get_llk",
    "ccu_measure_fp64_peak",
    "ccu_fmx_configure", "ccu_fmx_upload", "ccu_fmx_set_shard", "ccu_fmx_lmix", "ccu_fmx_get_pileups",
    "ccu_fmx_mstep", "ccu_fmx_estep", "ccu_fmx_get_llk", "ccu_fmx_get_clusters", "ccu_fmx_stats_dev",
    "ccu_fmx_assign_dev", "ccu_fmx_mstep_dev", "ccu_fmx_estep_dev", "ccu_fmx_download_results", "ccu_fmx_run_em",
    "ccu_fmx_get_timings", "ccu_fmx_selftest_division", "ccu_fmx_post_dev", "ccu_fmx_posteriors_dev",
    "ccu_fmx_estep_post_dev", "ccu_fmx_pair_distances", "ccu_fmx_get_pair_distances", "ccu_fmx_initial_clusters",
    "ccu_fmx_greedy_clusters", "ccu_fmx_get_voting_pairs", "ccu_fmx_peer_handles", "ccu_fmx_peer_open",
    "ccu_fmx_peer_close", "ccu_fmx_mstep_fused_dev", "ccu_fmx_estep_fused_dev", "ccu_fmx_rank_barrier_dev",
    "ccu_fmx_rank_barrier_failed", "ccu_fmx_run_em_multi", "ccu_demux_set_genotypes_mixed",
    "ccu_demux_genotypes_stage", "ccu_demux_genotypes_commit", "ccu_demux_results_dev",
    "ccu_demux_refine", "ccu_demux_refine_mixed", "ccu_demux_copy_genotypes", "ccu_demux_get_singlets",
]


def load_library():
    """dlopen libcramore_cuda.so (no GPU needed for this step)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                "%s not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(make -C cramore_b200/csrc). There is no CPU fallback." % LIB_PATH)
        lib = C.CDLL(LIB_PATH)
        lib.ccu_last_error.restype = C.c_char_p
        lib.ccu_last_error.argtypes = [C.c_void_p]
        lib.ccu_demux_best_header.restype = C.c_char_p
        for name in ABI_SYMBOLS:
            fn = getattr(lib, name)
            if name in ("ccu_fmx_stats_dev", "ccu_fmx_assign_dev", "ccu_fmx_post_dev", "ccu_demux_results_dev"):
                fn.restype = C.c_void_p
            elif name not in ("ccu_last_error", "ccu_demux_best_header"):
                fn.restype = C.c_int
        _lib = lib
    return _lib


def _ptr(a):
    return C.c_void_p(a.ctypes.data) if a is not None else C.c_void_p(0)


class EngineError(RuntimeError):
    pass


class DemuxEngine:
    """One GPU's demuxlet engine (mirrors the call sequence a patched cmdCramDemuxlet makes)."""

    def __init__(self, device=0):
        self.lib = load_library()
        self.h = C.c_void_p(0)
        rc = self.lib.ccu_create(C.c_int32(device), C.byref(self.h))
        if rc != 0:
            raise EngineError(self.lib.ccu_last_error(None).decode())
        self.nv = self.nA = 0
        self.nbcs = 0
        self._keep = []

    def close(self):
        if self.h:
            self.lib.ccu_destroy(self.h)
            self.h = C.c_void_p(0)

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc):
        if rc != 0:
            raise EngineError(self.lib.ccu_last_error(self.h).decode())

    def set_stream(self, cuda_stream_ptr):
        self._check(self.lib.ccu_set_stream(self.h, C.c_void_p(cuda_stream_ptr)))

    def synchronize(self):
        self._check(self.lib.ccu_synchronize(self.h))

    def configure(self, nv, alphas, doublet_prior=0.5):
        a = np.ascontiguousarray(alphas, np.float64)
        self._check(self.lib.ccu_demux_configure(self.h, C.c_int32(nv), C.c_int32(len(a)), _ptr(a),
                                                 C.c_double(doublet_prior)))
        self.nv, self.nA, self.alphas, self.doublet_prior = nv, len(a), a, doublet_prior

    def set_genotypes(self, gp_f32, snp_err, has_gp=None):
        gp = np.ascontiguousarray(gp_f32, np.float32)
        nsnps = gp.shape[0]
        assert gp.shape == (nsnps, self.nv, 3)
        err = np.ascontiguousarray(np.broadcast_to(np.asarray(snp_err, np.float64), (nsnps,)))
        hg = None if has_gp is None else np.ascontiguousarray(has_gp, np.uint8)
        self._check(self.lib.ccu_demux_set_genotypes(self.h, C.c_int64(nsnps), _ptr(gp), _ptr(err), _ptr(hg)))
        self.nsnps = nsnps

    def genotypes_stage(self, nsnps):
        """ccu_demux_genotypes_stage: device addresses of the float matrix [nsnps][nv][3], the per-SNP error (f64)
        and the has-genotype flags (u8), to be filled on the engine's stream before genotypes_commit."""
        gp, err, hg = C.c_void_p(0), C.c_void_p(0), C.c_void_p(0)
        self._check(self.lib.ccu_demux_genotypes_stage(self.h, C.c_int64(nsnps), C.byref(gp), C.byref(err), C.byref(hg)))
        return gp.value, err.value, hg.value

    def genotypes_commit(self, nsnps, use_has_gp=False):
        self._check(self.lib.ccu_demux_genotypes_commit(self.h, C.c_int64(nsnps), C.c_int32(int(use_has_gp))))
        self.nsnps = nsnps

    def copy_genotypes_from(self, other):
        """ccu_demux_copy_genotypes: take `other`'s mixed genotype matrix (device-to-device; NVLink across GPUs)."""
        self._check(self.lib.ccu_demux_copy_genotypes(self.h, other.h))
        self.nsnps = other.nsnps

    def results_dev(self):
        """(device pointer, nbcs) of the result records of the last run."""
        n = C.c_int64(0)
        p = self.lib.ccu_demux_results_dev(self.h, C.byref(n))
        return p, n.value

    @staticmethod
    def _csr(cell_ptr, snp_idx, read_ptr, al, bq):
        return (np.ascontiguousarray(cell_ptr, np.int64), np.ascontiguousarray(snp_idx, np.int32),
                np.ascontiguousarray(read_ptr, np.int64), np.ascontiguousarray(al, np.uint8),
                np.ascontiguousarray(bq, np.uint8))

    def score(self, cell_ptr, snp_idx, read_ptr, al, bq, out=None):
        """ccu_demux_score: host buffers in, results out (H2D + kernels + D2H)."""
        cp, si, rp, a, q = self._csr(cell_ptr, snp_idx, read_ptr, al, bq)
        nbcs = len(cp) - 1
        if out is None:
            out = np.zeros(nbcs, RESULT_DTYPE)
        self._check(self.lib.ccu_demux_score(self.h, C.c_int64(nbcs), _ptr(cp), _ptr(si), _ptr(rp), _ptr(a),
                                             _ptr(q), _ptr(out)))
        self.nbcs = nbcs
        self.nnz = int(cp[-1])
        return out

    def upload(self, cell_ptr, snp_idx, read_ptr, al, bq):
        cp, si, rp, a, q = self._csr(cell_ptr, snp_idx, read_ptr, al, bq)
        self.nbcs = len(cp) - 1
        self.nnz = int(cp[-1])
        self._check(self.lib.ccu_demux_upload(self.h, C.c_int64(self.nbcs), _ptr(cp), _ptr(si), _ptr(rp),
                                              _ptr(a), _ptr(q)))

    def run(self):
        self._check(self.lib.ccu_demux_run(self.h))

    def wait(self):
        self._check(self.lib.ccu_demux_wait(self.h))

    def download(self, out=None):
        if out is None:
            out = np.zeros(self.nbcs, RESULT_DTYPE)
        self._check(self.lib.ccu_demux_download(self.h, _ptr(out)))
        return out

    def timings(self):
        t = Timings()
        self._check(self.lib.ccu_get_timings(self.h, C.byref(t)))
        return t.as_dict()

    def tiles(self):
        out = np.zeros((self.nnz, self.nA * 9), np.float64)
        self._check(self.lib.ccu_demux_get_tiles(self.h, _ptr(out)))
        return out

    def genotypes(self, snp_begin=0, snp_end=None):
        snp_end = self.nsnps if snp_end is None else snp_end
        out = np.zeros((snp_end - snp_begin, self.nv, 3), np.float64)
        self._check(self.lib.ccu_demux_get_genotypes(self.h, C.c_int64(snp_begin), C.c_int64(snp_end), _ptr(out)))
        return out

    def llk(self, cell_begin=0, cell_end=None):
        cell_end = self.nbcs if cell_end is None else cell_end
        out = np.zeros((cell_end - cell_begin, self.nv, self.nv, self.nA), np.float64)
        self._check(self.lib.ccu_demux_get_llk(self.h, C.c_int64(cell_begin), C.c_int64(cell_end), _ptr(out)))
        return out

    def singlets(self, want_llk0=True):
        """(llk[nbcs][nv], llk0[nbcs] or None): every sample's singlet log-likelihood of the last run and llks00[0]."""
        llk = np.zeros((self.nbcs, self.nv), np.float64)
        llk0 = np.zeros(self.nbcs, np.float64) if want_llk0 else None
        self._check(self.lib.ccu_demux_get_singlets(self.h, _ptr(llk), _ptr(llk0)))
        return llk, llk0

    def fp64_peak_tflops(self, ms=200.0):
        v = C.c_double(0)
        self._check(self.lib.ccu_measure_fp64_peak(self.h, C.c_float(ms), C.byref(v)))
        return v.value


PILEUP_DTYPE = np.dtype([
    ("nreads", "<i4"), ("nref", "<i4"), ("nalt", "<i4"), ("_pad", "<i4"),
    ("gls", "<f8", (9,)), ("logdenom", "<f8")])

PAIR_DIST_DTYPE = np.dtype([
    ("id1", "<i4"), ("id2", "<i4"), ("nsnps", "<i4"), ("nread1", "<i4"), ("nread2", "<i4"), ("_pad", "<i4"),
    ("llk0", "<f8"), ("llk2", "<f8")])

FMX_RESULT_DTYPE = np.dtype([
    ("sBest", "<i4"), ("sNext", "<i4"), ("dBest1", "<i4"), ("dBest2", "<i4"), ("dNext1", "<i4"),
    ("dNext2", "<i4"), ("type", "<i4"), ("jBest", "<i4"), ("kBest", "<i4"), ("jNext", "<i4"),
    ("kNext", "<i4"), ("_pad", "<i4"),
    ("sngBestLLK", "<f8"), ("sngNextLLK", "<f8"), ("dblBestLLK", "<f8"), ("dblNextLLK", "<f8"),
    ("sumLLK", "<f8"), ("sngLLK", "<f8"), ("bestLLK", "<f8"), ("nextLLK", "<f8"),
    ("bestPP", "<f8"), ("sngPP", "<f8"), ("sngOnlyPP", "<f8")])


class FmxTimings(C.Structure):
    _fields_ = [("ms_pileup", C.c_float), ("ms_csc", C.c_float), ("ms_lmix", C.c_float), ("ms_cgp", C.c_float),
                ("ms_estep", C.c_float), ("ms_mstep", C.c_float), ("launches", C.c_int32), ("_pad", C.c_int32)]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_ if k != "_pad"}


class FreemuxEngine(DemuxEngine):
    """One GPU's freemuxlet engine: the call sequence a patched cmdCramFreemuxlet makes
    (cmd_cram_freemuxlet.cpp:83-653).  The multi-GPU driver is cramore_b200.fmx_dist."""

    def configure_fmx(self, nclust, doublet_prior=0.5, geno_error=0.0):
        self._check(self.lib.ccu_fmx_configure(self.h, C.c_int32(nclust), C.c_double(doublet_prior),
                                               C.c_double(geno_error)))
        self.K = nclust
        self.npairs = nclust * (nclust + 1) // 2

    def upload_fmx(self, cell_ptr, snp_idx, read_ptr, al, bq, af):
        cp, si, rp, a, q = self._csr(cell_ptr, snp_idx, read_ptr, al, bq)
        af = np.ascontiguousarray(af, np.float64)
        self.nbcs, self.nnz, self.nsnps = len(cp) - 1, int(cp[-1]), len(af)
        self.cell_begin, self.cell_end = 0, self.nbcs
        self._peer_world = None  # peer mappings belong to the arrays of one upload
        self._check(self.lib.ccu_fmx_upload(self.h, C.c_int64(self.nbcs), C.c_int64(self.nsnps), _ptr(cp), _ptr(si),
                                            _ptr(rp), _ptr(a), _ptr(q), _ptr(af)))

    def set_shard(self, cell_begin, cell_end, snp_begin, snp_end):
        self._check(self.lib.ccu_fmx_set_shard(self.h, C.c_int64(cell_begin), C.c_int64(cell_end),
                                               C.c_int64(snp_begin), C.c_int64(snp_end)))
        self.cell_begin, self.cell_end = cell_begin, cell_end

    def lmix(self):
        a, b = np.zeros(self.nbcs), np.zeros(self.nbcs)
        self._check(self.lib.ccu_fmx_lmix(self.h, _ptr(a), _ptr(b)))
        return a, b

    def pileups(self):
        out = np.zeros(self.nnz, PILEUP_DTYPE)
        self._check(self.lib.ccu_fmx_get_pileups(self.h, _ptr(out)))
        return out

    def mstep(self, assign):
        a = np.ascontiguousarray(assign, np.int32)
        assert len(a) == self.nbcs
        self._check(self.lib.ccu_fmx_mstep(self.h, _ptr(a)))

    def estep(self, apply_geno_error=False):
        n = self.cell_end - self.cell_begin
        out = np.zeros(n, FMX_RESULT_DTYPE)
        assign = np.zeros(n, np.int32)
        self._check(self.lib.ccu_fmx_estep(self.h, C.c_int32(1 if apply_geno_error else 0), _ptr(out), _ptr(assign)))
        return out, assign

    def fmx_llk(self):
        out = np.zeros((self.cell_end - self.cell_begin, self.npairs))
        self._check(self.lib.ccu_fmx_get_llk(self.h, _ptr(out)))
        return out

    def clusters(self):
        out = np.zeros((self.K, self.nsnps), PILEUP_DTYPE)
        self._check(self.lib.ccu_fmx_get_clusters(self.h, _ptr(out)))
        return out

    def stats_dev(self):
        """(device pointer, number of doubles) of the [K][nsnps][12] statistics array."""
        n = C.c_int64(0)
        p = self.lib.ccu_fmx_stats_dev(self.h, C.byref(n))
        return p, n.value

    def assign_dev(self):
        n = C.c_int64(0)
        p = self.lib.ccu_fmx_assign_dev(self.h, C.byref(n))
        return p, n.value

    def post_dev(self):
        n = C.c_int64(0)
        p = self.lib.ccu_fmx_post_dev(self.h, C.byref(n))
        return p, n.value

    def posteriors_dev(self, apply_geno_error=False):
        self._check(self.lib.ccu_fmx_posteriors_dev(self.h, C.c_int32(1 if apply_geno_error else 0)))

    def estep_post_dev(self):
        self._check(self.lib.ccu_fmx_estep_post_dev(self.h))

    # ---- peer-memory exchange (CUDA IPC over NVLink), see include/cramore_cuda.h ----
    IPC_BYTES = 4 * 64

    def peer_handles(self):
        buf = C.create_string_buffer(self.IPC_BYTES)
        self._check(self.lib.ccu_fmx_peer_handles(self.h, buf))
        return buf.raw

    def peer_open(self, rank, all_handles):
        """all_handles: the peer_handles() bytes of every rank, in rank order."""
        blob = b"".join(all_handles)
        if len(blob) != self.IPC_BYTES * len(all_handles):
            raise ValueError("peer_open: every rank contributes %d bytes" % self.IPC_BYTES)
        self._check(self.lib.ccu_fmx_peer_open(self.h, C.c_int32(len(all_handles)), C.c_int32(rank), blob))

    def peer_close(self):
        self._check(self.lib.ccu_fmx_peer_close(self.h))

    def mstep_fused_dev(self, apply_geno_error_next=False, want_stats=False, in_kernel_sync=True):
        """M-step of this shard's slab + posteriors of the next E-step, stored into every rank's arrays."""
        self._check(self.lib.ccu_fmx_mstep_fused_dev(self.h, C.c_int32(int(bool(apply_geno_error_next))),
                                                     C.c_int32(int(bool(want_stats))), C.c_int32(int(bool(in_kernel_sync)))))

    def estep_fused_dev(self, in_kernel_sync=True):
        """E-step + decision of this shard's droplets, assignments stored into every rank's vector."""
        self._check(self.lib.ccu_fmx_estep_fused_dev(self.h, C.c_int32(int(bool(in_kernel_sync)))))

    def rank_barrier_dev(self):
        self._check(self.lib.ccu_fmx_rank_barrier_dev(self.h))

    def rank_barrier_failed(self):
        """True when a kernel of the multi-GPU loop gave up waiting for a peer; last_error() then says who and where."""
        v = C.c_int32(0)
        self._check(self.lib.ccu_fmx_rank_barrier_failed(self.h, C.byref(v)))
        return bool(v.value)

    def last_error(self):
        return self.lib.ccu_last_error(self.h).decode()

    def mstep_dev(self):
        self._check(self.lib.ccu_fmx_mstep_dev(self.h))

    def estep_dev(self, apply_geno_error=False):
        self._check(self.lib.ccu_fmx_estep_dev(self.h, C.c_int32(1 if apply_geno_error else 0)))

    def download_results(self):
        out = np.zeros(self.cell_end - self.cell_begin, FMX_RESULT_DTYPE)
        self._check(self.lib.ccu_fmx_download_results(self.h, _ptr(out)))
        return out

    def run_em(self, init_assign, niter=10, geno_error_every_iter=False, stop_when_stable=False):
        a = np.ascontiguousarray(init_assign, np.int32)
        out = np.zeros(self.nbcs, FMX_RESULT_DTYPE)
        it = C.c_int32(0)
        self._check(self.lib.ccu_fmx_run_em(self.h, _ptr(a), C.c_int32(niter), C.c_int32(int(geno_error_every_iter)),
                                            C.c_int32(int(stop_when_stable)), _ptr(out), C.byref(it)))
        return out, it.value

    def pair_distances(self, download=True):
        """Sparse dropDs matrix of cmd_cram_freemuxlet.cpp:172-221: records sorted by (id1 > id2)."""
        n, nrec, ms = C.c_int64(0), C.c_int64(0), C.c_float(0)
        self._check(self.lib.ccu_fmx_pair_distances(self.h, C.byref(n), C.byref(nrec), C.byref(ms)))
        self.pair_stats = {"pairs": n.value, "records": nrec.value, "ms": ms.value}
        if not download:
            return None
        out = np.zeros(n.value, PAIR_DIST_DTYPE)
        self._check(self.lib.ccu_fmx_get_pair_distances(self.h, _ptr(out)))
        return out

    def voting_pairs(self, bf_thres=5.41):
        """After pair_distances(): only the pairs with |llk2 - llk0| > bf_thres (device-side compaction)."""
        n = C.c_int64(0)
        self._check(self.lib.ccu_fmx_get_voting_pairs(self.h, C.c_double(bf_thres), C.byref(n), None))
        out = np.zeros(n.value, PAIR_DIST_DTYPE)
        if n.value:
            self._check(self.lib.ccu_fmx_get_voting_pairs(self.h, C.c_double(bf_thres), C.byref(n), _ptr(out)))
        return out

    def selftest_division(self, nsets=1 << 22, seed=1):
        v = C.c_int64(-1)
        self._check(self.lib.ccu_fmx_selftest_division(self.h, C.c_int64(nsets), C.c_uint64(seed), C.byref(v)))
        return v.value

    def fmx_timings(self):
        t = FmxTimings()
        self._check(self.lib.ccu_fmx_get_timings(self.h, C.byref(t)))
        return t.as_dict()


def run_em_multi(engines, cell_ptr, snp_idx, init_assign, niter=10, geno_error_every_iter=False, want_clusters=False):
    """ccu_fmx_run_em_multi: the EM loop sharded over several FreemuxEngine handles of this process (each configured
    and uploaded with the same store), exchange through peer memory.  Returns (results[nbcs], clusters or None)."""
    lib = engines[0].lib
    n = len(engines)
    arr = (C.c_void_p * n)(*[e.h for e in engines])
    cp = np.ascontiguousarray(cell_ptr, np.int64)
    si = np.ascontiguousarray(snp_idx, np.int32)
    a = np.ascontiguousarray(init_assign, np.int32)
    out = np.zeros(engines[0].nbcs, FMX_RESULT_DTYPE)
    cl = np.zeros((engines[0].K, engines[0].nsnps), PILEUP_DTYPE) if want_clusters else None
    rc = lib.ccu_fmx_run_em_multi(arr, C.c_int32(n), _ptr(cp), _ptr(si), _ptr(a), C.c_int32(niter),
                                  C.c_int32(int(geno_error_every_iter)), _ptr(out), _ptr(cl))
    if rc:
        raise EngineError(lib.ccu_last_error(engines[0].h).decode())
    return out, cl


def initial_clusters(cell_scores, pairs, nclust, bf_thres=5.41, frac_init_clust=1.0, iter_init=10,
                     keep_init_missing=False, init=None, seed=1):
    """cmd_cram_freemuxlet.cpp:223-346 on the sparse pair list (host code of libcramore_cuda.so; uses rand())."""
    lib = load_library()
    scores = np.ascontiguousarray(cell_scores, np.float64)
    pairs = np.ascontiguousarray(pairs, PAIR_DIST_DTYPE)
    n = len(scores)
    clusts = np.full(n, -1, np.int32) if init is None else np.ascontiguousarray(init, np.int32).copy()
    rc = lib.ccu_fmx_initial_clusters(C.c_int64(n), C.c_int32(nclust), _ptr(scores), C.c_int64(len(pairs)), _ptr(pairs),
                                      C.c_double(bf_thres), C.c_double(frac_init_clust), C.c_int32(iter_init),
                                      C.c_int32(int(keep_init_missing)), C.c_int32(0 if init is None else 1),
                                      C.c_int32(seed), _ptr(clusts))
    if rc:
        raise ValueError("ccu_fmx_initial_clusters: bad arguments")
    return clusts


def greedy_clusters(cell_ptr, snp_idx, plp, af, cell_scores, nclust, nsnps, frac_init_clust=1.0,
                    singlet_score_thres=-1e300):
    """cmd_cram_freemux2.cpp:217-261 (host code of libcramore_cuda.so)."""
    lib = load_library()
    cell_ptr = np.ascontiguousarray(cell_ptr, np.int64)
    snp_idx = np.ascontiguousarray(snp_idx, np.int32)
    plp = np.ascontiguousarray(plp, PILEUP_DTYPE)
    af = np.ascontiguousarray(af, np.float64)
    scores = np.ascontiguousarray(cell_scores, np.float64)
    n = len(cell_ptr) - 1
    clusts = np.full(n, -1, np.int32)
    rc = lib.ccu_fmx_greedy_clusters(C.c_int64(n), C.c_int64(nsnps), C.c_int32(nclust), _ptr(cell_ptr), _ptr(snp_idx),
                                     _ptr(plp), _ptr(af), _ptr(scores), C.c_double(frac_init_clust),
                                     C.c_double(singlet_score_thres), _ptr(clusts))
    if rc:
        raise ValueError("ccu_fmx_greedy_clusters: bad arguments")
    return clusts


def refine(results, cell_ptr, snp_idx, read_ptr, al, bq, nv, alphas, gp_f32, snp_err, has_gp=None, nthreads=0):
    """ccu_demux_refine: the reported candidates re-scored in the reference's operation order (host code of the
    library, for byte-identical .best rows).  Returns a refined copy of `results`."""
    lib = load_library()
    cp, si, rp, a, q = DemuxEngine._csr(cell_ptr, snp_idx, read_ptr, al, bq)
    al_ = np.ascontiguousarray(alphas, np.float64)
    gp = np.ascontiguousarray(gp_f32, np.float32)
    nsnps = gp.shape[0]
    err = np.ascontiguousarray(np.broadcast_to(np.asarray(snp_err, np.float64), (nsnps,)))
    hg = None if has_gp is None else np.ascontiguousarray(has_gp, np.uint8)
    out = np.ascontiguousarray(results).copy()
    rc = lib.ccu_demux_refine(C.c_int64(len(cp) - 1), _ptr(cp), _ptr(si), _ptr(rp), _ptr(a), _ptr(q), C.c_int32(nv),
                              C.c_int32(len(al_)), _ptr(al_), C.c_int64(nsnps), _ptr(gp), _ptr(err), _ptr(hg),
                              C.c_int32(nthreads), _ptr(out))
    if rc:
        raise EngineError("ccu_demux_refine: bad arguments")
    return out


def classify(results, nv, alphas, doublet_prior=0.5):
    """Host decision of cmd_cram_demuxlet.cpp:921-991 for an array of results."""
    lib = load_library()
    a = np.ascontiguousarray(alphas, np.float64)
    res = np.ascontiguousarray(results)
    calls = np.zeros(len(res), CALL_DTYPE)
    for i in range(len(res)):
        rc = lib.ccu_demux_classify(C.c_void_p(res[i:i + 1].ctypes.data), C.c_int32(nv), C.c_int32(len(a)),
                                    _ptr(a), C.c_double(doublet_prior), C.c_void_p(calls[i:i + 1].ctypes.data))
        if rc != 0:
            raise EngineError("ccu_demux_classify failed")
    return calls


def best_header():
    return load_library().ccu_demux_best_header().decode()


def best_row(int_id, barcode, nsnps_cell, nreads_cell, result, nv, alphas, doublet_prior, sample_ids):
    lib = load_library()
    a = np.ascontiguousarray(alphas, np.float64)
    buf = C.create_string_buffer(4096)
    ids = (C.c_char_p * nv)(*[s.encode() for s in sample_ids])
    res = np.ascontiguousarray(result).reshape(1)
    n = lib.ccu_demux_best_row(buf, C.c_int32(4096), C.c_int32(int_id), barcode.encode(), C.c_uint32(nsnps_cell),
                               C.c_int32(nreads_cell), C.c_void_p(res.ctypes.data), C.c_int32(nv),
                               C.c_int32(len(a)), _ptr(a), C.c_double(doublet_prior), ids)
    if n < 0:
        raise EngineError("ccu_demux_best_row: result has undefined indices")
    return buf.raw[:n].decode()
