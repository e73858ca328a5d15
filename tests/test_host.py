"""CPU tests of the host-side logic: the C-ABI library loads and exports every symbol
include/cramore_cuda.h declares, the engine refuses to run without a GPU (no CPU fallback), the
host decision/row code matches the oracle, barcode sharding and the world_size-2 gather over gloo."""
import ctypes as C
import os
import re
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    names = []
    for fn in sorted(os.listdir(os.path.join(ROOT, "include"))):
        if fn.endswith(".h"):
            src = open(os.path.join(ROOT, "include", fn)).read()
            src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
            names += re.findall(r"\b(ccu_[a-z0-9_]+)\s*\(", src)
    return sorted(set(names))


def test_abi_exports_every_declared_symbol(built):
    from cramore_b200 import ABI_SYMBOLS, load_library
    lib = load_library()
    declared = _declared_symbols()
    assert len(declared) >= 20
    for name in declared:
        assert hasattr(lib, name), "libcramore_cuda.so does not export %s" % name
    assert sorted(ABI_SYMBOLS) == declared, "engine.py ABI_SYMBOLS out of sync with include/*.h"


def test_no_cpu_fallback(built):
    """Without an sm_100 device ccu_create must fail loudly and the product never touches oracle/."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from cramore_b200 import DemuxEngine, EngineError
    with pytest.raises(EngineError) as ei:
        DemuxEngine(0)
    assert "no CPU fallback" in str(ei.value) or "CUDA" in str(ei.value)
    pkg = os.path.join(ROOT, "cramore_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cpp", ".h")):
                txt = open(os.path.join(dirpath, f)).read()
                assert "pyoracle" not in txt and "liboracle" not in txt, "%s references the oracle" % f


def test_host_classify_and_row_match_oracle(built):
    from cramore_b200 import RESULT_DTYPE, best_header, best_row, classify
    from oracle import pyoracle
    rng = np.random.default_rng(3)
    alphas = [0.0, 0.25, 0.5]
    res = np.zeros(200, RESULT_DTYPE)
    res["sBest"], res["sNext"] = 1, 2
    res["dBest1"], res["dBest2"], res["dBestA"] = 0, 3, rng.integers(1, 3, 200)
    res["dNext1"], res["dNext2"], res["dNextA"] = 3, 0, rng.integers(1, 3, 200)
    res["sngBestLLK"] = -rng.uniform(10, 30, 200)
    res["sngNextLLK"] = res["sngBestLLK"] - rng.uniform(0, 5, 200)
    res["dblBestLLK"] = res["sngBestLLK"] + rng.uniform(-5, 5, 200)
    res["dblNextLLK"] = res["dblBestLLK"] - rng.uniform(0, 5, 200)
    res["sngLLK"] = res["sngBestLLK"] + 0.1
    res["sumLLK"] = np.maximum(res["sngLLK"], res["dblBestLLK"]) + 0.2
    mine = classify(res, 4, alphas, 0.5)
    theirs = pyoracle.demux_classify(res.view(pyoracle.DEMUX_RESULT_DTYPE), 4, alphas, 0.5)
    for f in mine.dtype.names:
        if f != "_pad":
            assert np.array_equal(mine[f], theirs[f]), f
    assert set(np.unique(mine["type"])) == {0, 1, 2}
    ids = ["S0", "S1", "S2", "S3"]
    for i in range(0, 200, 17):
        a = best_row(i, "BC%d-1" % i, 12, 15, res[i], 4, alphas, 0.5, ids)
        b = pyoracle.demux_best_row(i, "BC%d-1" % i, 12, 15, res[i:i + 1].view(pyoracle.DEMUX_RESULT_DTYPE)[0], 4,
                                    alphas, 0.5, ids)
        assert a == b
    # cmd_cram_demuxlet.cpp:629
    assert best_header() == ("INT_ID\tBARCODE\tNUM.SNPS\tNUM.READS\tDROPLET.TYPE\tBEST.GUESS\tBEST.LLK\tNEXT.GUESS\t"
                             "NEXT.LLK\tDIFF.LLK.BEST.NEXT\tBEST.POSTERIOR\tSNG.POSTERIOR\tSNG.BEST.GUESS\t"
                             "SNG.BEST.LLK\tSNG.NEXT.GUESS\tSNG.NEXT.LLK\tSNG.ONLY.POSTERIOR\tDBL.BEST.GUESS\t"
                             "DBL.BEST.LLK\tDIFF.LLK.SNG.DBL\n")


def test_balanced_ranges_and_slices():
    from cramore_b200.shard import balanced_ranges, slice_csr
    from cramore_b200.synth import make_dataset
    d = make_dataset("C1", nbcs=301, nsnps=500)
    for world in (1, 2, 3, 8):
        rs = balanced_ranges(d["cell_ptr"], world)
        assert rs[0][0] == 0 and rs[-1][1] == 301 and all(rs[i][1] == rs[i + 1][0] for i in range(world - 1))
        nnz = [int(d["cell_ptr"][e] - d["cell_ptr"][b]) for b, e in rs]
        assert max(nnz) - min(nnz) <= 2 * int(np.diff(d["cell_ptr"]).max())
        tot_reads = 0
        for b, e in rs:
            cp, si, rp, al, bq = slice_csr(d["cell_ptr"], d["snp_idx"], d["read_ptr"], d["al"], d["bq"], b, e)
            assert cp[0] == 0 and rp[0] == 0 and len(si) == cp[-1] and len(al) == rp[-1] == len(bq)
            tot_reads += len(al)
        assert tot_reads == len(d["al"])
    assert balanced_ranges(np.zeros(1, np.int64), 4) == [(0, 0)] * 4  # empty store


def test_synth_is_deterministic_and_hex_ordered():
    from cramore_b200.synth import _hex_order_key, make_dataset
    a = make_dataset("C1", nbcs=50, nsnps=300)
    b = make_dataset("C1", nbcs=50, nsnps=300)
    for k in ("gp", "cell_ptr", "snp_idx", "read_ptr", "al", "bq"):
        assert np.array_equal(a[k], b[k])
    assert np.all(np.diff(a["cell_ptr"]) >= 0) and np.all(np.diff(a["read_ptr"]) >= 1)
    for c in range(50):  # SNP ids ascend inside a droplet (std::map order, sc_drop_seq.h:165)
        s = a["snp_idx"][a["cell_ptr"][c]:a["cell_ptr"][c + 1]]
        assert np.all(np.diff(s) > 0)
    cnt = np.asarray([9, 10, 15, 16, 255, 256, 4095, 4096])
    left, nd = _hex_order_key(cnt)
    order = np.lexsort((nd, left))
    assert [format(int(x), "x") for x in cnt[order]] == sorted(format(int(x), "x") for x in cnt)


_WORKER = r"""
import os, sys
import numpy as np
import torch.distributed as dist
sys.path.insert(0, %(root)r)
from cramore_b200.shard import score_sharded
from cramore_b200.synth import make_dataset
dist.init_process_group("gloo")
d = make_dataset("C1", nbcs=97, nsnps=400)
DT = np.dtype([("cell", "<i8"), ("nent", "<i4"), ("nread", "<i4"), ("x", "<f8")])
calls = []
def fake_score(cp, si, rp, al, bq):   # stand-in for DemuxEngine.score on a CPU box: per-droplet digests
    calls.append(len(cp) - 1)
    out = np.zeros(len(cp) - 1, DT)
    out["nent"] = np.diff(cp)
    out["nread"] = rp[cp[1:]] - rp[cp[:-1]]
    out["x"] = [float(np.sum(si[cp[i]:cp[i + 1]].astype(np.float64))) for i in range(len(cp) - 1)]
    return out
full = score_sharded(fake_score, d["cell_ptr"], d["snp_idx"], d["read_ptr"], d["al"], d["bq"])
want = fake_score(d["cell_ptr"], d["snp_idx"], d["read_ptr"], d["al"], d["bq"])
assert len(full) == 97 and np.array_equal(full, want), "gathered shards differ from the single-process result"
assert 0 < calls[0] < 97
dist.barrier()
dist.destroy_process_group()
print("rank", dist.get_rank() if dist.is_initialized() else os.environ["RANK"], "ok")
"""


def test_world_size_2_gloo_shard_and_gather(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(_WORKER % {"root": ROOT})
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                        "--master-addr", "127.0.0.1", "--master-port", "29713", str(script)],
                       capture_output=True, text=True, timeout=300, env=env)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert r.stdout.count("ok") == 2


_FMX_WORKER = r"""
import os, sys
import numpy as np
import torch
import torch.distributed as dist
sys.path.insert(0, %(root)r)
from cramore_b200.fmx_dist import run_em_distributed, snp_slabs
from cramore_b200.shard import balanced_ranges
from cramore_b200.synth import make_dataset
from oracle import pyoracle

K, NITER, GE = 4, 3, 0.1
d = make_dataset("C4", nbcs=80, nsnps=600, rbar=80, nv=4)
plp = pyoracle.fmx_pileup(d["read_ptr"], d["al"], d["bq"], 0.5)


class OracleEngine:   # numpy stand-in with the FreemuxEngine methods the driver uses (CPU box: no GPU)
    def __init__(self):
        self.stats = torch.zeros(K * d["nsnps"] * 12, dtype=torch.float64)
        self.assign = torch.full((d["nbcs"],), -1, dtype=torch.int32)
        self.post = torch.zeros(d["nsnps"] * K * 3, dtype=torch.float64)
    def set_shard(self, cb, ce, sb, se):
        self.cb, self.ce, self.sb, self.se = cb, ce, sb, se
    def _clust(self):
        s = self.stats.numpy().reshape(K, d["nsnps"], 12)
        c = np.zeros((K, d["nsnps"]), pyoracle.PILEUP_DTYPE)
        c["gls"] = s[..., :9]
        c["nreads"], c["nref"], c["nalt"] = s[..., 9], s[..., 10], s[..., 11]
        return c
    def mstep_dev(self):
        c = pyoracle.fmx_mstep(d["cell_ptr"], d["snp_idx"], plp, self.assign.numpy(), K, d["nsnps"])
        s = np.zeros((K, d["nsnps"], 12))
        s[:, self.sb:self.se, :9] = c["gls"][:, self.sb:self.se]
        for i, f in enumerate(("nreads", "nref", "nalt")):
            s[:, self.sb:self.se, 9 + i] = c[f][:, self.sb:self.se]
        self.stats.copy_(torch.from_numpy(s.reshape(-1)))
    def estep_dev(self, apply):
        res, _ = pyoracle.fmx_estep(d["cell_ptr"], d["snp_idx"], plp, d["af"], self._clust(), 0.5, GE, apply)
        self.res = res[self.cb:self.ce]
        a = np.where(self.res["type"] == 0, self.res["jBest"], -1).astype(np.int32)
        self.assign[self.cb:self.ce] = torch.from_numpy(a)
    # posterior exchange: the oracle's E-step reads the statistics only through gls[0], gls[4], gls[8], so the
    # stand-in's "posterior" array carries exactly those three numbers per (SNP, cluster); slab only, zeros elsewhere
    def posteriors_dev(self, apply):
        self.apply = apply
        s = self.stats.numpy().reshape(K, d["nsnps"], 12)
        p = np.zeros((d["nsnps"], K, 3))
        p[self.sb:self.se] = s[:, self.sb:self.se][..., [0, 4, 8]].transpose(1, 0, 2)
        self.post.copy_(torch.from_numpy(p.reshape(-1)))
    def estep_post_dev(self):
        p = self.post.numpy().reshape(d["nsnps"], K, 3)
        c = np.zeros((K, d["nsnps"]), pyoracle.PILEUP_DTYPE)
        c["gls"][..., [0, 4, 8]] = p.transpose(1, 0, 2)
        res, _ = pyoracle.fmx_estep(d["cell_ptr"], d["snp_idx"], plp, d["af"], c, 0.5, GE, self.apply)
        self.res = res[self.cb:self.ce]
        a = np.where(self.res["type"] == 0, self.res["jBest"], -1).astype(np.int32)
        self.assign[self.cb:self.ce] = torch.from_numpy(a)
    def download_results(self):
        return self.res


dist.init_process_group("gloo")
world = dist.get_world_size()
rng = np.random.default_rng(1)
init = (d["s1"] %% K).astype(np.int32)
init[rng.random(len(init)) < 0.4] = -1
for mode in ("stats", "posteriors"):
    eng = OracleEngine()
    local, full = run_em_distributed(eng, eng.stats, eng.assign, balanced_ranges(d["cell_ptr"], world),
                                     snp_slabs(d["snp_idx"], d["nsnps"], world), init, niter=NITER,
                                     post=eng.post, exchange=mode)
    results = full if mode == "stats" else results
    assert np.array_equal(full, results), "the two exchange forms differ"
    if mode == "stats":
        stats_first = eng.stats.clone()
    else:
        assert torch.equal(eng.stats, stats_first), "final statistics differ between the exchange forms"
# single-process reference loop
clust = pyoracle.fmx_mstep(d["cell_ptr"], d["snp_idx"], plp, init, K, d["nsnps"])
for it in range(NITER):
    ref, _ = pyoracle.fmx_estep(d["cell_ptr"], d["snp_idx"], plp, d["af"], clust, 0.5, GE, it + 1 == NITER)
    a = np.where(ref["type"] == 0, ref["jBest"], -1).astype(np.int32)
    clust = pyoracle.fmx_mstep(d["cell_ptr"], d["snp_idx"], plp, a, K, d["nsnps"])
assert np.array_equal(full, ref), "sharded EM differs from the single-process loop"
assert np.array_equal(eng._clust()["gls"], clust["gls"]), "all-reduced statistics differ"
dist.barrier()
dist.destroy_process_group()
print("rank ok")
"""


def test_world_size_2_gloo_freemuxlet_em(tmp_path):
    """The multi-GPU EM driver (barcode-sharded E-step, SNP-slab M-step, all-reduce of the statistics)
    over gloo with an oracle-backed stand-in engine: bitwise equal to the single-process loop."""
    script = tmp_path / "fmx_worker.py"
    script.write_text(_FMX_WORKER % {"root": ROOT})
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                        "--master-addr", "127.0.0.1", "--master-port", "29714", str(script)],
                       capture_output=True, text=True, timeout=300, env=env)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert r.stdout.count("rank ok") == 2


# ---------------------------------------------------------------------------------------------
# the `cramore` command line (cramore_b200/host): flag surface and bit-exact CSR construction
# ---------------------------------------------------------------------------------------------
CRAMORE = os.path.join(ROOT, "cramore_b200", "host", "cramore")

REF_DEMUX_FLAGS = ["sam", "tag-group", "tag-UMI", "plp", "vcf", "field", "geno-error-offset", "geno-error-coeff", "r2-info",
                   "min-mac", "min-callrate", "sm", "sm-list", "out", "alpha", "doublet-prior", "sam-verbose",
                   "vcf-verbose", "cap-BQ", "min-BQ", "min-MQ", "min-TD", "excl-flag", "group-list", "min-total",
                   "min-umi", "min-snp"]  # cmd_cram_demuxlet.cpp:40-79, in order
REF_FMX_FLAGS = ["plp", "init-cluster", "out", "nsample", "aux-files", "verbose", "doublet-prior", "geno-error",
                 "bf-thres", "frac-init-clust", "iter-init", "keep-init-missing", "cap-BQ", "min-BQ", "group-list",
                 "min-total", "min-uniq", "min-snp"]  # cmd_cram_freemuxlet.cpp:34-63
REF_DSC_FLAGS = ["sam", "tag-group", "tag-UMI", "exclude-flag", "vcf", "sm", "sm-list", "out", "sam-verbose", "vcf-verbose",
                 "skip-umi", "cap-BQ", "min-BQ", "min-MQ", "min-TD", "excl-flag", "group-list", "min-total", "min-uniq",
                 "min-snp"]  # cmd_cram_dsc_pileup.cpp:39-71


def _help_flags(cmd):
    r = subprocess.run([CRAMORE, cmd, "--help"], capture_output=True, text=True)
    return re.findall(r"^\s+--([A-Za-z0-9-]+)\s+:", r.stderr, flags=re.M)


def test_cli_flag_surface(built):
    d = _help_flags("demuxlet")
    # beyond the reference's flags only engine options (--gpu-*) and the opt-in <out>.single writer
    assert d[:len(REF_DEMUX_FLAGS)] == REF_DEMUX_FLAGS and all(f.startswith("gpu-") or f == "write-single"
                                                              for f in d[len(REF_DEMUX_FLAGS):])
    f = _help_flags("freemuxlet")
    assert f[:len(REF_FMX_FLAGS)] == REF_FMX_FLAGS and all(x.startswith("gpu-") for x in f[len(REF_FMX_FLAGS):])
    assert _help_flags("dsc-pileup") == REF_DSC_FLAGS
    f2 = _help_flags("freemux2")
    assert "seed" in f2 and "randomize-singlet-score" in f2 and "min-umi" in f2 and "min-uniq" not in f2
    # unknown flags and a repeated scalar flag are errors, like paramList::Read (params.cpp:140-177, 478-480)
    r = subprocess.run([CRAMORE, "demuxlet", "--no-such-flag", "1"], capture_output=True, text=True)
    assert r.returncode != 0 and "ignored" in r.stderr
    r = subprocess.run([CRAMORE, "demuxlet", "--out", "a", "--out", "b"], capture_output=True, text=True)
    assert r.returncode != 0 and "Redundant" in r.stderr


@pytest.mark.parametrize("case", ["plain", "filters"])
def test_cli_csr_construction_bit_exact(built, tmp_path, case):
    """plp + VCF text -> the engine's flat inputs, without a GPU (--gpu-dump-inputs): CSR, read order, float GP,
    has_gp and per-SNP error must equal what load_from_plp / parse_posteriors build (bit-exact)."""
    from cramore_b200.synth import make_dataset, write_plp, write_vcf
    d = make_dataset("C1", nbcs=70, nsnps=150, rbar=160, nv=6, seed=31, with_barcodes=True)  # many multi-read entries
    pre = str(tmp_path / "in")
    write_plp(d, pre)
    write_vcf(d, pre + ".vcf", with_r2=True)
    args = [CRAMORE, "demuxlet", "--plp", pre, "--vcf", pre + ".vcf", "--out", str(tmp_path / "o"),
            "--gpu-dump-inputs", str(tmp_path / "dump")]
    keep = np.ones(d["nbcs"], bool)
    err = 0.10
    if case == "filters":
        grp = tmp_path / "grp.txt"
        sel = [b for i, b in enumerate(d["barcodes"]) if i % 3 != 1]
        grp.write_text("\n".join(sel) + "\n")
        keep = np.asarray([i % 3 != 1 for i in range(d["nbcs"])])
        args += ["--group-list", str(grp), "--geno-error-offset", "0.05", "--geno-error-coeff", "0.5", "--min-snp", "100"]
        err = 0.05 + (1 - 0.05) * (1 - float(np.float32(0.9))) * 0.5
        keep &= np.diff(d["cell_ptr"]) >= 100
    r = subprocess.run(args, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-2000:]

    def load(name, dt):
        return np.fromfile(str(tmp_path / ("dump." + name)), dtype=dt)

    # rows come in std::map barcode order, restricted to the droplets that pass the filters
    cells = [i for i in sorted(range(d["nbcs"]), key=lambda i: d["barcodes"][i]) if keep[i]]
    bcs = (tmp_path / "dump.barcodes.txt").read_text().split()
    assert bcs == [d["barcodes"][i] for i in cells]
    from cramore_b200.synth import _hex_order_key, plp_read_order
    from tests.parity import permute_cells
    # reads are numbered by a global counter over the KEPT droplets' .plp rows and iterate in the hex-string order
    # of that number (sc_drop_seq.cpp:365): re-derive the order for the surviving droplets
    o = plp_read_order(d)
    d2 = dict(d, al=np.asarray(d["al"])[o], bq=np.asarray(d["bq"])[o])      # file (counter) order
    kept_ids = [i for i in range(d["nbcs"]) if keep[i]]  # droplets surviving --group-list / --min-snp, id order
    counter = np.full(len(d["al"]), -1, np.int64)
    n = 0
    for i in kept_ids:
        r0, r1 = d["read_ptr"][d["cell_ptr"][i]], d["read_ptr"][d["cell_ptr"][i + 1]]
        counter[r0:r1] = np.arange(n, n + r1 - r0)
        n += r1 - r0
    al_f, bq_f = d2["al"].copy(), d2["bq"].copy()
    rp0 = np.asarray(d["read_ptr"])
    for e in np.nonzero(np.diff(rp0) > 1)[0]:
        c = counter[rp0[e]:rp0[e + 1]]
        if c[0] < 0:
            continue
        left, nd = _hex_order_key(c)
        perm = np.lexsort((nd, left))
        al_f[rp0[e]:rp0[e + 1]] = d2["al"][rp0[e]:rp0[e + 1]][perm]
        bq_f[rp0[e]:rp0[e + 1]] = d2["bq"][rp0[e]:rp0[e + 1]][perm]
    cp, si, rp, al, bq = permute_cells(dict(d, al=al_f, bq=bq_f), cells)
    assert np.array_equal(load("cell_ptr.i64", "<i8"), cp) and np.array_equal(load("snp_idx.i32", "<i4"), si)
    assert np.array_equal(load("read_ptr.i64", "<i8"), rp)
    assert np.array_equal(load("al.u8", "u1"), al) and np.array_equal(load("bq.u8", "u1"), bq)
    ac = np.asarray(d["geno"], np.int64).sum(axis=1)
    has = ((ac >= 1) & (2 * d["nv"] - ac >= 1)).astype(np.uint8)  # --min-mac 1 (bcf_filtered_reader.cpp:638)
    assert np.array_equal(load("has_gp.u8", "u1"), has)
    gp = load("gp.f32", "<f4").reshape(d["nsnps"], d["nv"], 3)
    assert np.array_equal(gp[has == 1], d["gp"][has == 1]) and np.all(gp[has == 0] == 0)
    assert np.array_equal(load("snp_err.f64", "<f8"), np.where(has == 1, err, 0.0))
    if case == "plain":
        ranks = {b: k for k, b in enumerate(sorted(d["barcodes"]))}
        assert np.array_equal(load("row_int_id.i32", "<i4"), [ranks[b] for b in bcs])


def test_cli_field_pl_posteriors(built, tmp_path):
    """--field PL: genotype posteriors from Phred-scaled likelihoods by ten EM iterations on the allele frequency
    (BCFFilteredReader::parse_likelihoods, bcf_filtered_reader.cpp:289-365), checked against a Python restatement of
    those lines through --gpu-dump-inputs (no GPU).  Parity unpinned: the reference's reader is htslib-bound."""
    from cramore_b200.synth import make_dataset, write_plp
    d = make_dataset("C1", nbcs=20, nsnps=60, rbar=40, nv=5, seed=9, with_barcodes=True)
    pre = str(tmp_path / "in")
    write_plp(d, pre)
    rng = np.random.default_rng(4)
    geno = np.asarray(d["geno"])
    pls = np.zeros((d["nsnps"], d["nv"], 3), np.int64)
    with open(pre + ".vcf", "w") as f:
        f.write("##fileformat=VCFv4.2\n##contig=<ID=1>\n")
        f.write("#CHROM\tPOS\tID\tREF\tALT\tQUAL\tFILTER\tINFO\tFORMAT\t" + "\t".join(d["sample_ids"]) + "\n")
        for s in range(d["nsnps"]):
            cols = []
            for j in range(d["nv"]):
                pl = rng.integers(5, 300, 3)  # above 255 is capped (PhredHelper.h:40)
                pl[geno[s, j]] = 0
                pls[s, j] = pl
                cols.append("%s:%d,%d,%d" % (["0/0", "0/1", "1/1"][geno[s, j]], *pl))
            f.write("1\t%d\t.\tA\tG\t.\tPASS\t.\tGT:PL\t%s\n" % (1000 + 10 * s, "\t".join(cols)))
    r = subprocess.run([CRAMORE, "demuxlet", "--plp", pre, "--vcf", pre + ".vcf", "--field", "PL", "--out",
                        str(tmp_path / "o"), "--gpu-dump-inputs", str(tmp_path / "dump")], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-2000:]
    gp = np.fromfile(str(tmp_path / "dump.gp.f32"), np.float32).reshape(d["nsnps"], d["nv"], 3)
    has = np.fromfile(str(tmp_path / "dump.has_gp.u8"), np.uint8)
    w = np.array([1.0, 2.0, 1.0])
    for s in range(d["nsnps"]):
        if not has[s]:
            continue  # dropped by --min-mac 1
        acs = np.array([0.5, 0.5])
        prob = np.power(0.1, np.minimum(pls[s], 255) * 0.1)
        for it in range(10):
            pri = np.array([acs[0] * acs[0], acs[1] * acs[0], acs[1] * acs[1]]) * w
            g = pri * prob
            g /= g.sum(axis=1, keepdims=True)
            new = np.array([(2 * g[:, 0] + g[:, 1]).sum(), (g[:, 1] + 2 * g[:, 2]).sum()])
            acs = new / (2 * d["nv"])
        np.testing.assert_allclose(gp[s], g.astype(np.float32), rtol=2e-6)
    assert has.sum() > 30


def test_cli_sam_path_matches_plp_path(built, tmp_path):
    """--sam (built-in BAM reader, bam_io.h) against --plp on the same reads: the BAM holds one short alignment per
    read of the synthetic store (CIGARs with soft/hard clips, an insertion, a deletion, a reference skip), duplicates
    of (droplet, SNP, UMI) that must lose to the first read, and reads the filter must drop.  The flattened engine
    inputs (--gpu-dump-inputs, no GPU) must agree entry by entry; the reference drops SNPs failing --min-mac on the
    --sam path and keeps them without genotypes on the --plp path, which is accounted for."""
    from cramore_b200.synth import make_dataset, write_bam, write_plp, write_vcf
    d = make_dataset("C1", nbcs=40, nsnps=150, rbar=90, nv=5, seed=13, with_barcodes=True)
    pre = str(tmp_path / "in")
    write_plp(d, pre)
    write_vcf(d, pre + ".vcf", with_r2=True, lead_decoy=True)
    write_bam(d, pre + ".bam", duplicate_every=5)
    common = ["--vcf", pre + ".vcf", "--out", str(tmp_path / "o"), "--geno-error-coeff", "0.5"]
    for mode, src in (("sam", ["--sam", pre + ".bam"]), ("plp", ["--plp", pre])):
        r = subprocess.run([CRAMORE, "demuxlet"] + src + common + ["--gpu-dump-inputs", str(tmp_path / mode)],
                           capture_output=True, text=True)
        assert r.returncode == 0, r.stderr[-2000:]

    def load(mode, name, dt):
        return np.fromfile(str(tmp_path / (mode + "." + name)), dtype=dt)

    has = load("plp", "has_gp.u8", np.uint8).astype(bool)
    assert has.sum() < len(has) and load("sam", "has_gp.u8", np.uint8).all()
    remap = np.cumsum(has)  # SNP id on the --sam path, whose SNP 0 is the decoy record
    assert (tmp_path / "sam.barcodes.txt").read_text() == (tmp_path / "plp.barcodes.txt").read_text()
    assert np.array_equal(load("sam", "row_int_id.i32", np.int32), load("plp", "row_int_id.i32", np.int32))
    nv = 5
    gp_p = load("plp", "gp.f32", np.float32).reshape(-1, nv * 3)[has]
    gp_s = load("sam", "gp.f32", np.float32).reshape(-1, nv * 3)
    err_s = load("sam", "snp_err.f64", np.float64)
    assert np.array_equal(gp_s[1:], gp_p)
    assert np.array_equal(err_s[1:], load("plp", "snp_err.f64", np.float64)[has])
    # the first variant of the file: genotype-error offset 1e-4 inside parse_posteriors, no error mix (:177, :204-209)
    assert err_s[0] == 0.0 and np.allclose(gp_s[0], np.tile([0.01, 0.98, 0.01], nv), atol=2e-4) and gp_s[0, 1] != np.float32(0.98)
    cps, cpp = load("sam", "cell_ptr.i64", np.int64), load("plp", "cell_ptr.i64", np.int64)
    sis, sip = load("sam", "snp_idx.i32", np.int32), load("plp", "snp_idx.i32", np.int32)
    rps, rpp = load("sam", "read_ptr.i64", np.int64), load("plp", "read_ptr.i64", np.int64)
    als, alp = load("sam", "al.u8", np.uint8), load("plp", "al.u8", np.uint8)
    bqs, bqp = load("sam", "bq.u8", np.uint8), load("plp", "bq.u8", np.uint8)
    nent = 0
    for c in range(len(cpp) - 1):
        want = [(remap[sip[e]], sorted(zip(alp[rpp[e]:rpp[e + 1]], bqp[rpp[e]:rpp[e + 1]])))
                for e in range(cpp[c], cpp[c + 1]) if has[sip[e]]]
        got = [(sis[e], sorted(zip(als[rps[e]:rps[e + 1]], bqs[rps[e]:rps[e + 1]]))) for e in range(cps[c], cps[c + 1])]
        assert got == want, c
        nent += len(got)
    assert nent > 2000


def test_bam_base_extraction_follows_the_reference_cigar_walk(built, tmp_path):
    """bam_get_base_and_qual_and_read_and_qual (hts_utils.cpp:279-358) through the CLI: hand-made alignments whose
    SNP base sits behind clips / insertions / deletions, inside a deletion (no observation), or behind '=' / 'X'
    operations, which that function ignores."""
    from cramore_b200.synth import encode_bam_record, write_bam_records
    pre = str(tmp_path / "x")
    with open(pre + ".vcf", "w") as f:
        f.write("##fileformat=VCFv4.2\n##contig=<ID=1>\n#CHROM\tPOS\tID\tREF\tALT\tQUAL\tFILTER\tINFO\tFORMAT\tS1\tS2\n")
        for pos in (101, 121, 141, 161, 181, 201):
            f.write("1\t%d\t.\tA\tG\t.\tPASS\t.\tGT:GP\t0/1:0.01,0.98,0.01\t0/0:0.98,0.01,0.01\n" % pos)
    q = [30] * 12

    def rec(pos0, cigar, seq, name, cb="AAAC-1", umi=None, qual=None):
        return encode_bam_record(0, pos0, 60, 0, cigar, seq, (qual or q)[:len(seq)], name, (("CB", cb), ("UB", umi or name)))

    records = [
        rec(98, [("M", 6)], "TTGTTT", "u1"),                         # SNP 101 (0-based 100) = read index 2: G -> alt
        rec(117, [("S", 3), ("M", 6)], "CCCTTTATT", "u2"),           # SNP 121: 3 clipped + offset 3 -> index 6: A -> ref
        rec(138, [("M", 2), ("D", 3), ("M", 4)], "TTTTTT", "u3"),    # SNP 141 (140) lies in the deletion 140..142: no read
        rec(158, [("M", 1), ("I", 2), ("M", 5)], "TCCTGTTT", "u4"),  # SNP 161 (160): index 1+2+1 = 4: G -> alt
        rec(178, [("=", 3), ("M", 6)], "TTTCTTTTT", "u5"),           # '=' is ignored by the walk: M starts at 178, SNP
                                                                      # 181 (180) -> index 2: T -> neither allele (code 2)
        rec(199, [("M", 6)], "TAGTTT", "u6", qual=[30, 5, 30, 30, 30, 30]),  # SNP 201 (200): index 1, BQ 5 < --min-BQ
    ]
    write_bam_records(pre + ".bam", [("1", 1000)], records)
    r = subprocess.run([CRAMORE, "demuxlet", "--sam", pre + ".bam", "--vcf", pre + ".vcf", "--out", str(tmp_path / "o"),
                        "--gpu-dump-inputs", str(tmp_path / "d")], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-2000:]
    snp = np.fromfile(str(tmp_path / "d.snp_idx.i32"), np.int32)
    al = np.fromfile(str(tmp_path / "d.al.u8"), np.uint8)
    bq = np.fromfile(str(tmp_path / "d.bq.u8"), np.uint8)
    # SNP 5 (position 201): u6 fails --min-BQ there, but u5 does produce an observation.  Its '=' bases are not
    # counted, so the walk ends at read index 6 < read length without having met position 200 and the function
    # hands out that base ('T', neither allele) for every later SNP in the window -- as the reference does
    assert snp.tolist() == [0, 1, 3, 4, 5] and al.tolist() == [1, 0, 1, 2, 2] and bq.tolist() == [20, 20, 20, 20, 20]


def _random_alignments(tmp_path, seed, with_samples=False):
    """A coordinate-sorted BAM over several contigs (one of them unknown to the VCF) with long gapped reads that
    span several SNPs, both strands, flags and mapping qualities on either side of the filters, missing tags, a small
    UMI alphabet (so that droplet/SNP/UMI collisions happen) -- and a VCF with an indel, a multi-allelic record and a
    contig the BAM lacks."""
    from cramore_b200.synth import encode_bam_record, write_bam_records
    rng = np.random.default_rng(seed)
    refs = [("1", 3000), ("2", 2500), ("GL7", 900), ("10", 2000), ("X", 1500)]   # GL7: not in the VCF
    vcf_contigs = ["1", "2", "10", "11", "X"]                                     # 11: not in the BAM
    pre = str(tmp_path / ("r%d" % seed))
    samples = ["S%d" % i for i in range(4)] if with_samples else []
    with open(pre + ".vcf", "w") as f:
        f.write("##fileformat=VCFv4.2\n" + "".join("##contig=<ID=%s>\n" % c for c in vcf_contigs))
        f.write("#CHROM\tPOS\tID\tREF\tALT\tQUAL\tFILTER\tINFO" + ("\tFORMAT\t" + "\t".join(samples) if samples else "") + "\n")
        for c in vcf_contigs:
            length = dict(refs).get(c, 800)
            pos = 40
            while pos < length - 40:
                kind = rng.integers(20)
                ref, alt = "ACGT"[rng.integers(4)], "ACGT"[rng.integers(4)]
                if alt == ref:
                    alt = "ACGT"[("ACGT".index(ref) + 1) % 4]
                if kind == 0:
                    ref += "TT"          # a deletion: rlen 3
                elif kind == 1 and not samples:
                    alt += ",N"          # three alleles
                info = "AF=%.4f" % rng.uniform(0.01, 0.99) if rng.integers(2) else "AC=%d;AN=%d" % (rng.integers(1, 50), 200)
                line = "%s\t%d\t.\t%s\t%s\t.\tPASS\t%s" % (c, pos, ref, alt, info)
                if samples:
                    gts = rng.integers(0, 3, len(samples))
                    line += "\tGT:GP\t" + "\t".join(["0/0:0.98,0.01,0.01", "0/1:0.01,0.98,0.01", "1/1:0.01,0.01,0.98"][g] for g in gts)
                f.write(line + "\n")
                pos += int(rng.integers(3, 60))
    bcs = ["%s-1" % "".join("ACGT"[i] for i in rng.integers(0, 4, 8)) for _ in range(12)]
    records = []
    for tid, (name, length) in enumerate(refs):
        starts = np.sort(rng.integers(0, length - 120, 500 if name != "GL7" else 60))
        for k, pos in enumerate(starts):
            cigar, rlen = [], 0
            if rng.integers(6) == 0:
                cigar.append(("H", int(rng.integers(1, 5))))
            if rng.integers(4) == 0:
                n = int(rng.integers(1, 6))
                cigar.append(("S", n))
                rlen += n
            for blk in range(int(rng.integers(1, 4))):
                n = int(rng.integers(5, 30))
                cigar.append(("M" if rng.integers(12) else "=X"[rng.integers(2)], n))
                rlen += n
                if blk < 2 and rng.integers(2):
                    op = "IDN"[rng.integers(3)]
                    n = int(rng.integers(1, 12))
                    cigar.append((op, n))
                    if op == "I":
                        rlen += n
            while cigar[-1][0] in "IDN":
                op, n = cigar.pop()
                if op == "I":
                    rlen -= n
            if rng.integers(5) == 0:
                n = int(rng.integers(1, 4))
                cigar.append(("S", n))
                rlen += n
            seq = "".join("ACGTN"[i] for i in rng.choice(5, rlen, p=[0.24, 0.24, 0.24, 0.24, 0.04]))
            qual = [int(q) for q in rng.integers(2, 42, rlen)]
            flag = (0x10 if rng.integers(2) else 0) | int(rng.choice([0, 0, 0, 0, 0, 0x400, 0x100, 0x4, 0x200, 0x800]))
            tags = []
            if rng.integers(25):
                tags.append(("CB", bcs[rng.integers(len(bcs))]))
            if rng.integers(25):
                tags.append(("UB", "".join("ACGT"[i] for i in rng.integers(0, 4, 2))))
            mapq = int(rng.choice([0, 3, 19, 20, 60, 255]))
            records.append(encode_bam_record(tid, int(pos), mapq, flag, cigar, seq, qual, "q%d_%d" % (tid, k), tuple(tags)))
    write_bam_records(pre + ".bam", refs, records)
    return pre, bcs


def _gunzip(path):
    import gzip
    with gzip.open(path, "rt") as f:
        return f.read()


@pytest.mark.parametrize("seed,extra", [(1, {}), (2, dict(min_td=3, cap_bq=30, min_bq=20)), (3, dict(skip_umi=True)),
                                        (4, dict(group=True, exclude_flag=0x10)), (5, dict(min_mq=0, excl_flag=0))])
def test_dsc_pileup_files_match_the_python_restatement(built, tmp_path, seed, extra):
    """`cramore dsc-pileup` (cmd_dsc_pileup.cpp + the streaming join of bam_io.h) against oracle/dsc_pileup_py.py, an
    independent reading of cmd_cram_dsc_pileup.cpp: the four output files byte for byte."""
    from oracle import dsc_pileup_py
    pre, bcs = _random_alignments(tmp_path, seed)
    args = [CRAMORE, "dsc-pileup", "--sam", pre + ".bam", "--vcf", pre + ".vcf", "--out", pre + ".out"]
    kw = {}
    for key, flag in (("min_td", "--min-TD"), ("cap_bq", "--cap-BQ"), ("min_bq", "--min-BQ"), ("min_mq", "--min-MQ"),
                      ("excl_flag", "--excl-flag"), ("exclude_flag", "--exclude-flag")):
        if key in extra:
            args += [flag, str(extra[key])]
            kw[key] = extra[key]
    if extra.get("skip_umi"):
        args.append("--skip-umi")
        kw["skip_umi"] = True
    if extra.get("group"):
        with open(pre + ".grp", "w") as f:
            f.write("\n".join(bcs[::2] + ["."]) + "\n")
        args += ["--group-list", pre + ".grp"]
        kw["group_list"] = set(bcs[::2] + ["."])
    r = subprocess.run(args, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-3000:]
    want = dsc_pileup_py.dsc_pileup(pre + ".bam", pre + ".vcf", **kw)
    assert len(want["plp"].splitlines()) > 150 // (2 if extra.get("group") else 1)
    for ext in ("cel", "var", "plp"):
        assert _gunzip(pre + ".out.%s.gz" % ext) == want[ext], ext
    if extra.get("skip_umi"):
        assert not os.path.exists(pre + ".out.umi.gz")
    else:
        assert _gunzip(pre + ".out.umi.gz") == want["umi"]


@pytest.mark.parametrize("seed", [11, 12])
def test_demuxlet_sam_join_matches_the_python_restatement(built, tmp_path, seed):
    """The droplet store `cramore demuxlet --sam` hands to the engine (--gpu-dump-inputs, no GPU) against the Python
    restatement of cmd_cram_demuxlet.cpp:222-395 on multi-contig gapped alignments: rows in barcode order, SNP ids in
    the order the VCF cursor met them, reads of an entry in UMI string order, first read of a UMI wins."""
    from oracle import dsc_pileup_py
    pre, bcs = _random_alignments(tmp_path, seed, with_samples=True)
    r = subprocess.run([CRAMORE, "demuxlet", "--sam", pre + ".bam", "--vcf", pre + ".vcf", "--out", pre + ".o", "--min-mac", "0",
                        "--min-callrate", "0", "--gpu-dump-inputs", pre + ".d"], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-3000:]
    want = dsc_pileup_py.dsc_pileup(pre + ".bam", pre + ".vcf", cap_bq=20, demuxlet_mode=True)
    per_cell = {}
    for snp, cells in want["snp_umis"].items():
        for c, umis in cells.items():
            per_cell.setdefault(c, {})[snp] = [umis[u] for u in sorted(umis)]
    rows = sorted((b, i) for i, b in enumerate(want["barcodes"]) if i in per_cell)
    assert (tmp_path / ("r%d.d.barcodes.txt" % seed)).read_text().split() == [b for b, _ in rows]
    snp_idx, als, bqs, cell_ptr, read_ptr = [], [], [], [0], [0]
    for _, i in rows:
        for snp in sorted(per_cell[i]):
            snp_idx.append(snp)
            als += [a for a, _ in per_cell[i][snp]]
            bqs += [q for _, q in per_cell[i][snp]]
            read_ptr.append(len(als))
        cell_ptr.append(len(snp_idx))
    assert len(snp_idx) > 200
    for name, dt, exp in (("cell_ptr.i64", np.int64, cell_ptr), ("snp_idx.i32", np.int32, snp_idx), ("read_ptr.i64", np.int64, read_ptr),
                          ("al.u8", np.uint8, als), ("bq.u8", np.uint8, bqs)):
        assert np.fromfile(pre + ".d." + name, dt).tolist() == exp, name
    assert len(np.fromfile(pre + ".d.has_gp.u8", np.uint8)) == len(want["snps"])


def test_dsc_pileup_output_feeds_demuxlet_plp(built, tmp_path):
    """BAM -> `cramore dsc-pileup` -> `cramore demuxlet --plp`: the same engine inputs as from the digital pileup the
    generator writes directly (reads of an entry compared as multisets: the UMI strings, hence the read order inside
    an entry, are not the generator's).  --sm restricts the VCF to one sample, which switches the allele frequency
    of the .var.gz file from INFO/AF to that sample's genotype (bcf_filtered_reader.cpp:875-878)."""
    from cramore_b200.synth import make_dataset, write_bam, write_plp, write_vcf
    d = make_dataset("C1", nbcs=30, nsnps=120, rbar=80, nv=4, seed=5, with_barcodes=True)
    pre = str(tmp_path / "in")
    write_plp(d, pre)
    write_vcf(d, pre + ".vcf")
    write_bam(d, pre + ".bam", duplicate_every=4)
    r = subprocess.run([CRAMORE, "dsc-pileup", "--sam", pre + ".bam", "--vcf", pre + ".vcf", "--out", pre + ".dsc", "--sm",
                        d["sample_ids"][0]], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-3000:]
    var = _gunzip(pre + ".dsc.var.gz").splitlines()[1:]
    geno0 = np.asarray(d["geno"])[:, 0]
    assert [ln.split("\t")[5] for ln in var[:len(geno0)]] == ["%.5f" % (g / 2.0) for g in geno0[:len(var)]]
    assert len(var) == d["nsnps"]
    dumps = {}
    for mode, prefix in (("dsc", pre + ".dsc"), ("gen", pre)):
        r = subprocess.run([CRAMORE, "demuxlet", "--plp", prefix, "--vcf", pre + ".vcf", "--out", str(tmp_path / "o"),
                            "--gpu-dump-inputs", str(tmp_path / mode)], capture_output=True, text=True)
        assert r.returncode == 0, r.stderr[-3000:]
        dumps[mode] = {n: np.fromfile(str(tmp_path / (mode + "." + n)), dt) for n, dt in (
            ("cell_ptr.i64", np.int64), ("snp_idx.i32", np.int32), ("read_ptr.i64", np.int64), ("al.u8", np.uint8),
            ("bq.u8", np.uint8), ("gp.f32", np.float32), ("snp_err.f64", np.float64), ("has_gp.u8", np.uint8),
            ("row_int_id.i32", np.int32))}
        dumps[mode]["bcs"] = (tmp_path / (mode + ".barcodes.txt")).read_text()
    a, b = dumps["dsc"], dumps["gen"]
    for n in ("cell_ptr.i64", "snp_idx.i32", "read_ptr.i64", "gp.f32", "snp_err.f64", "has_gp.u8", "row_int_id.i32", "bcs"):
        assert np.array_equal(a[n], b[n]), n
    rp = a["read_ptr.i64"]
    for e in range(len(rp) - 1):
        assert sorted(zip(a["al.u8"][rp[e]:rp[e + 1]], a["bq.u8"][rp[e]:rp[e + 1]])) == \
               sorted(zip(b["al.u8"][rp[e]:rp[e + 1]], b["bq.u8"][rp[e]:rp[e + 1]])), e


def test_vcf_plain_gzip_and_bgzf_give_the_same_inputs(built, tmp_path):
    """The text readers take plain files, gzip and BGZF (bgzip / bcftools -Oz; blocks inflated in parallel): the engine
    inputs built from the three forms of one VCF are identical, and a damaged block is an error, not silent garbage."""
    import gzip
    import shutil
    from cramore_b200.synth import bgzip, make_dataset, write_plp, write_vcf
    d = make_dataset("C1", nbcs=30, nsnps=3000, rbar=60, nv=12, seed=9, with_barcodes=True)   # ~0.5 MB of VCF: 9 blocks
    pre = str(tmp_path / "in")
    write_plp(d, pre)
    write_vcf(d, pre + ".vcf")
    with open(pre + ".vcf", "rb") as fi, gzip.open(pre + ".gz.vcf.gz", "wb") as fo:
        shutil.copyfileobj(fi, fo)
    bgzip(pre + ".vcf", pre + ".bgzf.vcf.gz")
    assert os.path.getsize(pre + ".bgzf.vcf.gz") < os.path.getsize(pre + ".vcf") / 3
    blobs = []
    for vcf in (pre + ".vcf", pre + ".gz.vcf.gz", pre + ".bgzf.vcf.gz"):
        dump = str(tmp_path / ("d%d" % len(blobs)))
        r = subprocess.run([CRAMORE, "demuxlet", "--plp", pre, "--vcf", vcf, "--out", str(tmp_path / "o"), "--gpu-dump-inputs", dump],
                           capture_output=True, text=True)
        assert r.returncode == 0, r.stderr[-2000:]
        blobs.append(b"".join(open(dump + "." + n, "rb").read() for n in ("gp.f32", "snp_err.f64", "has_gp.u8", "snp_idx.i32")))
    assert blobs[0] == blobs[1] == blobs[2] and len(blobs[0]) > 400000
    raw = bytearray(open(pre + ".bgzf.vcf.gz", "rb").read())
    raw[len(raw) // 2] ^= 0x5a
    open(pre + ".bad.vcf.gz", "wb").write(bytes(raw))
    r = subprocess.run([CRAMORE, "demuxlet", "--plp", pre, "--vcf", pre + ".bad.vcf.gz", "--out", str(tmp_path / "o"),
                        "--gpu-dump-inputs", str(tmp_path / "bad")], capture_output=True, text=True)
    assert r.returncode != 0 and "BGZF" in r.stderr


def test_bcf_input_equals_vcf_input(built, tmp_path):
    """A BCF2 file (written by an independent encoder, cramore_b200.synth.write_bcf) gives the engine inputs of the
    text VCF with the same records, bit for bit: float32 posteriors, the R2-dependent error, the --min-mac filter
    from GT, and the dsc-pileup .var.gz from the same sites."""
    from cramore_b200.synth import make_dataset, write_bam, write_bcf, write_plp, write_vcf
    d = make_dataset("C1", nbcs=25, nsnps=900, rbar=60, nv=9, seed=19, with_barcodes=True)
    pre = str(tmp_path / "in")
    write_plp(d, pre)
    write_vcf(d, pre + ".vcf", with_r2=True)
    write_bcf(d, pre + ".bcf", with_r2=True)
    blobs = []
    for vcf in (pre + ".vcf", pre + ".bcf"):
        for field in ("GP", "GT"):
            dump = str(tmp_path / ("d%d" % len(blobs)))
            r = subprocess.run([CRAMORE, "demuxlet", "--plp", pre, "--vcf", vcf, "--field", field, "--geno-error-coeff", "0.4",
                                "--out", str(tmp_path / "o"), "--gpu-dump-inputs", dump], capture_output=True, text=True)
            assert r.returncode == 0, r.stderr[-2000:]
            blobs.append(b"".join(open(dump + "." + n, "rb").read() for n in ("gp.f32", "snp_err.f64", "has_gp.u8", "snp_idx.i32"))
                         + open(dump + ".samples.txt", "rb").read())
    assert blobs[0] == blobs[2] and blobs[1] == blobs[3] and blobs[0] != blobs[1] and len(blobs[0]) > 90000
    has = np.fromfile(str(tmp_path / "d0.has_gp.u8"), np.uint8)
    assert 0 < has.sum() < len(has)   # the GT-based filter dropped some sites, in both forms alike
    # the same through the BAM x VCF join of dsc-pileup (--sm: allele frequencies from the genotypes)
    write_bam(d, pre + ".bam")
    outs = []
    for vcf in (pre + ".vcf", pre + ".bcf"):
        out = str(tmp_path / ("p%d" % len(outs)))
        r = subprocess.run([CRAMORE, "dsc-pileup", "--sam", pre + ".bam", "--vcf", vcf, "--out", out, "--sm", d["sample_ids"][2],
                            "--sm", d["sample_ids"][5]], capture_output=True, text=True)
        assert r.returncode == 0, r.stderr[-2000:]
        outs.append(_gunzip(out + ".var.gz") + _gunzip(out + ".plp.gz"))
    assert outs[0] == outs[1] and len(outs[0]) > 20000


def test_bcf_decoder_on_hand_encoded_records(built, tmp_path):
    """The BCF2 decoder on records built byte by byte from the specification (VCF 4.2, section 6.3): dictionary order
    with and without IDX=, typed integers of 8/16/32 bits, a length that needs the 15+ escape, missing values and
    end-of-vector padding, phased and half-missing genotypes, flags, strings, several filters, multi-allelic sites."""
    import struct
    from cramore_b200.synth import _bgzf_block
    text = ("##fileformat=VCFv4.2\n##FILTER=<ID=PASS,Description=\"ok\">\n##FILTER=<ID=q10,Description=\"low\">\n"
            "##contig=<ID=chrA>\n##contig=<ID=chrB>\n"
            "##INFO=<ID=DP,Number=1,Type=Integer,Description=\"d\">\n##INFO=<ID=DB,Number=0,Type=Flag,Description=\"f\">\n"
            "##INFO=<ID=AF,Number=A,Type=Float,Description=\"a\">\n##INFO=<ID=NOTE,Number=1,Type=String,Description=\"s\">\n"
            "##FORMAT=<ID=GT,Number=1,Type=String,Description=\"g\">\n##FORMAT=<ID=PL,Number=G,Type=Integer,Description=\"p\">\n"
            "##FORMAT=<ID=GP,Number=G,Type=Float,Description=\"p\",IDX=9>\n"
            "#CHROM\tPOS\tID\tREF\tALT\tQUAL\tFILTER\tINFO\tFORMAT\tS1\tS2\n")
    # dictionary: PASS 0, q10 1, DP 2, DB 3, AF 4, NOTE 5, GT 6, PL 7, GP 9 (IDX)
    hdr = text.encode() + b"\0"

    def tstr(sv):
        b = sv.encode()
        return (bytes([(len(b) << 4) | 7]) if len(b) < 15 else b"\xf7\x11" + bytes([len(b)])) + b

    def rec(chrom, pos0, rid, alleles, qual, filters, info, fmt, n_sample=2):
        sh = struct.pack("<iii", chrom, pos0, len(alleles[0])) + (struct.pack("<I", 0x7F800001) if qual is None else struct.pack("<f", qual))
        sh += struct.pack("<II", (len(alleles) << 16) | len(info), (len(fmt) << 24) | n_sample)
        sh += tstr(rid) + b"".join(tstr(a) for a in alleles)
        sh += bytes([(len(filters) << 4) | 1]) + bytes(filters) if filters else b"\x00"
        sh += b"".join(info)
        ind = b"".join(fmt)
        return struct.pack("<II", len(sh), len(ind)) + sh + ind

    r1 = rec(0, 99, "rs1", ["A", "G"], 30.5, [0],
             [b"\x11\x02" + b"\x12" + struct.pack("<h", 1000),            # DP=1000 as int16
              b"\x11\x03" + b"\x00",                                        # DB flag
              b"\x11\x04" + b"\x15" + struct.pack("<f", 0.25)],             # AF=0.25
             [b"\x11\x06" + b"\x21" + bytes([2, 5, 4, 0x81]),               # GT: 0|1 and haploid 1 (end-of-vector padded)
              b"\x11\x07" + b"\x32" + struct.pack("<hhhhhh", 0, 30, 300, 255, -32768, 0),   # PL, one missing
              b"\x11\x09" + b"\x35" + struct.pack("<fff", 0.98, 0.01, 0.01) + struct.pack("<III", 0x7F800001, 0x7F800002, 0x7F800002)])
    note = "a-string-of-more-than-fifteen-characters"
    r2 = rec(1, 1999999, "", ["AT", "A", "ATT"], None, [0, 1],
             [b"\x11\x02" + b"\x13" + struct.pack("<i", 70000),             # DP as int32
              b"\x11\x05" + tstr(note),
              b"\x11\x04" + b"\x25" + struct.pack("<f", 0.5) + struct.pack("<I", 0x7F800001)],   # AF=0.5,.
             [b"\x11\x06" + b"\x21" + bytes([0, 2, 6, 4])])                 # GT: ./0 and 2/1
    r3 = rec(0, 5, "x", ["C"], 0.0, [], [], [], n_sample=0)               # no ALT, no FILTER, no INFO, sites only
    data = b"BCF\2\2" + struct.pack("<I", len(hdr)) + hdr + r1 + r2 + r3
    path = str(tmp_path / "t.bcf")
    with open(path, "wb") as f:
        for i in range(0, len(data), 100):     # tiny blocks: records and even length fields straddle block borders
            f.write(_bgzf_block(data[i:i + 100]))
        f.write(_bgzf_block(b""))
    r = subprocess.run([CRAMORE, "vcf-text", path], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = r.stdout.splitlines()
    assert lines[:len(text.splitlines())] == text.splitlines()
    body = lines[len(text.splitlines()):]
    assert body == [
        "chrA\t100\trs1\tA\tG\t30.5\tPASS\tDP=1000;DB;AF=0.25\tGT:PL:GP\t0|1:0,30,300:0.980000019,0.00999999978,0.00999999978\t1:255,.,0:.",
        "chrB\t2000000\t.\tAT\tA,ATT\t.\tPASS;q10\tDP=70000;NOTE=%s;AF=0.5,.\tGT\t./0\t2/1" % note,
        "chrA\t6\tx\tC\t.\t0\t.\t.",
    ]


def _bam_to_sam(bam, sam, gz=False):
    """SAM text of a BAM written by the test encoder (decoded by the oracle's Python BAM parser)."""
    import gzip
    from oracle import dsc_pileup_py
    refs, recs = dsc_pileup_py.read_bam(bam)
    opener = gzip.open if gz else open
    with opener(sam, "wt") as f:
        f.write("@HD\tVN:1.6\tSO:coordinate\n" + "".join("@SQ\tSN:%s\tLN:100000\n" % n for n in refs) + "@PG\tID:test\n")
        for i, r in enumerate(recs):
            cigar = "".join("%d%s" % (n, op) for op, n in r["cigar"]) or "*"
            qual = "".join(chr(q + 33) for q in r["qual"]) or "*"
            tags = "".join("\t%s:Z:%s" % kv for kv in r["tags"].items()) + "\tNM:i:%d\tXS:A:+\tXF:f:0.5" % (i % 3)
            f.write("r%d\t%d\t%s\t%d\t%d\t%s\t*\t0\t0\t%s\t%s%s\n" % (
                i, r["flag"], refs[r["tid"]] if r["tid"] >= 0 else "*", r["pos"] + 1, r["mapq"], cigar, r["seq"] or "*", qual, tags))


@pytest.mark.parametrize("gz", [False, True])
def test_sam_text_input_equals_bam_input(built, tmp_path, gz):
    """--sam takes SAM text (plain or gzip) as well: every alignment line is re-encoded into the BAM record layout,
    so dsc-pileup writes the same four files from either form of the same alignments."""
    pre, _ = _random_alignments(tmp_path, 21)
    sam = pre + (".sam.gz" if gz else ".sam")
    _bam_to_sam(pre + ".bam", sam, gz)
    outs = []
    for src in (pre + ".bam", sam):
        out = str(tmp_path / ("o%d" % len(outs)))
        r = subprocess.run([CRAMORE, "dsc-pileup", "--sam", src, "--vcf", pre + ".vcf", "--out", out], capture_output=True, text=True)
        assert r.returncode == 0, r.stderr[-2000:]
        outs.append([_gunzip(out + "." + ext + ".gz") for ext in ("cel", "var", "plp", "umi")])
    assert outs[0] == outs[1] and len(outs[0][2].splitlines()) > 150
    open(pre + ".cram", "wb").write(b"CRAM\3\0" + b"\0" * 40)
    r = subprocess.run([CRAMORE, "dsc-pileup", "--sam", pre + ".cram", "--vcf", pre + ".vcf", "--out", str(tmp_path / "c")],
                       capture_output=True, text=True)
    assert r.returncode != 0 and "CRAM" in r.stderr


def test_binary_readers_survive_damaged_files(built, tmp_path):
    """Truncated and bit-flipped BAM / BCF payloads (re-wrapped as valid BGZF, so the damage reaches the record
    parsers instead of the CRC check): the readers either finish or stop with a FATAL ERROR -- never a signal."""
    import gzip
    from cramore_b200.synth import _bgzf_block, make_dataset, write_bam, write_bcf, write_vcf
    d = make_dataset("C1", nbcs=6, nsnps=40, rbar=30, nv=3, seed=2, with_barcodes=True)
    pre = str(tmp_path / "in")
    write_vcf(d, pre + ".vcf")
    write_bcf(d, pre + ".bcf", with_r2=True)
    write_bam(d, pre + ".bam")
    rng = np.random.default_rng(7)

    def wrap(data, path):
        with open(path, "wb") as f:
            for i in range(0, len(data), 4000):
                f.write(_bgzf_block(data[i:i + 4000]))
            f.write(_bgzf_block(b""))

    bad = str(tmp_path / "bad")
    for kind, cmd in (("bcf", lambda p: [CRAMORE, "vcf-text", p]),
                      ("bam", lambda p: [CRAMORE, "dsc-pileup", "--sam", p, "--vcf", pre + ".vcf", "--out", str(tmp_path / "o"),
                                         "--sm", d["sample_ids"][0]])):
        raw = gzip.open(pre + "." + kind, "rb").read()
        seen = set()
        for trial in range(int(os.environ.get("CRAMORE_FUZZ_TRIALS", "60"))):
            data = bytearray(raw)
            if trial % 3 == 0:
                data = data[:int(rng.integers(1, len(data)))]
            else:
                for _ in range(int(rng.integers(1, 6))):
                    data[int(rng.integers(0, len(data)))] = int(rng.integers(0, 256))
            wrap(bytes(data), bad + "." + kind)
            r = subprocess.run(cmd(bad + "." + kind), capture_output=True, timeout=60)
            seen.add(r.returncode)
            assert r.returncode in (0, 1, 134), (kind, trial, r.returncode, r.stderr[-500:])
        assert 134 in seen or 1 in seen   # some of the damage was noticed


def test_patched_reference_has_no_cpu_path(tmp_path):
    """oracle/_ref/ref_cramore_gpu (the reference's cmdCramDemuxlet spliced onto the engine) on a box without a GPU:
    it loads the inputs with the reference's own code and then stops with the reference's FATAL ERROR -- there is
    nothing to fall back to."""
    import torch
    exe = os.path.join(ROOT, "oracle", "_ref", "ref_cramore_gpu")
    if not os.path.exists(exe):
        pytest.skip("oracle/_ref/ref_cramore_gpu not built")
    if torch.cuda.is_available():
        pytest.skip("a GPU is present: covered by tests/test_cli_gpu.py")
    from cramore_b200.synth import make_dataset, write_plp, write_vcf
    d = make_dataset("C1", nbcs=5, nsnps=40, rbar=20, nv=3, with_barcodes=True)
    write_plp(d, str(tmp_path / "in"))
    write_vcf(d, str(tmp_path / "in.vcf"))
    r = subprocess.run([exe, "demuxlet", "--plp", "in", "--vcf", "in.vcf", "--out", "out"], capture_output=True, text=True,
                       cwd=str(tmp_path))
    assert r.returncode != 0 and "ccu_create" in r.stderr and not os.path.exists(str(tmp_path / "out.best"))


def test_bgzf_stream_across_many_batches(built, tmp_path):
    """The BGZF reader hands out batches of 512 blocks, the next one inflated by a helper thread while the current one
    is consumed: a text of several thousand tiny blocks (lines straddling block and batch borders) comes back intact,
    with one inflate thread and with many."""
    from cramore_b200.synth import _bgzf_block
    rng = np.random.default_rng(5)
    lines = ["##fileformat=VCFv4.2", "#CHROM\tPOS\tID\tREF\tALT\tQUAL\tFILTER\tINFO"]
    for i in range(20000):
        lines.append("1\t%d\t.\tA\tC\t.\tPASS\tX=%s" % (i + 1, "".join("ACGT"[j] for j in rng.integers(0, 4, int(rng.integers(0, 60))))))
    text = ("\n".join(lines) + "\n").encode()
    path = str(tmp_path / "many.vcf.gz")
    with open(path, "wb") as f:
        off = 0
        while off < len(text):
            n = int(rng.integers(1, 400))
            f.write(_bgzf_block(text[off:off + n]))
            if rng.integers(50) == 0:
                f.write(_bgzf_block(b""))   # empty blocks may appear anywhere (bgzip writes one at every flush)
            off += n
        f.write(_bgzf_block(b""))
    for threads in ("1", "7"):
        r = subprocess.run([CRAMORE, "vcf-text", path], capture_output=True, env=dict(os.environ, CRAMORE_BGZF_THREADS=threads))
        assert r.returncode == 0, r.stderr[-1000:]
        assert r.stdout == text


REF_GPU_BIN_DUMP = os.path.join(ROOT, "oracle", "_ref", "ref_cramore_gpu")


@pytest.mark.skipif(not os.path.exists(REF_GPU_BIN_DUMP), reason="oracle/_ref/ref_cramore_gpu not built (make -C oracle ref_gpu)")
@pytest.mark.parametrize("case", ["plain", "filters", "gt_r2"])
def test_own_loader_equals_the_references_own_maps(built, tmp_path, case):
    """CSR bit-exactness against the REFERENCE's data structures, not a re-derivation: the reference's own
    load_from_plp / add_cell / add_snp / add_umi (sc_drop_seq.cpp:27-91, 103-384, compiled unmodified into
    oracle/_ref/ref_cramore_gpu) fill sc_dropseq_lib_t, the patch of INTEGRATION.md section 1 flattens those maps, and
    CCU_PATCH_DUMP_INPUTS writes the result before any GPU call.  `cramore demuxlet --gpu-dump-inputs` (this repo's
    loader, host_io.h) must produce the same bytes: droplet order, SNP order, read order inside an entry, alleles,
    capped qualities, INT_IDs -- and the same genotype doubles once its float posteriors are mixed."""
    from cramore_b200.synth import make_dataset, write_plp, write_vcf
    from oracle import pyoracle
    d = make_dataset("C1", nbcs=90, nsnps=140, rbar=170, nv=5, seed=77, with_barcodes=True)  # many multi-read entries
    pre = str(tmp_path / "in")
    write_plp(d, pre)
    write_vcf(d, pre + ".vcf", with_r2=True)
    common = ["--plp", "in", "--vcf", "in.vcf", "--out", "o"]
    if case == "filters":
        (tmp_path / "grp.txt").write_text("\n".join(b for i, b in enumerate(d["barcodes"]) if i % 4 != 2) + "\n")
        common += ["--group-list", "grp.txt", "--min-snp", "95", "--min-umi", "150", "--cap-BQ", "17", "--min-BQ", "14",
                   "--geno-error-offset", "0.05", "--geno-error-coeff", "0.5"]
    if case == "gt_r2":
        common += ["--field", "GT", "--geno-error-offset", "0.02", "--geno-error-coeff", "0.3", "--min-mac", "2"]
    r = subprocess.run([REF_GPU_BIN_DUMP, "demuxlet"] + common, capture_output=True, text=True, cwd=str(tmp_path),
                       env=dict(os.environ, CCU_PATCH_DUMP_INPUTS="ref"))
    assert r.returncode == 0, r.stderr[-2000:]
    r = subprocess.run([CRAMORE, "demuxlet"] + common + ["--gpu-dump-inputs", "mine"], capture_output=True, text=True,
                       cwd=str(tmp_path))
    assert r.returncode == 0, r.stderr[-2000:]
    for name in ("cell_ptr.i64", "snp_idx.i32", "read_ptr.i64", "al.u8", "bq.u8", "has_gp.u8", "barcodes.txt",
                 "row_int_id.i32"):
        a, b = (tmp_path / ("ref." + name)).read_bytes(), (tmp_path / ("mine." + name)).read_bytes()
        assert len(a) > 0 and a == b, name
    # genotypes: the reference holds mixed doubles (sc_drop_seq.cpp:287-315), this build float posteriors + per-SNP
    # error that the engine mixes on the device with the same operations (tests/test_demux_gpu.py pins the kernel to
    # this very oracle function)
    has = np.fromfile(str(tmp_path / "mine.has_gp.u8"), dtype="u1")
    gp = np.fromfile(str(tmp_path / "mine.gp.f32"), dtype="<f4").reshape(len(has), -1, 3)
    err = np.fromfile(str(tmp_path / "mine.snp_err.f64"), dtype="<f8")
    mixed = pyoracle.mix_genotypes(gp, err, has)
    ref = np.fromfile(str(tmp_path / "ref.gps.f64"), dtype="<f8").reshape(mixed.shape)
    assert has.sum() > 20 and np.array_equal(mixed[has == 1], ref[has == 1])


def test_cli_rejects_unusable_alpha_grids_before_reading_anything(built, tmp_path):
    """The engine refuses grids the reference is undefined for (no doublet fraction, first value not 0) and grids
    beyond CCU_MAX_ALPHA; the command line says so, naming the option, before it opens any input file."""
    base = [CRAMORE, "demuxlet", "--plp", str(tmp_path / "absent"), "--vcf", str(tmp_path / "absent.vcf"), "--out",
            str(tmp_path / "o")]
    for extra, word in ((["--alpha", "0.1", "--alpha", "0.5"], "first value must be 0"),
                        (["--alpha", "0"], "at least one doublet fraction"),
                        (sum((["--alpha", repr(i / 64.0)] for i in range(40)), []), "at most 32")):
        r = subprocess.run(base + extra, capture_output=True, text=True)
        assert r.returncode != 0 and "--alpha" in r.stderr and word in r.stderr, r.stderr[-500:]
        assert "Cannot open" not in r.stderr  # refused before any attempt to open the inputs


def test_best_row_reports_a_short_buffer(built):
    import ctypes as C
    from cramore_b200 import RESULT_DTYPE, load_library
    lib = load_library()
    res = np.zeros(1, RESULT_DTYPE)
    res["sNext"] = 1
    res["dBest2"] = 1
    res["dBestA"] = 1
    a = np.asarray([0.0, 0.5])
    ids = (C.c_char_p * 2)(b"S" * 300, b"T" * 300)
    small = C.create_string_buffer(256)
    n = lib.ccu_demux_best_row(small, C.c_int32(256), C.c_int32(0), b"BC", C.c_uint32(1), C.c_int32(1),
                               C.c_void_p(res.ctypes.data), C.c_int32(2), C.c_int32(2), C.c_void_p(a.ctypes.data),
                               C.c_double(0.5), ids)
    assert n == -2
    big = C.create_string_buffer(8192)
    n = lib.ccu_demux_best_row(big, C.c_int32(8192), C.c_int32(0), b"BC", C.c_uint32(1), C.c_int32(1),
                               C.c_void_p(res.ctypes.data), C.c_int32(2), C.c_int32(2), C.c_void_p(a.ctypes.data),
                               C.c_double(0.5), ids)
    assert n > 2000 and big.raw[n - 1:n] == b"\n"
