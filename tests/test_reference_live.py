"""CPU tests against the LIVE reference: oracle/_ref/ref_cramore -- the reference's own cmd_cram_demuxlet.cpp /
cmd_cram_freemuxlet.cpp / sc_drop_seq.cpp compiled unmodified behind the zlib I/O shim -- is run on freshly drawn
random cases (sizes, sample counts, alpha grids, priors, error settings) and the oracle has to reproduce its output
files: .best byte for byte, .clust1.samples.gz rows to the last digit.  Complements the committed fixtures
(tests/golden/ref_*.json), which cover a fixed set of cases everywhere; this sweep runs wherever the reference
binary has been built (here and on the GPU box, where oracle/_ref travels) and is skipped otherwise."""
import gzip
import os
import subprocess

import numpy as np
import pytest

from cramore_b200.synth import make_dataset, write_plp, write_vcf
from oracle import pyoracle

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.path.join(ROOT, "oracle", "_ref", "ref_cramore")
pytestmark = pytest.mark.skipif(not os.path.exists(REF), reason="oracle/_ref/ref_cramore not built (make -C oracle ref)")


def _run_ref(args, cwd):
    r = subprocess.run([REF] + args, capture_output=True, text=True, cwd=cwd)
    assert r.returncode == 0, r.stderr[-3000:]


def _complement_free_grid(rng):
    """0 first, then distinct values in (0, 1) none of which is the complement of another;
    0.5 itself may appear."""
    pool = [0.05, 0.1, 0.15, 0.2, 0.25, 0.3, 0.35, 0.4, 0.45, 0.5, 0.6, 0.65, 0.75, 0.8, 0.9]
    chosen = []
    for a in rng.permutation(pool)[:int(rng.integers(1, 6))]:
        if all(abs(a + b - 1.0) > 1e-9 or a == 0.5 for b in chosen):
            chosen.append(float(a))
    return [0.0] + sorted(chosen)


@pytest.mark.parametrize("seed", range(40))
def test_oracle_demuxlet_equals_live_reference(tmp_path, seed):
    rng = np.random.default_rng(1000 + seed)
    nv = int(rng.integers(2, 10))
    d = make_dataset("C1", nbcs=int(rng.integers(8, 40)), nsnps=int(rng.integers(60, 400)), rbar=int(rng.integers(10, 120)),
                     nv=nv, seed=5000 + seed, with_barcodes=True)
    alphas = _complement_free_grid(rng)
    prior = float(rng.choice([0.1, 0.3, 0.5, 0.8]))
    off = float(rng.choice([0.0, 0.01, 0.1, 0.3]))
    coeff = float(rng.choice([0.0, 0.0, 0.5]))
    field = str(rng.choice(["GP", "GP", "GT"]))
    pre = str(tmp_path / "in")
    write_plp(d, pre)
    write_vcf(d, pre + ".vcf", with_r2=True)
    args = ["demuxlet", "--plp", "in", "--vcf", "in.vcf", "--field", field, "--out", "out", "--doublet-prior", repr(prior),
            "--geno-error-offset", repr(off), "--geno-error-coeff", repr(coeff)]
    for a in alphas:
        args += ["--alpha", repr(a)]
    _run_ref(args, str(tmp_path))
    ref_lines = open(str(tmp_path / "out.best")).read().splitlines(keepends=True)

    gp = np.asarray(d["gp"])
    if field == "GT":
        gp = np.zeros((d["nsnps"], nv, 3), np.float32)
        np.put_along_axis(gp, np.asarray(d["geno"], np.int64)[:, :, None], np.float32(1.0), axis=2)
    err = off + ((1 - off) * (1 - float(np.float32(0.9))) * coeff if coeff > 0 else 0.0)
    ac = np.asarray(d["geno"], np.int64).sum(axis=1)
    has_gp = ((ac >= 1) & (2 * nv - ac >= 1)).astype(np.uint8)  # --min-mac 1 (cmd_cram_demuxlet.cpp:27)
    al = np.asarray(alphas, np.float64)
    gps = pyoracle.mix_genotypes(gp, np.full(d["nsnps"], err), has_gp)
    res, _, _ = pyoracle.demux_score(d["cell_ptr"], d["snp_idx"], d["read_ptr"], d["al"], d["bq"], gps, al, prior,
                                     has_gp=has_gp)
    order = sorted(range(d["nbcs"]), key=lambda i: d["barcodes"][i])
    cp, rp = d["cell_ptr"], d["read_ptr"]
    assert len(ref_lines) == 1 + d["nbcs"]
    for rank, c in enumerate(order):
        mine = pyoracle.demux_best_row(rank, d["barcodes"][c], int(cp[c + 1] - cp[c]), int(rp[cp[c + 1]] - rp[cp[c]]),
                                       res[c], nv, al, prior, d["sample_ids"])
        assert mine == ref_lines[1 + rank], (seed, field, alphas, prior, off, coeff)


@pytest.mark.parametrize("seed", range(16))
def test_oracle_freemuxlet_equals_live_reference(tmp_path, seed):
    rng = np.random.default_rng(2000 + seed)
    K = int(rng.integers(2, 8))
    d = make_dataset("C4", nbcs=int(rng.integers(30, 90)), nsnps=int(rng.integers(200, 600)), rbar=int(rng.integers(40, 110)),
                     nv=int(rng.integers(2, 6)), seed=6000 + seed, with_barcodes=True)
    ge = float(rng.choice([0.0, 0.02, 0.1]))
    prior = float(rng.choice([0.2, 0.5, 0.7]))
    pre = str(tmp_path / "in")
    bcs = write_plp(d, pre)
    init = (d["s1"] % K).astype(np.int32)
    init[d["is_dbl"]] = -1
    init[rng.random(len(init)) < 0.3] = -1
    with open(str(tmp_path / "init.txt"), "w") as f:
        for b, c in zip(bcs, init):
            f.write("%s\t%d\n" % (b, c))
    _run_ref(["freemuxlet", "--plp", "in", "--out", "out", "--nsample", str(K), "--init-cluster", "init.txt", "--iter-init", "0",
              "--keep-init-missing", "--geno-error", repr(ge), "--doublet-prior", repr(prior)], str(tmp_path))
    ref_rows = gzip.open(str(tmp_path / "out.clust1.samples.gz"), "rt").read().splitlines()[1:]

    plp = pyoracle.fmx_pileup(d["read_ptr"], d["al"], d["bq"], 0.5)
    af = np.asarray(["%.5f" % a for a in d["af"]], np.float64)
    clust = pyoracle.fmx_mstep(d["cell_ptr"], d["snp_idx"], plp, init, K, d["nsnps"])
    for it in range(10):
        res, _ = pyoracle.fmx_estep(d["cell_ptr"], d["snp_idx"], plp, af, clust, prior, ge, it + 1 == 10)
        assign = np.where(res["type"] == 0, res["jBest"], -1).astype(np.int32)
        clust = pyoracle.fmx_mstep(d["cell_ptr"], d["snp_idx"], plp, assign, K, d["nsnps"])
    assert len(ref_rows) == d["nbcs"]
    names = {0: "SNG", 1: "DBL", 2: "AMB"}
    for c, line in enumerate(ref_rows):
        f = line.split("\t")
        # INT_ID BARCODE NUM.SNPS NUM.READS DROPLET.TYPE BEST.GUESS BEST.LLK NEXT.GUESS NEXT.LLK ... (:709-711)
        assert f[1] == bcs[c] and f[4] == names[int(res["type"][c])], (seed, c, line)
        assert f[5] == "%d,%d" % (res["jBest"][c], res["kBest"][c]) and f[7] == "%d,%d" % (res["jNext"][c], res["kNext"][c])
        assert f[6] == "%.2f" % res["bestLLK"][c] and f[8] == "%.2f" % res["nextLLK"][c], (seed, c, line)


@pytest.mark.parametrize("seed", range(10))
def test_freemux2_equals_live_reference(tmp_path, seed):
    """freemux2 (cmd_cram_freemux2.cpp) on random cases: the library's greedy cluster-pileup initialisation gives the
    reference's .clust0, and the oracle's EM in the freemux2 flavour (genotype error on every iteration, stop once no
    call changes) ends in the reference's .clust1.samples.gz rows."""
    from cramore_b200.engine import greedy_clusters
    rng = np.random.default_rng(5000 + seed)
    K = int(rng.integers(2, 7))
    d = make_dataset("C4", nbcs=int(rng.integers(30, 90)), nsnps=int(rng.integers(200, 500)), rbar=int(rng.integers(40, 110)),
                     nv=int(rng.integers(2, 6)), seed=9000 + seed, with_barcodes=True)
    ge = float(rng.choice([0.02, 0.1]))
    prior = float(rng.choice([0.3, 0.5]))
    frac = float(rng.choice([0.5, 1.0]))
    bcs = write_plp(d, str(tmp_path / "in"))
    _run_ref(["freemux2", "--plp", "in", "--out", "out", "--nsample", str(K), "--aux-files", "--seed", "11", "--geno-error", repr(ge),
              "--doublet-prior", repr(prior), "--frac-init-clust", repr(frac)], str(tmp_path))
    c0 = [int(ln.split("\t")[2]) for ln in gzip.open(str(tmp_path / "out.clust0.samples.gz"), "rt").read().splitlines()[1:]]
    ref_rows = gzip.open(str(tmp_path / "out.clust1.samples.gz"), "rt").read().splitlines()[1:]

    # freemux2's read filter defaults cap base qualities at 20 (cmd_cram_freemux2.cpp:24); the synthetic ones already are
    plp = pyoracle.fmx_pileup(d["read_ptr"], d["al"], d["bq"], 0.5)
    af = np.asarray(["%.5f" % a for a in d["af"]], np.float64)
    l0, l2 = pyoracle.fmx_lmix(d["cell_ptr"], d["snp_idx"], plp, af)
    clusts = greedy_clusters(d["cell_ptr"], d["snp_idx"], plp, af, l2 - l0, K, d["nsnps"], frac_init_clust=frac)
    assert clusts.tolist() == c0, seed
    clust = pyoracle.fmx_mstep(d["cell_ptr"], d["snp_idx"], plp, clusts, K, d["nsnps"])
    prev = None
    for it in range(10):
        res, _ = pyoracle.fmx_estep(d["cell_ptr"], d["snp_idx"], plp, af, clust, prior, ge, True)
        assign = np.where(res["type"] == 0, res["jBest"], -1).astype(np.int32)
        clust = pyoracle.fmx_mstep(d["cell_ptr"], d["snp_idx"], plp, assign, K, d["nsnps"])
        key = np.where(res["type"] == 0, res["sBest"], -1 - res["type"])
        if prev is not None and np.array_equal(key, prev):
            break
        prev = key
    names = {0: "SNG", 1: "DBL", 2: "AMB"}
    assert len(ref_rows) == d["nbcs"]
    for c, line in enumerate(ref_rows):
        f = line.split("\t")
        assert f[1] == bcs[c] and f[4] == names[int(res["type"][c])], (seed, c, line)
        assert f[5] == "%d,%d" % (res["jBest"][c], res["kBest"][c]) and f[6] == "%.2f" % res["bestLLK"][c], (seed, c, line)


@pytest.mark.parametrize("seed", range(8))
def test_initial_clustering_equals_live_reference(tmp_path, seed):
    """freemuxlet without --init-cluster on random cases: pairwise distances (oracle) + the library's host votes with
    the reference's rand() call order give the .clust0 the reference itself wrote, for random thresholds and fractions."""
    from cramore_b200.engine import initial_clusters
    from tests import refcase
    rng = np.random.default_rng(7000 + seed)
    K = int(rng.integers(2, 6))
    d = make_dataset("C4", nbcs=int(rng.integers(30, 70)), nsnps=int(rng.integers(150, 400)), rbar=int(rng.integers(60, 120)),
                     nv=int(rng.integers(2, 5)), seed=9500 + seed, with_barcodes=True)
    bf = float(rng.choice([2.0, 5.41, 8.0]))
    frac = float(rng.choice([0.6, 1.0]))
    write_plp(d, str(tmp_path / "in"))
    _run_ref(["freemuxlet", "--plp", "in", "--out", "out", "--nsample", str(K), "--aux-files", "--bf-thres", repr(bf),
              "--frac-init-clust", repr(frac)], str(tmp_path))
    c0 = [int(ln.split("\t")[2]) for ln in gzip.open(str(tmp_path / "out.clust0.samples.gz"), "rt").read().splitlines()[1:]]
    plp = pyoracle.fmx_pileup(d["read_ptr"], d["al"], d["bq"], 0.5)
    af = np.asarray(["%.5f" % a for a in d["af"]], np.float64)
    l0, l2 = pyoracle.fmx_lmix(d["cell_ptr"], d["snp_idx"], plp, af)
    dense = pyoracle.fmx_pair_dist(d["cell_ptr"], d["snp_idx"], plp, af, d["nsnps"])
    clusts = initial_clusters(l2 - l0, refcase.sparse_pairs(dense), K, bf_thres=bf, frac_init_clust=frac, iter_init=10,
                              keep_init_missing=False, init=None, seed=1)
    assert clusts.tolist() == c0, (seed, K, bf, frac)
