"""CPU tests: the oracle against outputs of the REFERENCE's own compiled engine
(tests/golden/ref_*.json, made by tests/golden/make_ref_fixtures.py from
/root/reference/cmd_cram_demuxlet.cpp / cmd_cram_freemuxlet.cpp / sc_drop_seq.cpp compiled unmodified
behind the zlib I/O shim).  These pin the restatement in oracle/ to the reference itself."""
import numpy as np
import pytest

from oracle import pyoracle
from tests import refcase

TYPES = {"SNG": 0, "DBL": 1, "AMB": 2}
TIGHT = 1e-12


@pytest.mark.parametrize("name", refcase.DEMUX_CASES)
def test_oracle_demuxlet_equals_reference(name):
    fx = refcase.load(name)
    d, alphas, prior, snp_err, has_gp = refcase.demux_inputs(fx)
    gps = pyoracle.mix_genotypes(d["gp"], snp_err, has_gp)
    res, _, _ = pyoracle.demux_score(d["cell_ptr"], d["snp_idx"], d["read_ptr"], d["al"], d["bq"], gps, alphas, prior,
                                     has_gp=has_gp)
    calls = pyoracle.demux_classify(res, d["nv"], alphas, prior)
    ref = refcase.best_rows(fx)
    assert len(ref) == d["nbcs"]
    order = sorted(range(d["nbcs"]), key=lambda i: d["barcodes"][i])  # rows come in std::map barcode order (:636)
    ids = d["sample_ids"]
    cp, rp = d["cell_ptr"], d["read_ptr"]
    for rank, c in enumerate(order):
        r = ref[d["barcodes"][c]]
        assert r["int_id"] == rank
        assert r["nsnps"] == cp[c + 1] - cp[c] and r["nreads"] == rp[cp[c + 1]] - rp[cp[c]]
        assert TYPES[r["type"]] == calls["type"][c]
        # exact equality of every likelihood the reference printed: same operations in the same order
        for mine, theirs in ((res["sngBestLLK"][c], r["sngBestLLK"]), (res["sngNextLLK"][c], r["sngNextLLK"]),
                             (res["dblBestLLK"][c], r["dblBestLLK"]), (calls["bestLLK"][c], r["bestLLK"]),
                             (calls["nextLLK"][c], r["nextLLK"])):
            assert mine == theirs or abs(mine - theirs) <= TIGHT * abs(theirs), (name, c, mine, theirs)
        for mine, theirs in ((calls["bestPP"][c], r["bestPP"]), (calls["sngPP"][c], r["sngPP"]),
                             (calls["sngOnlyPP"][c], r["sngOnlyPP"])):
            assert abs(mine - theirs) <= 1e-10 * max(1.0, abs(theirs)), (name, c, mine, theirs)
        assert ids[res["sBest"][c]] == r["sBest"] and ids[res["sNext"][c]] == r["sNext"]
        assert (ids[res["dBest1"][c]], ids[res["dBest2"][c]]) == (r["dBest1"], r["dBest2"])
        assert alphas[res["dBestA"][c]] == r["dBestAlpha"]
        assert (ids[calls["jBest"][c]], ids[calls["kBest"][c]], alphas[calls["alphaBest"][c]]) == \
            (r["best1"], r["best2"], r["bestA"])
        assert (ids[calls["jNext"][c]], ids[calls["kNext"][c]], alphas[calls["alphaNext"][c]]) == \
            (r["next1"], r["next2"], r["nextA"])
    # and the text rows are byte-identical to the reference's .best
    lines = fx["best_text"].splitlines(keepends=True)
    for rank, c in enumerate(order):
        mine = pyoracle.demux_best_row(rank, d["barcodes"][c], int(cp[c + 1] - cp[c]), int(rp[cp[c + 1]] - rp[cp[c]]),
                                       res[c], d["nv"], alphas, prior, ids)
        assert mine == lines[1 + rank]


@pytest.mark.parametrize("name", refcase.FMX_CASES)
def test_oracle_freemuxlet_equals_reference(name):
    fx = refcase.load(name)
    d = refcase.dataset(fx)
    K = fx["spec"]["K"]
    args = fx["spec"]["args"]
    ge = float(args[args.index("--geno-error") + 1]) if "--geno-error" in args else 0.0
    prior = float(args[args.index("--doublet-prior") + 1]) if "--doublet-prior" in args else 0.5
    # freemuxlet's own read filter defaults: --cap-BQ 40 --min-BQ 13 (cmd_cram_freemuxlet.cpp:27-28); synthetic
    # base qualities are already capped at 20, so the CSR is what load_from_plp builds
    plp = pyoracle.fmx_pileup(d["read_ptr"], d["al"], d["bq"], 0.5)
    af = np.asarray(["%.5f" % a for a in d["af"]], np.float64)  # AF as it round-trips through .var.gz
    l0, l2 = pyoracle.fmx_lmix(d["cell_ptr"], d["snp_idx"], plp, af)
    lm = [r for r in fx["lmix_raw"] if len(r) == 8]
    assert len(lm) == d["nbcs"]
    for c, r in enumerate(lm):
        assert abs(l0[c] - float(r[4])) <= TIGHT * abs(float(r[4])) and abs(l2[c] - float(r[5])) <= TIGHT * abs(float(r[5]))
    clust = pyoracle.fmx_mstep(d["cell_ptr"], d["snp_idx"], plp, np.asarray(fx["init"], np.int32), K, d["nsnps"])
    for it in range(10):  # cmd_cram_freemuxlet.cpp:456-457
        res, _ = pyoracle.fmx_estep(d["cell_ptr"], d["snp_idx"], plp, af, clust, prior, ge, it + 1 == 10)
        assign = np.where(res["type"] == 0, res["jBest"], -1).astype(np.int32)
        clust = pyoracle.fmx_mstep(d["cell_ptr"], d["snp_idx"], plp, assign, K, d["nsnps"])
    rows = refcase.fmx_rows(fx)
    assert len(rows) == d["nbcs"]
    for c, r in enumerate(rows):
        assert TYPES[r["type"]] == res["type"][c], (c, r["type"], res["type"][c])
        for k in ("jBest", "kBest", "jNext", "kNext", "sBest", "sNext", "dBest1", "dBest2"):
            assert r[k] == res[k][c], (c, k)
        for k in ("bestLLK", "nextLLK", "sngBestLLK", "sngNextLLK", "dblBestLLK", "bestPP"):
            assert abs(res[k][c] - r[k]) <= TIGHT * abs(r[k]), (c, k, res[k][c], r[k])
        for k in ("sngPP", "sngOnlyPP"):
            assert abs(res[k][c] - r[k]) <= 1e-10, (c, k)
    # final cluster pileups as printed into clust1.vcf.gz (:670-704): DP, AD, PL and GP of every cluster
    assert len(fx["vcf_sites"]) >= 40
    for site in fx["vcf_sites"]:
        s = (site["pos"] - 1000) // 10
        gps = [(1. - af[s]) * (1. - af[s]), 2. * af[s] * (1. - af[s]), af[s] * af[s]]
        for k in range(K):
            gt1, gt2, gq, dp, nref, nalt, pl0, pl1, pl2, pp0, pp1, pp2 = site["clusters"][k]
            assert (dp, nref, nalt) == (clust["nreads"][k, s], clust["nref"][k, s], clust["nalt"][k, s])
            g = clust["gls"][k, s][[0, 4, 8]]
            mx = g.max()
            pp = np.asarray([gps[i] * (g[i] / mx) + 1e-100 for i in range(3)])
            pp /= pp.sum()
            np.testing.assert_allclose(pp, [pp0, pp1, pp2], rtol=1e-11, atol=1e-300)
            assert [int(-10.0 * np.log10(g[i] / mx)) for i in range(3)] == [pl0, pl1, pl2]


@pytest.mark.parametrize("name", refcase.FMX_INIT_CASES)
def test_initial_clustering_equals_reference(name, built):
    """cmd_cram_freemuxlet.cpp:166-346 against the reference's own run (no --init-cluster, or --init-cluster plus
    the ten sweeps): the oracle's pairwise distances equal every row of .ldist.gz, and the host clustering
    (ccu_fmx_initial_clusters: same rand() call sequence, same vote order) reproduces .clust0 exactly."""
    from cramore_b200.engine import initial_clusters
    fx = refcase.load(name)
    d, plp, af, prm = refcase.fmx_init_inputs(fx)
    K = fx["spec"]["K"]
    l0, l2 = pyoracle.fmx_lmix(d["cell_ptr"], d["snp_idx"], plp, af)
    dense = pyoracle.fmx_pair_dist(d["cell_ptr"], d["snp_idx"], plp, af, d["nsnps"])
    for r in fx["ldist_raw"]:
        a, b = int(r[0]), int(r[1])
        dd = dense[max(a, b), min(a, b)]
        assert (dd["nsnps"], dd["nread1"], dd["nread2"]) == (int(r[2]), int(r[3]), int(r[4])), (a, b)
        for mine, theirs in ((dd["llk0"], float(r[6])), (dd["llk2"], float(r[7]))):
            assert mine == theirs or abs(mine - theirs) <= TIGHT * abs(theirs), (a, b, mine, theirs)
    if fx["init"] is None:
        assert len(fx["ldist_raw"]) > 500
    init = None if fx["init"] is None else np.asarray(fx["init"], np.int32)
    clusts = initial_clusters(l2 - l0, refcase.sparse_pairs(dense), K, bf_thres=prm["bf_thres"],
                              frac_init_clust=prm["frac_init_clust"], iter_init=10,
                              keep_init_missing=prm["keep_init_missing"], init=init, seed=1)
    assert clusts.tolist() == fx["clust0"]
    # and the EM that follows, started from those clusters, ends where the reference ended
    clust = pyoracle.fmx_mstep(d["cell_ptr"], d["snp_idx"], plp, clusts, K, d["nsnps"])
    for it in range(10):
        res, _ = pyoracle.fmx_estep(d["cell_ptr"], d["snp_idx"], plp, af, clust, 0.5, prm["geno_error"], it + 1 == 10)
        assign = np.where(res["type"] == 0, res["jBest"], -1).astype(np.int32)
        clust = pyoracle.fmx_mstep(d["cell_ptr"], d["snp_idx"], plp, assign, K, d["nsnps"])
    rows = refcase.fmx_rows(fx)
    assert len(rows) == d["nbcs"]
    for c, r in enumerate(rows):
        assert TYPES[r["type"]] == res["type"][c] and (r["jBest"], r["kBest"]) == (res["jBest"][c], res["kBest"][c])
        assert abs(res["bestLLK"][c] - r["bestLLK"]) <= TIGHT * abs(r["bestLLK"])


@pytest.mark.parametrize("name", refcase.FMX2_CASES)
def test_freemux2_equals_reference(name, built):
    """cmd_cram_freemux2.cpp against the reference's own run: the greedy cluster-pileup initialisation
    (ccu_fmx_greedy_clusters, :217-261) reproduces .clust0 exactly, and the oracle's EM in the freemux2 flavour
    (geno_error on every iteration, stop once no call changes, :374-605) ends in the reference's rows."""
    from cramore_b200.engine import greedy_clusters
    fx = refcase.load(name)
    d, plp, af, prm = refcase.fmx_init_inputs(fx)
    K = fx["spec"]["K"]
    l0, l2 = pyoracle.fmx_lmix(d["cell_ptr"], d["snp_idx"], plp, af)
    if fx["init"] is None:
        clusts = greedy_clusters(d["cell_ptr"], d["snp_idx"], plp, af, l2 - l0, K, d["nsnps"],
                                 frac_init_clust=prm["frac_init_clust"])
    else:
        clusts = np.asarray(fx["init"], np.int32)
    assert clusts.tolist() == fx["clust0"]
    clust = pyoracle.fmx_mstep(d["cell_ptr"], d["snp_idx"], plp, clusts, K, d["nsnps"])
    prev = None
    for it in range(10):
        res, _ = pyoracle.fmx_estep(d["cell_ptr"], d["snp_idx"], plp, af, clust, prm["doublet_prior"], prm["geno_error"], True)
        assign = np.where(res["type"] == 0, res["jBest"], -1).astype(np.int32)
        clust = pyoracle.fmx_mstep(d["cell_ptr"], d["snp_idx"], plp, assign, K, d["nsnps"])
        key = np.where(res["type"] == 0, res["sBest"], -1 - res["type"])
        if prev is not None and np.array_equal(key, prev):
            break
        prev = key
    rows = refcase.fmx_rows(fx)
    assert len(rows) == d["nbcs"]
    for c, r in enumerate(rows):
        assert TYPES[r["type"]] == res["type"][c], (c, r["type"], res["type"][c])
        assert (r["jBest"], r["kBest"]) == (res["jBest"][c], res["kBest"][c])
        assert abs(res["bestLLK"][c] - r["bestLLK"]) <= TIGHT * abs(r["bestLLK"]), (c, res["bestLLK"][c], r["bestLLK"])
