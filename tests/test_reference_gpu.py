"""-m gpu: the CUDA engine (through the C ABI) against outputs of the REFERENCE's own compiled engine
(tests/golden/ref_*.json; see tests/golden/make_ref_fixtures.py and oracle/ref_shim/)."""
import numpy as np
import pytest

from tests import refcase

pytestmark = pytest.mark.gpu
RTOL = 1e-9  # north_star: per-cell singlet/doublet log-likelihoods within 1e-9 relative
TYPES = {"SNG": 0, "DBL": 1, "AMB": 2}


@pytest.mark.parametrize("name", refcase.DEMUX_CASES)
def test_engine_demuxlet_equals_reference(engine, name):
    from cramore_b200 import best_row, classify
    from cramore_b200.engine import refine
    fx = refcase.load(name)
    d, alphas, prior, snp_err, has_gp = refcase.demux_inputs(fx)
    engine.configure(d["nv"], alphas, prior)
    engine.set_genotypes(d["gp"], snp_err, has_gp)
    res = engine.score(d["cell_ptr"], d["snp_idx"], d["read_ptr"], d["al"], d["bq"])
    calls = classify(res, d["nv"], alphas, prior)
    ref = refcase.best_rows(fx)
    ids = d["sample_ids"]
    cp, rp = d["cell_ptr"], d["read_ptr"]
    order = sorted(range(d["nbcs"]), key=lambda i: d["barcodes"][i])
    lines = fx["best_text"].splitlines(keepends=True)
    for rank, c in enumerate(order):
        r = ref[d["barcodes"][c]]
        assert TYPES[r["type"]] == calls["type"][c], (name, c)
        for mine, theirs in ((res["sngBestLLK"][c], r["sngBestLLK"]), (res["sngNextLLK"][c], r["sngNextLLK"]),
                             (res["dblBestLLK"][c], r["dblBestLLK"]), (calls["bestLLK"][c], r["bestLLK"]),
                             (calls["nextLLK"][c], r["nextLLK"])):
            assert abs(mine - theirs) <= RTOL * abs(theirs) + 1e-13, (name, c, mine, theirs)
        np.testing.assert_allclose([calls["sngPP"][c], calls["sngOnlyPP"][c]], [r["sngPP"], r["sngOnlyPP"]],
                                   rtol=1e-7, atol=1e-12)
        assert ids[res["sBest"][c]] == r["sBest"] and ids[res["sNext"][c]] == r["sNext"]
        # alpha = 0.5 orientations (j,k)/(k,j) are one hypothesis; the reference picks by last-bit noise (App. A.5)
        mine_d = (ids[res["dBest1"][c]], ids[res["dBest2"][c]])
        theirs_d = (r["dBest1"], r["dBest2"])
        assert alphas[res["dBestA"][c]] == r["dBestAlpha"]
        assert mine_d == theirs_d or (r["dBestAlpha"] == 0.5 and mine_d == theirs_d[::-1]), (name, c)
        assert ids[calls["jBest"][c]] in (r["best1"], r["best2"])
    # ---- the print: the named hypotheses re-scored in the reference's operation order (ccu_demux_refine) -> the
    # orientation of mirror pairs, every index and every printed byte are the reference's
    fixed = refine(res, d["cell_ptr"], d["snp_idx"], d["read_ptr"], d["al"], d["bq"], d["nv"], alphas, d["gp"], snp_err,
                   has_gp)
    for rank, c in enumerate(order):
        r = ref[d["barcodes"][c]]
        for mine, theirs in ((fixed["sngBestLLK"][c], r["sngBestLLK"]), (fixed["sngNextLLK"][c], r["sngNextLLK"]),
                             (fixed["dblBestLLK"][c], r["dblBestLLK"])):
            assert mine == theirs, (name, c, mine, theirs)  # the reference's own doubles (%.17g round trip)
        assert (ids[fixed["dBest1"][c]], ids[fixed["dBest2"][c]], alphas[fixed["dBestA"][c]]) == \
            (r["dBest1"], r["dBest2"], r["dBestAlpha"]), (name, c)
        row = best_row(rank, d["barcodes"][c], int(cp[c + 1] - cp[c]), int(rp[cp[c + 1]] - rp[cp[c]]), fixed[c], d["nv"],
                       alphas, prior, ids)
        assert row == lines[1 + rank], (name, c, row, lines[1 + rank])


@pytest.mark.parametrize("name", refcase.FMX_CASES)
def test_engine_freemuxlet_equals_reference(built, name):
    from cramore_b200 import FreemuxEngine
    fx = refcase.load(name)
    d = refcase.dataset(fx)
    K = fx["spec"]["K"]
    args = fx["spec"]["args"]
    ge = float(args[args.index("--geno-error") + 1]) if "--geno-error" in args else 0.0
    prior = float(args[args.index("--doublet-prior") + 1]) if "--doublet-prior" in args else 0.5
    af = np.asarray(["%.5f" % a for a in d["af"]], np.float64)
    eng = FreemuxEngine(0)
    try:
        eng.configure_fmx(K, prior, ge)
        eng.upload_fmx(d["cell_ptr"], d["snp_idx"], d["read_ptr"], d["al"], d["bq"], af)
        l0, l2 = eng.lmix()
        lm = [r for r in fx["lmix_raw"] if len(r) == 8]
        np.testing.assert_allclose(l0, [float(r[4]) for r in lm], rtol=RTOL)
        np.testing.assert_allclose(l2, [float(r[5]) for r in lm], rtol=RTOL)
        res, iters = eng.run_em(np.asarray(fx["init"], np.int32), niter=10)
        clust = eng.clusters()
    finally:
        eng.close()
    rows = refcase.fmx_rows(fx)
    for c, r in enumerate(rows):
        assert TYPES[r["type"]] == res["type"][c], (c,)
        for k in ("jBest", "kBest", "jNext", "kNext", "sBest", "sNext", "dBest1", "dBest2"):
            assert r[k] == res[k][c], (c, k)
        for k in ("bestLLK", "nextLLK", "sngBestLLK", "sngNextLLK", "dblBestLLK", "bestPP"):
            assert abs(res[k][c] - r[k]) <= RTOL * abs(r[k]) + 1e-12, (c, k, res[k][c], r[k])
        for k in ("sngPP", "sngOnlyPP"):
            assert abs(res[k][c] - r[k]) <= 1e-7, (c, k)
    for site in fx["vcf_sites"]:
        s = (site["pos"] - 1000) // 10
        gps = [(1. - af[s]) * (1. - af[s]), 2. * af[s] * (1. - af[s]), af[s] * af[s]]
        for k in range(K):
            gt1, gt2, gq, dp, nref, nalt, pl0, pl1, pl2, pp0, pp1, pp2 = site["clusters"][k]
            assert (dp, nref, nalt) == (clust["nreads"][k, s], clust["nref"][k, s], clust["nalt"][k, s])
            g = clust["gls"][k, s][[0, 4, 8]]
            pp = np.asarray([gps[i] * (g[i] / g.max()) + 1e-100 for i in range(3)])
            np.testing.assert_allclose(pp / pp.sum(), [pp0, pp1, pp2], rtol=1e-11, atol=1e-300)
