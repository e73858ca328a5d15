"""Shared comparison helpers: CUDA engine vs CPU oracle (tolerances of BASELINE.json north_star)."""
import numpy as np

from oracle import pyoracle

RTOL = 1e-9  # per-cell singlet/doublet log-likelihoods: 1e-9 relative (north_star)
ATOL = 1e-13  # floor for LLKs that are ~0 (droplets whose reads carry no information: |LLK| ~ 1e-8)


def oracle_run(d, doublet_prior=0.5, has_gp=None, want_llk=False, want_tiles=False, cells=None):
    gps = pyoracle.mix_genotypes(d["gp"], d["snp_err"], has_gp)
    b, e = (0, None) if cells is None else cells
    res, llk, tiles = pyoracle.demux_score(d["cell_ptr"], d["snp_idx"], d["read_ptr"], d["al"], d["bq"], gps,
                                           d["alphas"], doublet_prior, has_gp=has_gp, cell_begin=b, cell_end=e,
                                           want_llk=want_llk, want_tiles=want_tiles)
    return gps, res, llk, tiles


def engine_run(eng, d, doublet_prior=0.5, has_gp=None):
    eng.configure(d["nv"], d["alphas"], doublet_prior)
    eng.set_genotypes(d["gp"], d["snp_err"], has_gp)
    return eng.score(d["cell_ptr"], d["snp_idx"], d["read_ptr"], d["al"], d["bq"])


def _same_pair(a, b, alpha):
    """(j,k,n) equality; alpha == 0.5 orientations (j,k)/(k,j) are the same hypothesis and the
    reference itself picks between them by last-bit rounding."""
    if a == b:
        return True
    return a[2] == b[2] and a[2] >= 0 and alpha[a[2]] == 0.5 and (a[0], a[1]) == (b[1], b[0])


def assert_results_match(res, ref, alphas, ref_llk=None, rtol=RTOL):
    alphas = np.asarray(alphas)
    n = len(ref)
    assert len(res) == n
    for f in ("sngBestLLK", "sngNextLLK", "dblBestLLK", "dblNextLLK"):
        np.testing.assert_allclose(res[f], ref[f], rtol=rtol, atol=ATOL, err_msg=f)
    for f in ("sumLLK", "sngLLK"):  # consumed only through exp(x - sumLLK): absolute accuracy matters
        np.testing.assert_allclose(res[f], ref[f], rtol=rtol, atol=1e-9, err_msg=f)
    for i in range(n):
        for which in ("s",):
            for tag in ("Best", "Next"):
                a, b = int(res[which + tag][i]), int(ref[which + tag][i])
                if a != b:  # only acceptable if the two samples tie within tolerance in the oracle
                    assert ref_llk is not None, "singlet %s differs at cell %d: %d vs %d" % (tag, i, a, b)
                    va, vb = ref_llk[i, a, 0, 0], ref_llk[i, b, 0, 0]
                    assert abs(va - vb) <= rtol * abs(vb) + ATOL, (i, tag, a, b, va, vb)
        for tag in ("Best", "Next"):
            a = (int(res["d%s1" % tag][i]), int(res["d%s2" % tag][i]), int(res["d%sA" % tag][i]))
            b = (int(ref["d%s1" % tag][i]), int(ref["d%s2" % tag][i]), int(ref["d%sA" % tag][i]))
            if _same_pair(a, b, alphas):
                continue
            assert ref_llk is not None, "doublet %s differs at cell %d: %s vs %s" % (tag, i, a, b)
            va, vb = ref_llk[i, a[0], a[1], a[2]], ref_llk[i, b[0], b[1], b[2]]
            assert abs(va - vb) <= rtol * abs(vb) + ATOL, (i, tag, a, b, va, vb)


def assert_calls_match(res, ref, nv, alphas, doublet_prior=0.5):
    """DROPLET.TYPE and BEST/NEXT guesses from the product's host classifier vs the oracle's."""
    from cramore_b200 import classify
    mine = classify(res, nv, alphas, doublet_prior)
    theirs = pyoracle.demux_classify(ref, nv, alphas, doublet_prior)
    alphas = np.asarray(alphas)
    assert np.array_equal(mine["type"], theirs["type"])
    for i in range(len(ref)):
        a = (int(mine["jBest"][i]), int(mine["kBest"][i]), int(mine["alphaBest"][i]))
        b = (int(theirs["jBest"][i]), int(theirs["kBest"][i]), int(theirs["alphaBest"][i]))
        assert _same_pair(a, b, alphas), (i, a, b)
    np.testing.assert_allclose(mine["bestLLK"], theirs["bestLLK"], rtol=RTOL)
    np.testing.assert_allclose(mine["sngPP"], theirs["sngPP"], rtol=1e-6, atol=1e-12)
    np.testing.assert_allclose(mine["sngOnlyPP"], theirs["sngOnlyPP"], rtol=1e-6, atol=1e-12)
    return mine, theirs


def assert_llk_match(llk, ref_llk, rtol=RTOL):
    """Engine llksAB dump vs oracle: compared where the engine defines it ([j][0][0], [j][k][n>=1])."""
    defined = ~np.isnan(llk)
    assert defined[:, :, 0, 0].all() and defined[:, :, :, 1:].all()
    np.testing.assert_allclose(llk[defined], ref_llk[defined], rtol=rtol, atol=ATOL)


def permute_cells(d, order):
    """CSR with the droplets re-ordered (order[i] = old id of new droplet i)."""
    cp, rp = np.asarray(d["cell_ptr"], np.int64), np.asarray(d["read_ptr"], np.int64)
    order = np.asarray(order, np.int64)
    ecnt = np.diff(cp)[order]
    new_cp = np.concatenate([[0], np.cumsum(ecnt)])
    ent_idx = np.repeat(cp[order] - new_cp[:-1], ecnt) + np.arange(int(new_cp[-1]))
    rcnt = np.diff(rp)[ent_idx]
    new_rp = np.concatenate([[0], np.cumsum(rcnt)])
    read_idx = np.repeat(rp[ent_idx] - new_rp[:-1], rcnt) + np.arange(int(new_rp[-1]))
    return (new_cp, np.asarray(d["snp_idx"])[ent_idx], new_rp, np.asarray(d["al"])[read_idx],
            np.asarray(d["bq"])[read_idx])
