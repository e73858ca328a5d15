"""CPU tests (no GPU): the oracle against the hand-derived known answers of SURVEY.md App. C, the
reference's own PhredHelper.cpp (oracle/_ref) and a second, independent pure-Python restatement of
cmd_cram_demuxlet.cpp:655-906 on small random cases."""
import json
import math
import os

import numpy as np
import pytest

from oracle import pyoracle

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def golden():
    return json.load(open(os.path.join(HERE, "golden", "appendix_c.json")))


# ---------------------------------------------------------------- independent python restatement
def py_phred():
    err = [0.75 if q <= 1 else math.pow(0.1, q * 0.1) for q in range(256)]  # PhredHelper.cpp:29-33
    return err, [1.0 - e for e in err]


def py_tile(al, bq, alpha):
    """cmd_cram_demuxlet.cpp:655-725, literal loops."""
    err, mat = py_phred()
    nA = len(alpha)
    pG = [1.0] * (nA * 9)
    for a, q in zip(al, bq):
        if a == 2:
            continue
        pR = mat[q] if a == 0 else err[q] / 3.0
        pA = mat[q] if a == 1 else err[q] / 3.0
        mx = 0.0
        for k in range(nA):
            for l in range(3):
                for m in range(3):
                    p = 0.5 * l + (m - l) * 0.5 * alpha[k]
                    pG[k * 9 + l * 3 + m] *= (pR * (1.0 - p) + pA * p)
                    mx = max(mx, pG[k * 9 + l * 3 + m])
        pG = [v / mx for v in pG]
    pG = [v + 1e-10 for v in pG]
    mx = max(pG)
    return [v / mx for v in pG]


def py_log_add(a, b):
    return a + math.log(1.0 + math.exp(b - a)) if a > b else b + math.log(1.0 + math.exp(a - b))


def py_demux(cell_ptr, snp_idx, read_ptr, al, bq, gps, alpha, prior):
    """cmd_cram_demuxlet.cpp:636-906 for every cell; returns (llk[cells][nv][nv][nA], results)."""
    nv, nA = len(gps[0]), len(alpha)
    out_llk, out = [], []
    for c in range(len(cell_ptr) - 1):
        llk = np.zeros((nv, nv, nA))
        for e in range(cell_ptr[c], cell_ptr[c + 1]):
            r0, r1 = read_ptr[e], read_ptr[e + 1]
            pG = py_tile(al[r0:r1], bq[r0:r1], alpha)
            g = gps[snp_idx[e]]
            for j in range(nv):
                for k in range(nv):
                    s = [0.0] * nA
                    for l in range(3):
                        for m in range(3):
                            p = g[j][l] * g[k][m]
                            for n in range(nA):
                                s[n] += p * pG[n * 9 + l * 3 + m]
                    for n in range(nA):
                        llk[j, k, n] += math.log(s[n])
        lsp = math.log((1.0 - prior) / nv)
        ld1 = math.log(prior / nv / (nv - 1.0) / (nA - 1.0))
        ld2 = math.log(prior / nv / (nv - 1.0) / (nA - 1.0) * 2)
        sumLLK = sngLLK = -1e-300
        for j in range(nv):
            sumLLK = py_log_add(sumLLK, llk[j, 0, 0] + lsp)
            sngLLK = py_log_add(sngLLK, llk[j, 0, 0] + lsp)
            for k in range(nv):
                if j == k:
                    continue
                for n in range(1, nA):
                    if alpha[n] == 0.5:
                        if k > j:
                            continue
                        sumLLK = py_log_add(sumLLK, llk[j, k, n] + ld2)
                    else:
                        sumLLK = py_log_add(sumLLK, llk[j, k, n] + ld1)
        sb = sn = -1
        sbl = snl = -1e300
        for j in range(nv):
            v = llk[j, 0, 0]
            if sbl < v:
                sn, snl, sb, sbl = sb, sbl, j, v
            elif snl < v:
                sn, snl = j, v
        db = dn = (-1, -1, -1)
        dbl = dnl = -1e300
        for j in range(nv):
            for k in range(nv):
                if j == k:
                    continue
                for n in range(1, nA):
                    v = llk[j, k, n]
                    if dbl < v:
                        dn, dnl, db, dbl = db, dbl, (j, k, n), v
                    elif dnl < v:
                        dn, dnl = (j, k, n), v
        out_llk.append(llk)
        out.append(dict(sBest=sb, sNext=sn, sngBestLLK=sbl, sngNextLLK=snl, dBest=db, dNext=dn, dblBestLLK=dbl,
                        dblNextLLK=dnl, sumLLK=sumLLK, sngLLK=sngLLK))
    return np.asarray(out_llk), out


# ---------------------------------------------------------------- tests
def test_phred_tables_match_reference_build():
    err, mat = pyoracle.phred_tables()
    perr, pmat = py_phred()
    assert np.array_equal(err, np.asarray(perr)) and np.array_equal(mat, np.asarray(pmat))
    ref = pyoracle.ref_phred_tables()  # the reference's own PhredHelper.cpp compiled into oracle/_ref
    if ref is None:
        pytest.skip("oracle/_ref not built (no /root/reference on this machine)")
    assert np.array_equal(err, ref[0]) and np.array_equal(mat, ref[1])


def test_log_add():
    for a, b in [(-1.0, -2.0), (-2.0, -1.0), (-1e-300, -700.0), (-50.0, -50.0), (0.0, -1e300)]:
        assert pyoracle.log_add(a, b) == py_log_add(a, b)


def test_appendix_c_tiles(golden):
    """Tile known answers computed by hand from cmd_cram_demuxlet.cpp:666-725."""
    rp = golden["read_ptr"]
    for e in range(2):
        t = pyoracle.demux_tile(golden["al"][rp[e]:rp[e + 1]], golden["bq"][rp[e]:rp[e + 1]], golden["alpha"])
        assert np.array_equal(t, np.asarray(golden["tiles"][e], np.float64))


def test_appendix_c_llk_and_decision(golden):
    gps = np.asarray(golden["gps"], np.float64)
    res, llk, tiles = pyoracle.demux_score(golden["cell_ptr"], golden["snp_idx"], golden["read_ptr"], golden["al"],
                                           golden["bq"], gps, golden["alpha"], golden["doublet_prior"],
                                           want_llk=True, want_tiles=True)
    assert np.array_equal(tiles, np.asarray(golden["tiles"]))
    assert np.array_equal(llk[0], np.asarray(golden["llk"]))  # bit-exact: same operation order
    r = res[0]
    assert r["sumLLK"] == golden["sumLLK"] and r["sngLLK"] == golden["sngLLK"]
    assert (r["sBest"], r["sNext"]) == (golden["sBest"], golden["sNext"])
    assert [r["dBest1"], r["dBest2"], r["dBestA"]] == golden["dBest"]
    assert [r["dNext1"], r["dNext2"], r["dNextA"]] == golden["dNext"]
    call = pyoracle.demux_classify(res, 2, golden["alpha"], golden["doublet_prior"])[0]
    assert call["type"] == 2  # AMB
    assert call["jBest"] == 0 and call["kBest"] == 0
    assert abs(call["sngPP"] - golden["sngPP"]) < 1e-15


@pytest.mark.parametrize("seed,nv,alpha", [(1, 2, [0, 0.5]), (2, 3, [0, 0.25, 0.5]), (3, 4, [0, 0.1, 0.3])])
def test_oracle_vs_python_restatement(seed, nv, alpha):
    rng = np.random.default_rng(seed)
    nsnps, ncell = 12, 4
    gp = rng.dirichlet([1, 1, 1], size=(nsnps, nv)).astype(np.float32)
    gps = pyoracle.mix_genotypes(gp, np.full(nsnps, 0.1))
    cp, si, rp, al, bq = [0], [], [0], [], []
    for _ in range(ncell):
        for s in sorted(rng.choice(nsnps, size=rng.integers(1, 7), replace=False)):
            si.append(int(s))
            for _ in range(int(rng.integers(1, 4))):
                al.append(int(rng.choice([0, 1, 2], p=[0.5, 0.45, 0.05])))
                bq.append(int(rng.integers(13, 41)))
            rp.append(len(al))
        cp.append(len(si))
    res, llk, _ = pyoracle.demux_score(cp, si, rp, al, bq, gps, alpha, 0.5, want_llk=True)
    pllk, pres = py_demux(cp, si, rp, al, bq, gps.tolist(), alpha, 0.5)
    np.testing.assert_allclose(llk, pllk, rtol=1e-13, atol=1e-13)
    for c in range(ncell):
        assert res["sBest"][c] == pres[c]["sBest"] and res["sNext"][c] == pres[c]["sNext"]
        assert (res["dBest1"][c], res["dBest2"][c], res["dBestA"][c]) == pres[c]["dBest"]
        assert (res["dNext1"][c], res["dNext2"][c], res["dNextA"][c]) == pres[c]["dNext"]
        for f in ("sngBestLLK", "dblBestLLK", "sumLLK", "sngLLK"):
            assert abs(res[f][c] - pres[c][f]) <= 1e-12 * max(1.0, abs(pres[c][f]))


def test_mix_genotypes_formula():
    """sc_drop_seq.cpp:287-315: avg seeded with 1e-10, error clamped to [0, 0.999]."""
    rng = np.random.default_rng(7)
    gp = rng.dirichlet([1, 1, 1], size=(5, 3)).astype(np.float32)
    err = np.asarray([0.1, 0.0, 1.5, -0.2, 0.5])
    has = np.asarray([1, 1, 1, 1, 0], np.uint8)
    out = pyoracle.mix_genotypes(gp, err, has)
    for s in range(5):
        if not has[s]:
            assert np.all(out[s] == 0)
            continue
        g = gp[s].astype(np.float64)
        avg = np.full(3, 1e-10)
        for i in range(3):
            avg = avg + g[i]
        avg = avg / ((avg[0] + avg[1]) + avg[2])
        e = min(max(err[s], 0.0), 0.999)
        want = (1 - e) * g + e * avg if e > 0 else g
        np.testing.assert_allclose(out[s], want, rtol=1e-15)


def test_single_read_tile_is_affine_mixture():
    """The identity the CUDA fast path rests on: with at most one informative read the tile
    satisfies pG[n][l][m] = (1-alpha_n) pG[0][l][.] + alpha_n pG[0][m][.] (to rounding)."""
    alpha = [round(0.05 * i, 2) for i in range(11)]
    for al, bq in [([0], [20]), ([1], [13]), ([1, 2], [40, 20]), ([2, 2], [20, 20]), ([], [])]:
        t = pyoracle.demux_tile(al, bq, alpha).reshape(11, 3, 3)
        v = t[0, :, 0]
        for n, a in enumerate(alpha):
            want = (1 - a) * v[:, None] + a * v[None, :]
            np.testing.assert_allclose(t[n], want, rtol=2e-12)


def test_best_row_format():
    """Header and row layout of cmd_cram_demuxlet.cpp:629 / :993-1013."""
    d = dict(cell_ptr=[0, 1], snp_idx=[0], read_ptr=[0, 1], al=[0], bq=[20])
    gps = np.asarray([[[0.9, 0.08, 0.02], [0.05, 0.15, 0.80]]])
    res, _, _ = pyoracle.demux_score(d["cell_ptr"], d["snp_idx"], d["read_ptr"], d["al"], d["bq"], gps, [0, 0.5], 0.5)
    row = pyoracle.demux_best_row(3, "ACGT-1", 1, 1, res[0], 2, [0, 0.5], 0.5, ["A", "B"])
    f = row.rstrip("\n").split("\t")
    assert len(f) == 20 and f[0] == "3" and f[1] == "ACGT-1" and f[4] in ("SNG", "DBL", "AMB")
    assert f[5].count(",") == 2 and row.endswith("\n")
