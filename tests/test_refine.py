"""CPU tests of ccu_demux_refine (cramore_b200/csrc/demux_refine.cpp): the host stage that re-scores the candidates a
.best row names in the reference's own operation order.  The engine's output is emulated from the oracle's: mirror
pairs put into the canonical j < k orientation and every likelihood disturbed at the 1e-13 level (what the GPU's
product-domain accumulation leaves); after the refinement every index and every likelihood must equal the reference's
own run bit for bit, and the text rows must be the reference's bytes."""
import numpy as np
import pytest

from cramore_b200 import best_row
from cramore_b200.engine import refine
from oracle import pyoracle
from tests import refcase


def engine_like(res, alphas, rng):
    """What the GPU hands back for the oracle's exact result: same hypotheses, engine orientation, ~1e-13 noise."""
    out = res.copy()
    for c in range(len(out)):
        r = out[c]
        n1, n2 = r["dBestA"], r["dNextA"]
        mirrors = (n1 >= 1 and n2 >= 1 and alphas[n2] == 1.0 - alphas[n1] and r["dBest1"] == r["dNext2"]
                   and r["dBest2"] == r["dNext1"])
        if mirrors and r["dBest1"] > r["dBest2"]:  # demux_finalize_kernel reports such a pair as j < k first
            for a, b in (("dBest1", "dNext1"), ("dBest2", "dNext2"), ("dBestA", "dNextA"), ("dblBestLLK", "dblNextLLK")):
                out[a][c], out[b][c] = r[b], r[a]
        if rng.random() < 0.3:  # and it may rank two close singlets the other way round
            for a, b in (("sBest", "sNext"), ("sngBestLLK", "sngNextLLK")):
                out[a][c], out[b][c] = out[b][c], out[a][c]
    for f in ("sngBestLLK", "sngNextLLK", "dblBestLLK", "dblNextLLK"):
        out[f] *= 1.0 + rng.uniform(-2e-13, 2e-13, len(out))
    return out


@pytest.mark.parametrize("name", refcase.DEMUX_CASES)
def test_refined_rows_are_the_references_bytes(name, built):
    fx = refcase.load(name)
    d, alphas, prior, snp_err, has_gp = refcase.demux_inputs(fx)
    gps = pyoracle.mix_genotypes(d["gp"], snp_err, has_gp)
    exact, _, _ = pyoracle.demux_score(d["cell_ptr"], d["snp_idx"], d["read_ptr"], d["al"], d["bq"], gps, alphas, prior,
                                       has_gp=has_gp)
    rng = np.random.default_rng(5)
    noisy = engine_like(exact, alphas, rng)
    assert not np.array_equal(noisy["dblBestLLK"], exact["dblBestLLK"])
    for nthreads in (1, 4):
        fixed = refine(noisy, d["cell_ptr"], d["snp_idx"], d["read_ptr"], d["al"], d["bq"], d["nv"], alphas, d["gp"],
                       snp_err, has_gp, nthreads=nthreads)
        for f in ("sBest", "sNext", "dBest1", "dBest2", "dBestA", "dNext1", "dNext2", "dNextA", "sngBestLLK", "sngNextLLK",
                  "dblBestLLK", "dblNextLLK"):
            assert np.array_equal(fixed[f], exact[f]), (name, f, np.nonzero(fixed[f] != exact[f])[0][:5])
    # the text rows written from the refined records are the reference's .best
    lines = fx["best_text"].splitlines(keepends=True)
    order = sorted(range(d["nbcs"]), key=lambda i: d["barcodes"][i])
    cp, rp = d["cell_ptr"], d["read_ptr"]
    fixed["sumLLK"], fixed["sngLLK"] = exact["sumLLK"], exact["sngLLK"]  # untouched by the refinement
    for rank, c in enumerate(order):
        mine = best_row(rank, d["barcodes"][c], int(cp[c + 1] - cp[c]), int(rp[cp[c + 1]] - rp[cp[c]]), fixed[c], d["nv"],
                        alphas, prior, d["sample_ids"])
        assert mine == lines[1 + rank]


def test_refine_mixed_form_and_missing_genotypes(built):
    """The double-pointer form (scl.snps[s].gps, NULL where the VCF had no record) gives the same answer."""
    import ctypes as C
    from cramore_b200 import load_library
    fx = refcase.load("demux_missing_gp_r2")
    d, alphas, prior, snp_err, has_gp = refcase.demux_inputs(fx)
    gps = pyoracle.mix_genotypes(d["gp"], snp_err, has_gp)
    exact, _, _ = pyoracle.demux_score(d["cell_ptr"], d["snp_idx"], d["read_ptr"], d["al"], d["bq"], gps, alphas, prior,
                                       has_gp=has_gp)
    noisy = engine_like(exact, alphas, np.random.default_rng(9))
    ptrs = (C.c_void_p * d["nsnps"])(*[gps[s].ctypes.data if has_gp[s] else None for s in range(d["nsnps"])])
    out = noisy.copy()
    cp = np.ascontiguousarray(d["cell_ptr"], np.int64)
    si = np.ascontiguousarray(d["snp_idx"], np.int32)
    rp = np.ascontiguousarray(d["read_ptr"], np.int64)
    al = np.ascontiguousarray(d["al"], np.uint8)
    bq = np.ascontiguousarray(d["bq"], np.uint8)
    a = np.ascontiguousarray(alphas, np.float64)
    rc = load_library().ccu_demux_refine_mixed(C.c_int64(d["nbcs"]), C.c_void_p(cp.ctypes.data), C.c_void_p(si.ctypes.data),
                                               C.c_void_p(rp.ctypes.data), C.c_void_p(al.ctypes.data),
                                               C.c_void_p(bq.ctypes.data), C.c_int32(d["nv"]), C.c_int32(len(a)),
                                               C.c_void_p(a.ctypes.data), C.c_int64(d["nsnps"]), ptrs, C.c_int32(2),
                                               C.c_void_p(out.ctypes.data))
    assert rc == 0
    for f in ("sBest", "sNext", "dBest1", "dBest2", "dBestA", "dNext1", "dNext2", "dNextA", "sngBestLLK", "sngNextLLK",
              "dblBestLLK", "dblNextLLK"):
        assert np.array_equal(out[f], exact[f]), f


def test_refine_rejects_bad_arguments(built):
    d, alphas = {"nv": 4}, np.asarray([0.0, 0.5])
    res = np.zeros(1, pyoracle.DEMUX_RESULT_DTYPE)
    gp = np.zeros((3, 4, 3), np.float32)
    from cramore_b200.engine import EngineError
    with pytest.raises(EngineError):  # SNP id outside the genotype matrix
        refine(res, [0, 1], [7], [0, 1], [0], [20], 4, alphas, gp, 0.1)


@pytest.mark.parametrize("seed", range(6))
def test_refine_on_grids_with_complements_against_the_live_reference(built, tmp_path, seed):
    """Grids holding alpha together with 1 - alpha: (j,k,alpha) and (k,j,1-alpha) are one model, and which of the two
    the reference prints is rounding in its own summation.  The refinement re-scores both and reproduces the
    reference's file (oracle/_ref/ref_cramore run on the same inputs)."""
    import os
    from cramore_b200.synth import make_dataset, write_plp, write_vcf
    from tests.test_reference_live import REF as REF_BIN, _run_ref
    if not os.path.exists(REF_BIN):
        pytest.skip("oracle/_ref/ref_cramore not built")
    rng = np.random.default_rng(700 + seed)
    nv = int(rng.integers(2, 9))
    d = make_dataset("C1", nbcs=int(rng.integers(10, 40)), nsnps=int(rng.integers(60, 300)), rbar=int(rng.integers(20, 120)),
                     nv=nv, seed=9100 + seed, with_barcodes=True)
    alphas = [[0.0, 0.3, 0.7], [0.0, 0.25, 0.5, 0.75], [0.0, 0.1, 0.4, 0.6, 0.9], [0.0, 0.2, 0.5, 0.8],
              [0.0, 0.45, 0.55], [0.0, 0.05, 0.35, 0.5, 0.65]][seed]
    pre = str(tmp_path / "in")
    write_plp(d, pre)
    write_vcf(d, pre + ".vcf")
    args = ["demuxlet", "--plp", "in", "--vcf", "in.vcf", "--field", "GP", "--out", "out"]
    for a in alphas:
        args += ["--alpha", repr(a)]
    _run_ref(args, str(tmp_path))
    ref_lines = open(str(tmp_path / "out.best")).read().splitlines(keepends=True)
    ac = np.asarray(d["geno"], np.int64).sum(axis=1)
    has_gp = ((ac >= 1) & (2 * nv - ac >= 1)).astype(np.uint8)  # --min-mac 1 (cmd_cram_demuxlet.cpp:27)
    al = np.asarray(alphas, np.float64)
    snp_err = np.full(d["nsnps"], 0.1)
    gps = pyoracle.mix_genotypes(d["gp"], snp_err, has_gp)
    exact, _, _ = pyoracle.demux_score(d["cell_ptr"], d["snp_idx"], d["read_ptr"], d["al"], d["bq"], gps, al, 0.5,
                                       has_gp=has_gp)
    noisy = engine_like(exact, al, rng)
    # the engine may also report the two orientations the other way round
    for c in range(len(noisy)):
        if rng.random() < 0.5:
            r = noisy[c].copy()
            n1, n2 = r["dBestA"], r["dNextA"]
            if al[n2] == 1.0 - al[n1] and r["dBest1"] == r["dNext2"] and r["dBest2"] == r["dNext1"]:
                for a, b in (("dBest1", "dNext1"), ("dBest2", "dNext2"), ("dBestA", "dNextA"), ("dblBestLLK", "dblNextLLK")):
                    noisy[a][c], noisy[b][c] = r[b], r[a]
    fixed = refine(noisy, d["cell_ptr"], d["snp_idx"], d["read_ptr"], d["al"], d["bq"], nv, al, d["gp"], snp_err, has_gp)
    fixed["sumLLK"], fixed["sngLLK"] = exact["sumLLK"], exact["sngLLK"]
    order = sorted(range(d["nbcs"]), key=lambda i: d["barcodes"][i])
    cp, rp = d["cell_ptr"], d["read_ptr"]
    for rank, c in enumerate(order):
        mine = best_row(rank, d["barcodes"][c], int(cp[c + 1] - cp[c]), int(rp[cp[c + 1]] - rp[cp[c]]), fixed[c], nv, al,
                        0.5, d["sample_ids"])
        assert mine == ref_lines[1 + rank], (seed, c)
