"""-m gpu parity tests: CUDA engine (through the C ABI) vs the CPU oracle on the same inputs."""
import numpy as np
import pytest

from cramore_b200.synth import make_dataset
from tests.parity import (assert_calls_match, assert_llk_match, assert_results_match, engine_run, oracle_run)

pytestmark = pytest.mark.gpu


def test_genotype_mixing_bit_exact(engine):
    d = make_dataset("C1", nbcs=4, nsnps=777)
    has_gp = (np.arange(777) % 5 != 0).astype(np.uint8)
    d["snp_err"] = np.linspace(-0.1, 1.2, 777)  # exercises both clamps and err == 0
    engine.configure(d["nv"], d["alphas"])
    engine.set_genotypes(d["gp"], d["snp_err"], has_gp)
    gps, _, _, _ = oracle_run(d, has_gp=has_gp, cells=(0, 0))
    assert np.array_equal(engine.genotypes(), gps)


# 2, 3, 5, 9, 11, 21 and 32 alphas: 1, 1, 2, 3, 4, 6 and 9 tile values per lane in the tile kernel
@pytest.mark.parametrize("alphas", [[0, 0.5], [0, 0.25, 0.5], [0, 0.1, 0.2, 0.35, 0.5],
                                    [round(0.0625 * i, 4) for i in range(9)],
                                    [round(0.05 * i, 2) for i in range(11)],
                                    [round(0.025 * i, 3) for i in range(21)],
                                    [round(0.5 * i / 31, 6) for i in range(32)]])
def test_tiles_bit_exact(engine, alphas):
    d = make_dataset("C1", nbcs=50, nsnps=300, rbar=120)  # ~120 reads over 300 SNPs: multi-read entries
    d["alphas"] = np.asarray(alphas, np.float64)
    engine_run(engine, d)
    _, _, _, tiles = oracle_run(d, want_tiles=True)
    assert np.array_equal(engine.tiles(), tiles)


def test_appendix_c_known_answer(engine):
    """SURVEY.md App. C (hand-derived from cmd_cram_demuxlet.cpp:655-821)."""
    import json, os
    g = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "appendix_c.json")))
    # the engine takes float GPs + mixing; feed GPs that are exact in float and err = 0
    # -> use the golden case through doubles is impossible, so check the float-rounded variant vs oracle
    gp = np.asarray(g["gps"], np.float32)
    d = dict(nv=2, alphas=np.asarray(g["alpha"]), gp=gp, snp_err=np.zeros(2), cell_ptr=g["cell_ptr"],
             snp_idx=g["snp_idx"], read_ptr=g["read_ptr"], al=g["al"], bq=g["bq"])
    res = engine_run(engine, d)
    assert np.array_equal(engine.tiles(), np.asarray(g["tiles"]))
    _, ref, ref_llk, _ = oracle_run(d, want_llk=True)
    assert_results_match(res, ref, d["alphas"], ref_llk)
    assert_llk_match(engine.llk(), ref_llk)
    # float rounding of the GPs moves the LLKs by ~1e-8 only: the golden doubles still match loosely
    np.testing.assert_allclose(engine.llk()[0][:, :, 1], np.asarray(g["llk"])[:, :, 1], rtol=1e-6)
    calls, _ = assert_calls_match(res, ref, 2, d["alphas"])
    assert calls["type"][0] == 2  # AMB


@pytest.mark.parametrize("nv,alphas", [
    (2, [0, 0.5]), (3, [0, 0.3]), (8, [0, 0.5]), (33, [0, 0.25, 0.5]), (64, [0, 0.1, 0.2, 0.3, 0.4, 0.5]),
    (65, [round(0.05 * i, 2) for i in range(11)]), (40, [0, 0.1, 0.2, 0.3, 0.35, 0.4, 0.45, 0.5]),
    (5, [0, 0.25, 0.75, 1.0]),                      # grid beyond 0.5: the one-SNP form of the affine path
    (34, [round(0.025 * i, 3) for i in range(21)]),  # two chunks of ten alphas: runtime chunk offset
])
def test_llk_matrix_and_calls(engine, nv, alphas):
    d = make_dataset("C1", nbcs=12, nsnps=500, nv=nv, rbar=60)
    d["alphas"] = np.asarray(alphas, np.float64)
    res = engine_run(engine, d)
    _, ref, ref_llk, _ = oracle_run(d, want_llk=True)
    assert_llk_match(engine.llk(), ref_llk)
    assert_results_match(res, ref, d["alphas"], ref_llk)
    assert_calls_match(res, ref, nv, d["alphas"])


def test_c1_full(engine):
    """BASELINE configs[0]: 8 samples, 20k SNPs, 1,000 barcodes, ~50 reads/cell, alpha {0, 0.5}."""
    d = make_dataset("C1")
    res = engine_run(engine, d)
    _, ref, ref_llk, _ = oracle_run(d, want_llk=True)
    assert_results_match(res, ref, d["alphas"], ref_llk)
    calls, _ = assert_calls_match(res, ref, d["nv"], d["alphas"])
    # sanity of the simulation itself: confident singlets point at the simulated sample
    sng = (calls["type"] == 0) & ~d["is_dbl"]
    assert sng.sum() > 300 and (calls["jBest"][sng] == d["s1"][sng]).mean() > 0.99


def test_edge_cases(engine):
    """empty droplet, droplet whose SNPs all lack genotypes, reads with allele code 2 only,
    a very deep (cell,SNP) entry, ragged droplets."""
    rng = np.random.default_rng(5)
    nv, nsnps = 5, 40
    d = make_dataset("C1", nbcs=1, nsnps=nsnps, nv=nv, rbar=5)
    has_gp = np.ones(nsnps, np.uint8)
    has_gp[[3, 4, 5]] = 0
    cells = [
        [],                                            # no entries at all
        [(3, [(0, 20)]), (4, [(1, 15)])],              # only SNPs without genotypes
        [(7, [(2, 20), (2, 13)])],                     # allele code 2 only
        [(9, [(int(a), int(q)) for a, q in zip(rng.integers(0, 2, 600), rng.integers(13, 21, 600))])],  # deep
        [(s, [(int(rng.integers(0, 2)), 20)]) for s in range(0, 40, 3)],
        [(1, [(0, 20)] * 3), (3, [(1, 20)]), (20, [(1, 13), (0, 14), (2, 20)])],
    ]
    cp, si, rp, al, bq = [0], [], [0], [], []
    for c in cells:
        for s, reads in c:
            si.append(s)
            for a, q in reads:
                al.append(a)
                bq.append(q)
            rp.append(len(al))
        cp.append(len(si))
    d.update(cell_ptr=np.asarray(cp), snp_idx=np.asarray(si, np.int32), read_ptr=np.asarray(rp),
             al=np.asarray(al, np.uint8), bq=np.asarray(bq, np.uint8), alphas=np.asarray([0, 0.25, 0.5]))
    res = engine_run(engine, d, has_gp=has_gp)
    _, ref, ref_llk, tiles = oracle_run(d, has_gp=has_gp, want_llk=True, want_tiles=True)
    assert np.array_equal(engine.tiles(), tiles)
    assert_llk_match(engine.llk(), ref_llk)
    assert_results_match(res, ref, d["alphas"], ref_llk)
    assert res["sngBestLLK"][0] == 0.0 and res["sngBestLLK"][1] == 0.0  # no informative SNP: all LLKs are 0


def test_c2_shape_sample(engine):
    """BASELINE configs[1] shape (64 samples, 11 alphas, ~200 reads/cell) on a barcode sample the
    oracle finishes in seconds."""
    d = make_dataset("C2", nbcs=6, nsnps=20000)
    res = engine_run(engine, d)
    _, ref, ref_llk, _ = oracle_run(d, want_llk=True)
    assert_llk_match(engine.llk(), ref_llk)
    assert_results_match(res, ref, d["alphas"], ref_llk)
    assert_calls_match(res, ref, d["nv"], d["alphas"])


def test_high_depth_shape_sample(engine):
    """BASELINE configs[4] shape: ~2,000 reads per cell, 128 samples, 20% doublets."""
    d = make_dataset("C5", nbcs=2, nsnps=20000)
    res = engine_run(engine, d)
    _, ref, ref_llk, _ = oracle_run(d, want_llk=True)
    assert_llk_match(engine.llk(), ref_llk)
    assert_results_match(res, ref, d["alphas"], ref_llk)


def test_full_size_properties(engine):
    """Size-independent properties at a size the oracle cannot reach: barcode-shard invariance
    (bitwise) and invariance to permuting the droplet order."""
    from cramore_b200.shard import balanced_ranges, slice_csr
    d = make_dataset("C2", nbcs=600, nsnps=200_000)
    res = engine_run(engine, d)
    parts = []
    for b, e in balanced_ranges(d["cell_ptr"], 3):
        parts.append(engine.score(*slice_csr(d["cell_ptr"], d["snp_idx"], d["read_ptr"], d["al"], d["bq"], b, e)))
    assert np.array_equal(np.concatenate(parts), res)
    # reversed droplet order
    from tests.parity import permute_cells
    order = np.arange(d["nbcs"])[::-1]
    res2 = engine.score(*permute_cells(d, order))
    assert np.array_equal(res2[::-1], res)
    # simulated truth is recovered
    from cramore_b200 import classify
    calls = classify(res, d["nv"], d["alphas"])
    sng = (calls["type"] == 0)
    assert (calls["jBest"][sng] == d["s1"][sng]).mean() > 0.99
    dbl = d["is_dbl"] & (calls["type"] == 1)
    assert dbl.sum() > 20
    truth = np.stack([np.minimum(d["s1"], d["s2"]), np.maximum(d["s1"], d["s2"])], 1)[dbl]
    got = np.stack([np.minimum(calls["jBest"], calls["kBest"]), np.maximum(calls["jBest"], calls["kBest"])], 1)[dbl]
    assert (truth == got).all(1).mean() > 0.95


def test_scoring_in_memory_bounded_ranges(engine, monkeypatch):
    """ccu_demux_score splits the droplet store into ranges whose scratch fits the device memory budget
    (CCU_DEMUX_MAX_SCRATCH_MB forces it here): results are bitwise those of the single pass."""
    d = make_dataset("C2", nbcs=300, nsnps=20000)
    whole = engine_run(engine, d)
    launches_whole = engine.timings()["launches"]
    monkeypatch.setenv("CCU_DEMUX_MAX_SCRATCH_MB", "8")  # ~1.4 KB per entry at 64 samples x 11 alphas
    parts = engine_run(engine, d)
    assert engine.timings()["launches"] > 3 * launches_whole  # several ranges were run
    monkeypatch.delenv("CCU_DEMUX_MAX_SCRATCH_MB")
    assert np.array_equal(whole, parts)


def test_scoring_in_two_overlapped_parts(engine, monkeypatch):
    """A large store is scored in two parts, the second part's upload (on a second stream, into a second set of store
    arrays, pointer arrays rebased on the device) under the first part's kernels: results are bitwise those of the single
    pass (CCU_DEMUX_NO_PIPELINE=1), and the inspection calls refuse a handle that holds only the last part."""
    from cramore_b200.engine import EngineError
    d = make_dataset("C2", nbcs=6000, nsnps=200_000)
    assert d["cell_ptr"][-1] >= 1 << 20
    parts = engine_run(engine, d)
    launches_parts = engine.timings()["launches"]
    with pytest.raises(EngineError):
        engine.llk(0, 2)
    monkeypatch.setenv("CCU_DEMUX_NO_PIPELINE", "1")
    whole = engine_run(engine, d)
    assert launches_parts > engine.timings()["launches"]  # two passes against one
    monkeypatch.delenv("CCU_DEMUX_NO_PIPELINE")
    assert np.array_equal(whole, parts)
    assert engine.llk(0, 2).shape[0] == 2


def test_two_samples_and_widest_pool(engine):
    """nv = 2 (one doublet pair; a single sample is refused like one alpha is: the reference's doublet scan would index
    sample -1) and the C3 shape (256 samples, 11 alphas: 65,280 ordered pairs per alpha) on two droplets."""
    from cramore_b200.engine import EngineError
    with pytest.raises(EngineError):
        engine.configure(1, [0, 0.5], 0.5)
    d = make_dataset("C1", nbcs=6, nsnps=300, nv=2, rbar=40)
    res = engine_run(engine, d)
    _, ref, ref_llk, _ = oracle_run(d, want_llk=True)
    assert_llk_match(engine.llk(), ref_llk)
    assert_results_match(res, ref, d["alphas"], ref_llk)
    assert_calls_match(res, ref, 2, d["alphas"])
    d = make_dataset("C3", nbcs=2, nsnps=5000)
    assert d["nv"] == 256 and len(d["alphas"]) == 11
    res = engine_run(engine, d)
    _, ref, ref_llk, _ = oracle_run(d, want_llk=True)
    assert_llk_match(engine.llk(), ref_llk)
    assert_results_match(res, ref, d["alphas"], ref_llk)
    assert_calls_match(res, ref, d["nv"], d["alphas"])


def test_exact_ties_follow_the_scan_order(engine):
    """Samples with identical genotypes tie EXACTLY in every likelihood: best / next must be the ones the reference's
    strict '<' scans keep (cmd_cram_demuxlet.cpp:827-837, :883-906) -- the first in scan order wins, the runner-up
    is the next one met -- for singlets, and for doublets on a grid without alpha = 0.5."""
    for alphas in ([0, 0.5], [0, 0.2, 0.4]):
        d = make_dataset("C1", nbcs=10, nsnps=400, nv=6, rbar=80, seed=77)
        gp = np.array(d["gp"], np.float32).reshape(-1, 6, 3)
        gp[:, 3] = gp[:, 1]   # sample 3 == sample 1
        gp[:, 5] = gp[:, 1]   # sample 5 == sample 1
        gp[:, 4] = gp[:, 0]   # sample 4 == sample 0
        d["gp"] = gp.reshape(np.asarray(d["gp"]).shape)
        d["alphas"] = np.asarray(alphas, np.float64)
        res = engine_run(engine, d)
        _, ref, ref_llk, _ = oracle_run(d, want_llk=True)
        assert_llk_match(engine.llk(), ref_llk)
        for f in ("sBest", "sNext"):
            assert np.array_equal(res[f], ref[f]), (alphas, f)
        if 0.5 in alphas:
            # with alpha = 0.5 the two orientations of a pair differ by rounding in the reference, and which of several
            # exactly tied pairs ends up as runner-up depends on that last bit: equal likelihoods are what can be asked
            assert_results_match(res, ref, d["alphas"], ref_llk)
        else:
            for tag in ("Best", "Next"):
                for x in "12A":
                    assert np.array_equal(res["d%s%s" % (tag, x)], ref["d%s%s" % (tag, x)]), (alphas, tag, x)
        np.testing.assert_allclose(res["dblBestLLK"], ref["dblBestLLK"], rtol=1e-9)
        np.testing.assert_allclose(res["dblNextLLK"], ref["dblNextLLK"], rtol=1e-9)


def test_tiled_epilogue_forms_agree(engine, monkeypatch):
    """The tiled K2 reduces an item's products in the product domain (likelihoods rescaled to the warp's largest
    exponent, REDUX arg-max) and keeps the (mantissa, exponent) form for warps that would lose a candidate below
    2^-1022; CCU_DEMUX_EXACT_EPILOGUE=1 sends every warp through the latter.  Candidates are exact in both (same indices,
    same bits), the sums agree to rounding -- on a pool with exactly tied samples, at 40 samples x 8 alphas (two alpha
    chunks) and at the C2 shape (one chunk of ten)."""
    for nv, alphas in ((40, [0, 0.1, 0.2, 0.3, 0.35, 0.4, 0.45, 0.5]), (64, [round(0.05 * i, 2) for i in range(11)])):
        d = make_dataset("C1", nbcs=24, nsnps=1500, nv=nv, rbar=120, seed=5)
        gp = np.array(d["gp"], np.float32).reshape(-1, nv, 3)
        gp[:, 7] = gp[:, 2]    # exact ties between samples, inside a thread's pair and across lanes / warps
        gp[:, 33] = gp[:, 2]
        gp[:, 20] = gp[:, 19]
        d["gp"] = gp.reshape(np.asarray(d["gp"]).shape)
        d["alphas"] = np.asarray(alphas, np.float64)
        fast = engine_run(engine, d)
        _, ref, ref_llk, _ = oracle_run(d, want_llk=True)
        assert_results_match(fast, ref, d["alphas"], ref_llk)
        monkeypatch.setenv("CCU_DEMUX_EXACT_EPILOGUE", "1")
        exact = engine_run(engine, d)
        monkeypatch.delenv("CCU_DEMUX_EXACT_EPILOGUE")
        for f in ("sBest", "sNext", "dBest1", "dBest2", "dBestA", "dNext1", "dNext2", "dNextA",
                  "sngBestLLK", "sngNextLLK", "dblBestLLK", "dblNextLLK"):
            assert np.array_equal(fast[f], exact[f]), (nv, f)
        for f in ("sumLLK", "sngLLK"):
            np.testing.assert_allclose(fast[f], exact[f], rtol=1e-13, atol=0, err_msg=f)


def test_twenty_thousand_reads_on_one_snp(engine):
    """One (droplet, SNP) entry with 20,000 reads: the per-read product of cmd_cram_demuxlet.cpp:655-700 is renormalised
    by its running maximum in the reference; the tile kernel must survive the same depth without underflow."""
    rng = np.random.default_rng(3)
    d = make_dataset("C1", nbcs=1, nsnps=50, nv=4, rbar=5)
    n = 20000
    al = (rng.random(n) < 0.5).astype(np.uint8)
    d.update(cell_ptr=np.asarray([0, 2]), snp_idx=np.asarray([7, 9], np.int32), read_ptr=np.asarray([0, n, n + 1]),
             al=np.concatenate([al, [1]]).astype(np.uint8), bq=np.concatenate([rng.integers(13, 21, n), [20]]).astype(np.uint8))
    res = engine_run(engine, d)
    _, ref, ref_llk, tiles = oracle_run(d, want_llk=True, want_tiles=True)
    assert np.array_equal(engine.tiles(), tiles)
    assert_llk_match(engine.llk(), ref_llk)
    assert_results_match(res, ref, d["alphas"], ref_llk)


def test_argument_errors_are_reported_not_computed(engine):
    from cramore_b200.engine import EngineError
    d = make_dataset("C1", nbcs=2, nsnps=50, nv=3, rbar=10)
    with pytest.raises(EngineError):
        engine.configure(3, [0.5], 0.5)          # one alpha: the prior divides by nA - 1 (:794)
    with pytest.raises(EngineError):
        engine.configure(3, [0.1, 0.5], 0.5)     # element 0 must be 0
    engine.configure(3, [0, 0.5], 0.5)
    engine.set_genotypes(d["gp"], d["snp_err"], None)
    bad = np.asarray(d["snp_idx"]).copy()
    bad[0] = 50                                   # SNP id out of range
    with pytest.raises(EngineError):
        engine.score(d["cell_ptr"], bad, d["read_ptr"], d["al"], d["bq"])
    res = engine.score(d["cell_ptr"], d["snp_idx"], d["read_ptr"], d["al"], d["bq"])   # the handle is still usable
    assert len(res) == 2
