"""-m gpu parity tests of the freemuxlet kernels (through the C ABI) vs the CPU oracle."""
import os

import numpy as np
import pytest

from cramore_b200.synth import make_dataset
from oracle import pyoracle

pytestmark = pytest.mark.gpu
RTOL = 1e-9


@pytest.fixture(scope="module")
def feng(built):
    from cramore_b200 import FreemuxEngine
    eng = FreemuxEngine(0)
    yield eng
    eng.close()


def small_c4(nbcs=300, nsnps=3000, rbar=150, seed=None):
    d = make_dataset("C4", nbcs=nbcs, nsnps=nsnps, rbar=rbar, seed=seed)
    d["plp"] = pyoracle.fmx_pileup(d["read_ptr"], d["al"], d["bq"], 0.5)
    return d


def noisy_init(d, K, frac_unassigned=0.2, frac_wrong=0.05, seed=3):
    rng = np.random.default_rng(seed)
    a = (d["s1"] % K).astype(np.int32)
    wrong = rng.random(len(a)) < frac_wrong
    a[wrong] = rng.integers(0, K, wrong.sum())
    a[d["is_dbl"]] = -1
    a[rng.random(len(a)) < frac_unassigned] = -1
    return a


def upload(feng, d, K, geno_error=0.0, prior=0.5):
    feng.configure_fmx(K, prior, geno_error)
    feng.upload_fmx(d["cell_ptr"], d["snp_idx"], d["read_ptr"], d["al"], d["bq"], d["af"])


def assert_fmx_results(res, ref, ref_llk, K):
    for f in ("sngBestLLK", "sngNextLLK", "dblBestLLK", "dblNextLLK", "sumLLK", "sngLLK", "bestLLK", "nextLLK", "bestPP"):
        np.testing.assert_allclose(res[f], ref[f], rtol=RTOL, atol=1e-9, err_msg=f)
    np.testing.assert_allclose(res["sngPP"], ref["sngPP"], rtol=1e-6, atol=1e-12)
    np.testing.assert_allclose(res["sngOnlyPP"], ref["sngOnlyPP"], rtol=1e-6, atol=1e-12)

    def pair(j, k):
        return j * (j + 1) // 2 + k

    for i in range(len(ref)):
        for a, b in (("sBest", "sBest"), ("sNext", "sNext")):
            x, y = int(res[a][i]), int(ref[b][i])
            if x != y:  # acceptable only for a tie within tolerance
                assert abs(ref_llk[i, pair(x, x)] - ref_llk[i, pair(y, y)]) <= RTOL * abs(ref_llk[i, pair(y, y)])
        for t in ("Best", "Next"):
            x = (int(res["d%s1" % t][i]), int(res["d%s2" % t][i]))
            y = (int(ref["d%s1" % t][i]), int(ref["d%s2" % t][i]))
            if x != y:
                assert abs(ref_llk[i, pair(*x)] - ref_llk[i, pair(*y)]) <= RTOL * abs(ref_llk[i, pair(*y)])
    assert np.array_equal(res["type"], ref["type"])
    same = res["sBest"] == ref["sBest"]
    assert np.array_equal(res["jBest"][same], ref["jBest"][same])


def test_shared_reciprocal_division_is_ieee(feng):
    """F1/F3 divide nine likelihoods by one sum through a shared refined reciprocal: 9 * 2^26 quotients
    must equal __ddiv_rn bit for bit (including the operand ranges that take the guarded slow path)."""
    assert feng.selftest_division(1 << 26, seed=12345) == 0
    assert feng.selftest_division(1 << 20, seed=7) == 0


def test_pileups_bit_exact(feng):
    d = small_c4(nbcs=200, nsnps=500, rbar=300)  # ~300 reads over 500 SNPs: many multi-read entries
    upload(feng, d, 4)
    got = feng.pileups()
    ref = d["plp"]
    for f in ("nreads", "nref", "nalt"):
        assert np.array_equal(got[f], ref[f]), f
    assert np.array_equal(got["gls"], ref["gls"])
    np.testing.assert_allclose(got["logdenom"], ref["logdenom"], rtol=1e-12, atol=1e-13)


def test_lmix_scores(feng):
    d = small_c4()
    upload(feng, d, 4)
    l0, l2 = feng.lmix()
    r0, r2 = pyoracle.fmx_lmix(d["cell_ptr"], d["snp_idx"], d["plp"], d["af"])
    np.testing.assert_allclose(l0, r0, rtol=RTOL)
    np.testing.assert_allclose(l2, r2, rtol=RTOL)


# shapes: ~100 droplets per SNP (some SNP beyond 128: the general kernel), ~15 per SNP (balanced work lists, one
# pass), ~65 per SNP (balanced work lists; with K = 2 the per-thread lists exceed one pass)
@pytest.mark.parametrize("shape", [(400, 800, 200), (300, 3000, 150), (400, 1200, 200)])
@pytest.mark.parametrize("K", [2, 5, 16])
def test_mstep_bit_exact(feng, K, shape):
    d = small_c4(nbcs=shape[0], nsnps=shape[1], rbar=shape[2])
    upload(feng, d, K)
    assign = noisy_init(d, K)
    feng.mstep(assign)
    got = feng.clusters()
    ref = pyoracle.fmx_mstep(d["cell_ptr"], d["snp_idx"], d["plp"], assign, K, d["nsnps"])
    for f in ("nreads", "nref", "nalt"):
        assert np.array_equal(got[f], ref[f]), f
    assert np.array_equal(got["gls"], ref["gls"])  # ordered, clamped fold: bit-exact


@pytest.mark.parametrize("K", [4, 16])
def test_estep_deep_droplets_take_the_log_form(feng, K):
    """~2,900 covered SNPs per droplet: the next singlet lies more than 708 nats (2^-1022) below the best one, so the
    product-domain epilogue cannot show it and must hand the droplet to the log-domain form; same results either way."""
    d = small_c4(nbcs=24, nsnps=4000, rbar=5000, seed=11)
    upload(feng, d, K)
    assign = (d["s1"] % K).astype(np.int32)
    assign[d["is_dbl"]] = -1
    feng.mstep(assign)
    res, _ = feng.estep(apply_geno_error=False)
    clust = pyoracle.fmx_mstep(d["cell_ptr"], d["snp_idx"], d["plp"], assign, K, d["nsnps"])
    ref, ref_llk = pyoracle.fmx_estep(d["cell_ptr"], d["snp_idx"], d["plp"], d["af"], clust, 0.5, 0.0, False, want_llk=True)
    assert (ref["sngBestLLK"] - ref["sngNextLLK"]).max() > 750.0  # the case is what it claims to be
    assert_fmx_results(res, ref, ref_llk, K)
    np.testing.assert_allclose(feng.fmx_llk(), ref_llk, rtol=RTOL, atol=1e-12)


@pytest.mark.parametrize("K,geno_error", [(2, 0.0), (3, 0.1), (16, 0.0), (16, 0.1), (20, 0.05), (33, 0.0)])
def test_estep_llk_and_decisions(feng, K, geno_error):
    d = small_c4(nbcs=120, nsnps=2000, rbar=150)
    upload(feng, d, K, geno_error)
    assign = noisy_init(d, K)
    feng.mstep(assign)
    res, nxt = feng.estep(apply_geno_error=geno_error > 0)
    clust = pyoracle.fmx_mstep(d["cell_ptr"], d["snp_idx"], d["plp"], assign, K, d["nsnps"])
    ref, ref_llk = pyoracle.fmx_estep(d["cell_ptr"], d["snp_idx"], d["plp"], d["af"], clust, 0.5, geno_error,
                                      geno_error > 0, want_llk=True)
    np.testing.assert_allclose(feng.fmx_llk(), ref_llk, rtol=RTOL, atol=1e-12)
    assert_fmx_results(res, ref, ref_llk, K)
    want_next = np.where(ref["type"] == 0, ref["jBest"], -1)
    same = res["sBest"] == ref["sBest"]
    assert np.array_equal(nxt[same], want_next[same])


def test_em_loop_matches_oracle(feng):
    """cmd_cram_freemuxlet.cpp:457-653: initial M-step, then E/decision/M iterations, geno_error on the
    last iteration only; final cluster statistics are bit-identical when every assignment agrees."""
    K, niter, ge = 16, 4, 0.1
    d = small_c4(nbcs=600, nsnps=4000, rbar=200)
    upload(feng, d, K, ge)
    init = noisy_init(d, K, frac_unassigned=0.5)
    res, iters = feng.run_em(init, niter=niter)
    assert iters == niter
    clust = pyoracle.fmx_mstep(d["cell_ptr"], d["snp_idx"], d["plp"], init, K, d["nsnps"])
    for it in range(niter):
        ref, ref_llk = pyoracle.fmx_estep(d["cell_ptr"], d["snp_idx"], d["plp"], d["af"], clust, 0.5, ge,
                                          it + 1 == niter, want_llk=True)
        assign = np.where(ref["type"] == 0, ref["jBest"], -1).astype(np.int32)
        clust = pyoracle.fmx_mstep(d["cell_ptr"], d["snp_idx"], d["plp"], assign, K, d["nsnps"])
    assert_fmx_results(res, ref, ref_llk, K)
    got = feng.clusters()
    assert np.array_equal(got["gls"], clust["gls"]) and np.array_equal(got["nreads"], clust["nreads"])
    # the simulation is recovered: confident singlets sit in the cluster of their sample
    sng = (res["type"] == 0) & ~d["is_dbl"]
    assert sng.sum() > 300 and (res["jBest"][sng] == d["s1"][sng] % K).mean() > 0.98


def test_freemux2_flavour_early_stop(feng):
    """cmd_cram_freemux2.cpp: geno_error on every iteration and stop once no call changes."""
    K = 8
    d = make_dataset("C4", nbcs=300, nsnps=3000, rbar=200, nv=8)
    upload(feng, d, K, 0.1)
    init = noisy_init(d, K, frac_unassigned=0.3)
    res, iters = feng.run_em(init, niter=30, geno_error_every_iter=True, stop_when_stable=True)
    assert 2 <= iters < 30
    res2, iters2 = feng.run_em(init, niter=iters, geno_error_every_iter=True)
    assert np.array_equal(res, res2)


def test_shard_invariance_and_slab_sum(feng):
    """Multi-GPU partition emulated on one GPU: E-step over barcode shards and M-step over SNP slabs
    reproduce the unsharded results bitwise; slab statistics are disjoint, so their sum (what the NCCL
    all-reduce computes) is the full array exactly."""
    import ctypes as C
    K = 6
    d = small_c4(nbcs=250, nsnps=1500, rbar=150)
    upload(feng, d, K)
    assign = noisy_init(d, K)
    feng.mstep(assign)
    full_stats = np.array(feng.clusters())
    res_full, nxt_full = feng.estep()
    parts = []
    for b, e in ((0, 90), (90, 250)):
        feng.set_shard(b, e, 0, d["nsnps"])
        r, _ = feng.estep()
        parts.append(r)
    assert np.array_equal(np.concatenate(parts), res_full)
    total = None
    for b, e in ((0, 400), (400, 1100), (1100, 1500)):
        feng.set_shard(0, d["nbcs"], b, e)
        feng.mstep(assign)
        c = feng.clusters()
        outside = np.ones(d["nsnps"], bool)
        outside[b:e] = False
        assert np.all(c["gls"][:, outside] == 0) and np.all(c["nreads"][:, outside] == 0)
        s = np.concatenate([c["gls"], c["nreads"][..., None].astype(float), c["nref"][..., None].astype(float),
                            c["nalt"][..., None].astype(float)], axis=2)
        total = s if total is None else total + s
    assert np.array_equal(total[..., :9], full_stats["gls"])
    assert np.array_equal(total[..., 9], full_stats["nreads"])


def test_posterior_exchange_path_is_bitwise_equal(feng):
    """ccu_fmx_posteriors_dev + ccu_fmx_estep_post_dev (the multi-GPU per-iteration form) give the results of
    ccu_fmx_estep_dev bitwise; slab posteriors are disjoint, so their sum over slabs is the full array."""
    K = 7
    d = small_c4(nbcs=200, nsnps=1200, rbar=150)
    upload(feng, d, K, 0.1)
    assign = noisy_init(d, K)
    feng.mstep(assign)
    feng.estep_dev(True)
    ref = feng.download_results()
    feng.posteriors_dev(True)
    feng.estep_post_dev()
    assert np.array_equal(feng.download_results(), ref)
    ptr, n = feng.post_dev()
    assert n == d["nsnps"] * K * 3

    def post_host():
        import torch
        from cramore_b200.fmx_dist import posterior_view
        feng.synchronize()
        return posterior_view(feng, torch.device("cuda", 0)).cpu().numpy().reshape(d["nsnps"], K, 3).copy()

    full = post_host()
    total = np.zeros_like(full)
    for b, e in ((0, 500), (500, 1200)):
        feng.set_shard(0, d["nbcs"], b, e)
        feng.mstep(assign)
        feng.posteriors_dev(True)
        part = post_host()
        assert np.all(part[:b] == 0) and np.all(part[e:] == 0)
        total += part
    assert np.array_equal(total, full)
    feng.set_shard(0, d["nbcs"], 0, d["nsnps"])


def test_pair_distances_match_oracle(feng):
    """ccu_fmx_pair_distances (sparse, sorted) against the dense dropDs matrix of the oracle
    (cmd_cram_freemuxlet.cpp:172-221): same pairs, exact counts, log-likelihood sums to 1e-12."""
    from tests.refcase import sparse_pairs
    d = small_c4(nbcs=180, nsnps=900, rbar=120)
    upload(feng, d, 4)
    got = feng.pair_distances()
    ref = sparse_pairs(pyoracle.fmx_pair_dist(d["cell_ptr"], d["snp_idx"], d["plp"], d["af"], d["nsnps"]))
    assert len(got) == len(ref) and feng.pair_stats["records"] == int(ref["nsnps"].sum())
    for f in ("id1", "id2", "nsnps", "nread1", "nread2"):
        assert np.array_equal(got[f], ref[f]), f
    np.testing.assert_allclose(got["llk0"], ref["llk0"], rtol=1e-12)
    np.testing.assert_allclose(got["llk2"], ref["llk2"], rtol=1e-12)
    assert np.all(np.diff(got["id1"].astype(np.int64) * d["nbcs"] + got["id2"]) > 0)
    # device-side compaction of the pairs that can vote
    for thres in (5.41, 1.0):
        sel = feng.voting_pairs(thres)
        want = got[np.abs(got["llk2"] - got["llk0"]) > thres]
        assert np.array_equal(sel, want) and (thres > 2 or len(sel) > 0)


@pytest.mark.parametrize("name", ["fmx_init_k3", "fmx_init_k4_partial"])
def test_initial_clustering_from_gpu_distances_equals_reference(feng, name):
    """GPU pair distances + host votes reproduce the reference's own .clust0 (its rand() sequence included)."""
    from cramore_b200.engine import initial_clusters
    from tests import refcase
    fx = refcase.load(name)
    d, plp, af, prm = refcase.fmx_init_inputs(fx)
    K = fx["spec"]["K"]
    feng.configure_fmx(K, 0.5, prm["geno_error"])
    feng.upload_fmx(d["cell_ptr"], d["snp_idx"], d["read_ptr"], d["al"], d["bq"], af)
    l0, l2 = feng.lmix()
    pairs = feng.pair_distances()
    for plist in (pairs, feng.voting_pairs(prm["bf_thres"])):  # the full list and the compacted one vote alike
        clusts = initial_clusters(l2 - l0, plist, K, bf_thres=prm["bf_thres"], frac_init_clust=prm["frac_init_clust"],
                                  iter_init=10, keep_init_missing=prm["keep_init_missing"], seed=1)
        assert clusts.tolist() == fx["clust0"]


_PEER_WORKER = r"""
import os, sys
import numpy as np
import torch
import torch.distributed as dist
sys.path.insert(0, %(root)r)
from cramore_b200 import FreemuxEngine
from cramore_b200.fmx_dist import device_views, posterior_view, run_em_distributed, snp_slabs
from cramore_b200.shard import balanced_ranges
from cramore_b200.synth import make_dataset

dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
ndev = torch.cuda.device_count()
dev = torch.device("cuda", rank %% ndev)  # one GPU per rank where the box has them (stores cross NVLink); else shared
torch.cuda.set_device(dev)
torch.cuda.set_stream(torch.cuda.Stream(device=dev))
K, NITER = 5, 4                          # odd K: a slab may start at an odd element (the 8-byte store path)
d = make_dataset("C4", nbcs=300, nsnps=2001, rbar=120, nv=5)
init = (d["s1"] %% K).astype(np.int32)
init[d["is_dbl"]] = -1
eng = FreemuxEngine(rank %% ndev)
eng.set_stream(torch.cuda.current_stream().cuda_stream)
eng.configure_fmx(K, 0.5, 0.05)
eng.upload_fmx(d["cell_ptr"], d["snp_idx"], d["read_ptr"], d["al"], d["bq"], d["af"])
stats, assign = device_views(eng, dev)
post = posterior_view(eng, dev)
ranges = balanced_ranges(d["cell_ptr"], world)
slabs = snp_slabs(d["snp_idx"], d["nsnps"], world)
out = {}
for mode in ("push", "push-nccl-ordered", "posteriors"):
    loc, full = run_em_distributed(eng, stats, assign, ranges, slabs, init, niter=NITER, post=post,
                                   exchange=mode.split("-")[0], geno_error_every_iter=True,
                                   push_barrier="nccl" if mode.endswith("ordered") else "kernel")
    torch.cuda.synchronize()
    out[mode] = (full.copy(), stats.cpu().numpy().copy())
assert np.array_equal(out["push"][0], out["push-nccl-ordered"][0]) and np.array_equal(out["push"][1], out["push-nccl-ordered"][1])
assert np.array_equal(out["push"][0], out["posteriors"][0]), "results differ between the exchanges"
assert np.array_equal(out["push"][1], out["posteriors"][1]), "cluster statistics differ between the exchanges"
if rank == 0:
    ref = FreemuxEngine(0)
    ref.configure_fmx(K, 0.5, 0.05)
    ref.upload_fmx(d["cell_ptr"], d["snp_idx"], d["read_ptr"], d["al"], d["bq"], d["af"])
    want, _ = ref.run_em(init, niter=NITER, geno_error_every_iter=True)
    assert np.array_equal(want, out["push"][0]), "peer-memory exchange differs from the single-process loop"
    assert np.array_equal(ref.clusters()["gls"].reshape(-1, 9), out["push"][1].reshape(-1, 12)[:, :9]), "cluster pileups differ"
    assert (want["type"] == 0).sum() > 100
    ref.close()
dist.barrier()
eng.close()
dist.destroy_process_group()
print("rank", rank, "ok")
"""


def test_peer_memory_exchange_two_processes(built, tmp_path):
    """The CUDA-IPC exchange of the multi-GPU freemuxlet loop (the M-step kernel storing its slab's posteriors and the
    decision kernel its assignments into the other rank's arrays, the ranks ordered by epoch flags inside those
    kernels or by a collective between them) with two processes: on two GPUs where the box has them (the stores then
    cross NVLink), else both on the one GPU.  Bit-identical to the all-reduce exchange and to the single-process
    loop.  gloo carries the host-side plumbing (NCCL refuses two ranks on one device)."""
    import subprocess
    import sys
    script = tmp_path / "peer_worker.py"
    script.write_text(_PEER_WORKER % {"root": os.path.dirname(os.path.dirname(os.path.abspath(__file__)))})
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                        "--master-addr", "127.0.0.1", "--master-port", "29733", str(script)],
                       capture_output=True, text=True, timeout=300, env=env)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-3000:]
    assert r.stdout.count("ok") == 2


_TIMEOUT_WORKER = r"""
import os, sys, time
import numpy as np
import torch
import torch.distributed as dist
sys.path.insert(0, %(root)r)
from cramore_b200 import FreemuxEngine
from cramore_b200.synth import make_dataset

dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
ndev = torch.cuda.device_count()
torch.cuda.set_device(rank %% ndev)
K = 4
d = make_dataset("C4", nbcs=200, nsnps=1000, rbar=80, nv=4)
eng = FreemuxEngine(rank %% ndev)
eng.configure_fmx(K, 0.5, 0.0)
eng.upload_fmx(d["cell_ptr"], d["snp_idx"], d["read_ptr"], d["al"], d["bq"], d["af"])
half = d["nbcs"] // 2
eng.set_shard(rank * half, (rank + 1) * half, rank * 500, (rank + 1) * 500)
everyone = [None] * world
dist.all_gather_object(everyone, eng.peer_handles())
eng.peer_open(rank, everyone)
if rank == 0:
    # rank 1 never launches its M-step: the E-step's wait at its head must give up (CCU_FMX_BARRIER_TIMEOUT_S = 1),
    # raise the flag, and every later kernel of the loop must return at once instead of computing on stale data
    t0 = time.time()
    eng.mstep_fused_dev(False, False, True)
    eng.estep_fused_dev(True)
    eng.mstep_fused_dev(False, False, True)
    eng.estep_fused_dev(True)
    eng.synchronize()
    dt = time.time() - t0
    assert eng.rank_barrier_failed(), "the wait did not time out"
    msg = eng.last_error()
    assert "rank 0" in msg and "ranks behind: 1" in msg and "CCU_FMX_BARRIER_TIMEOUT_S" in msg, msg
    assert 0.5 < dt < 20.0, dt
    print("timeout reported after %%.1f s: %%s" %% (dt, msg))
dist.barrier()
eng.close()
dist.destroy_process_group()
print("rank", rank, "ok")
"""


def test_peer_wait_times_out_instead_of_hanging(built, tmp_path):
    """A rank whose peer never arrives: the kernel waiting at its head gives up after CCU_FMX_BARRIER_TIMEOUT_S, raises
    the error flag, the remaining kernels of the loop return at once, and the host gets a message naming the rank,
    the epoch and the ranks that were behind."""
    import subprocess
    import sys
    script = tmp_path / "timeout_worker.py"
    script.write_text(_TIMEOUT_WORKER % {"root": os.path.dirname(os.path.dirname(os.path.abspath(__file__)))})
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", CCU_FMX_BARRIER_TIMEOUT_S="1")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                        "--master-addr", "127.0.0.1", "--master-port", "29735", str(script)],
                       capture_output=True, text=True, timeout=300, env=env)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-3000:]
    assert r.stdout.count("ok") == 2 and "timeout reported" in r.stdout


@pytest.mark.parametrize("nhandles", [2, 3])
@pytest.mark.parametrize("K", [6, 20])  # 20: every entry by the bilinear form, so the peers get every posterior triple
def test_in_process_multi_handle_em_is_bitwise_single(built, nhandles, K):
    """ccu_fmx_run_em_multi -- one host thread per handle, push kernels and the flag barrier between handles of ONE
    process -- with several handles on this box's GPU (on a multi-GPU box: one per device): results and the stitched
    cluster pileups are bit-identical to the single-handle loop."""
    import torch
    from cramore_b200 import FreemuxEngine
    from cramore_b200.engine import run_em_multi
    ndev = torch.cuda.device_count()
    NITER = 5
    d = make_dataset("C4", nbcs=500, nsnps=2500, rbar=120, nv=6)
    init = (d["s1"] % K).astype(np.int32)
    init[d["is_dbl"]] = -1
    engs = []
    try:
        for i in range(nhandles):
            e = FreemuxEngine(i % ndev)
            e.configure_fmx(K, 0.4, 0.03)
            e.upload_fmx(d["cell_ptr"], d["snp_idx"], d["read_ptr"], d["al"], d["bq"], d["af"])
            engs.append(e)
        res, clusters = run_em_multi(engs, d["cell_ptr"], d["snp_idx"], init, niter=NITER, want_clusters=True)
        ref = FreemuxEngine(0)
        engs.append(ref)
        ref.configure_fmx(K, 0.4, 0.03)
        ref.upload_fmx(d["cell_ptr"], d["snp_idx"], d["read_ptr"], d["al"], d["bq"], d["af"])
        want, _ = ref.run_em(init, niter=NITER)
        assert np.array_equal(res, want)
        assert (want["type"] == 0).sum() > 200
        wc = ref.clusters()
        for f in wc.dtype.names:
            assert np.array_equal(clusters[f], wc[f]), f
    finally:
        for e in engs:
            e.close()
