"""-m gpu: the `cramore` command line (cramore_b200/host) end to end -- digital pileup + VCF files in, output
files out -- against the files the REFERENCE's own compiled commands wrote for the same inputs
(tests/golden/ref_*.json)."""
import gzip
import os
import subprocess

import numpy as np
import pytest

from cramore_b200.synth import write_plp, write_vcf
from tests import refcase

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CRAMORE = os.path.join(ROOT, "cramore_b200", "host", "cramore")


def _numeric_row_match(a, b, rtol_llk=0.011):
    """Same row up to the last printed digit of a %.2lf / %.2lg field (1e-9-level differences can flip a rounding)."""
    fa, fb = a.rstrip("\n").split("\t"), b.rstrip("\n").split("\t")
    if len(fa) != len(fb):
        return False
    for x, y in zip(fa, fb):
        if x == y:
            continue
        try:
            if abs(float(x) - float(y)) > rtol_llk * max(1.0, abs(float(y)) * 0.01):
                return False
        except ValueError:
            return False
    return True


def _compare_best(mine, ref):
    """.best of this build against the reference's: the same bytes.  The engine's likelihoods agree with the
    reference's to ~1e-13; the rows are identical as text because the hypotheses they name are re-scored on the host in
    the reference's own operation order (ccu_demux_refine), which also settles the orientation of alpha = 0.5 pairs."""
    assert len(mine) == len(ref)
    for a, b in zip(mine, ref):
        assert a == b, (a, b)


@pytest.mark.parametrize("name", refcase.DEMUX_CASES)
def test_cli_demuxlet_best_file(built, tmp_path, name):
    fx = refcase.load(name)
    d = refcase.dataset(fx)
    spec = fx["spec"]
    pre = str(tmp_path / "in")
    write_plp(d, pre)
    write_vcf(d, str(tmp_path / "full.vcf"), with_r2=spec.get("with_r2", False))
    vcf = str(tmp_path / "in.vcf")
    with open(vcf, "w") as f:
        snp = 0
        for ln in open(str(tmp_path / "full.vcf")):
            if ln.startswith("#"):
                f.write(ln)
                continue
            if not spec.get("drop_every") or snp % spec["drop_every"] != 0:
                f.write(ln)
            snp += 1
    args = [CRAMORE, "demuxlet", "--plp", pre, "--vcf", vcf, "--field", spec.get("field", "GP"), "--out", str(tmp_path / "out")] + spec["args"]
    for a in (spec["alphas"] or []):
        args += ["--alpha", repr(float(a))]
    r = subprocess.run(args, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-3000:]
    # binary sidecar cache: --gpu-dump-inputs once, then --gpu-load-inputs without --plp/--vcf gives the same file
    dump = str(tmp_path / "cache")
    r2 = subprocess.run(args + ["--gpu-dump-inputs", dump], capture_output=True, text=True)
    assert r2.returncode == 0, r2.stderr[-3000:]
    keep = [a for a in args if a not in ("--plp", pre, "--vcf", vcf)]
    keep[keep.index(str(tmp_path / "out"))] = str(tmp_path / "out_cached")
    r3 = subprocess.run(keep + ["--gpu-load-inputs", dump], capture_output=True, text=True)
    assert r3.returncode == 0, r3.stderr[-3000:]
    assert open(str(tmp_path / "out_cached.best")).read() == open(str(tmp_path / "out.best")).read()
    mine = open(str(tmp_path / "out.best")).read().splitlines(keepends=True)
    ref = fx["best_text"].splitlines(keepends=True)
    _compare_best(mine, ref)


@pytest.mark.parametrize("name", refcase.FMX_CASES)
def test_cli_freemuxlet_files(built, tmp_path, name):
    fx = refcase.load(name)
    d = refcase.dataset(fx)
    K = fx["spec"]["K"]
    pre = str(tmp_path / "in")
    bcs = write_plp(d, pre)
    with open(str(tmp_path / "init.txt"), "w") as f:
        for b, c in zip(bcs, fx["init"]):
            f.write("%s\t%d\n" % (b, c))
    args = [CRAMORE, "freemuxlet", "--plp", pre, "--out", str(tmp_path / "out"), "--nsample", str(K), "--init-cluster",
            str(tmp_path / "init.txt"), "--iter-init", "0", "--keep-init-missing"] + fx["spec"]["args"]
    r = subprocess.run(args, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-3000:]
    mine = gzip.open(str(tmp_path / "out.clust1.samples.gz"), "rt").read().splitlines(keepends=True)
    ref = fx["samples_text"].splitlines(keepends=True)
    assert len(mine) == len(ref) and mine[0] == ref[0]
    exact = sum(a == b for a, b in zip(mine, ref))
    for a, b in zip(mine[1:], ref[1:]):
        assert a == b or _numeric_row_match(a, b), (a, b)
    assert exact >= 0.95 * len(ref)
    # .lmix and the cluster VCF exist with the reference's headers
    lm = open(str(tmp_path / "out.lmix")).read().splitlines()
    assert lm[0] == "INT_ID\tBARCODE\tNSNPs\tNREADs\tDBL.LLK\tSNG.LLK\tLOG.BF\tBFpSNP" and len(lm) == d["nbcs"] + 1
    ref_lm = [r for r in fx["lmix_raw"] if len(r) == 8]
    for ln, rr in zip(lm[1:], ref_lm):
        f = ln.split("\t")
        assert f[1] == rr[1] and abs(float(f[4]) - float(rr[4])) <= 0.011 and abs(float(f[5]) - float(rr[5])) <= 0.011
    vcf = gzip.open(str(tmp_path / "out.clust1.vcf.gz"), "rt").read().splitlines()
    body = [ln for ln in vcf if not ln.startswith("#")]
    observed = len(np.unique(d["snp_idx"]))
    assert len(body) == observed and vcf[2] == "##source=cramore-freemuxlet"
    for ln, site in zip(body, fx["vcf_sites"]):
        f = ln.split("\t")
        assert int(f[1]) == site["pos"] and len(f) == 9 + K
        for k in range(K):
            gt, gq, dp, ad, pl, gp = f[9 + k].split(":")
            c = site["clusters"][k]
            assert int(dp) == c[3] and ad == "%d,%d" % (c[4], c[5]) and pl == "%d,%d,%d" % (c[6], c[7], c[8])


@pytest.mark.parametrize("name", refcase.FMX_INIT_CASES)
def test_cli_freemuxlet_initial_clustering(built, tmp_path, name):
    """`cramore freemuxlet` without --init-cluster (or with it and the sweeps on): the GPU pair distances and the
    host votes give the reference's own .clust0.samples.gz, and the EM ends in the reference's .clust1.samples.gz."""
    fx = refcase.load(name)
    d = refcase.dataset(fx)
    K = fx["spec"]["K"]
    pre = str(tmp_path / "in")
    bcs = write_plp(d, pre)
    args = [CRAMORE, "freemuxlet", "--plp", pre, "--out", str(tmp_path / "out"), "--nsample", str(K), "--aux-files"] + \
        fx["spec"]["args"]
    if fx["init"] is not None:
        with open(str(tmp_path / "init.txt"), "w") as f:
            for b, c in zip(bcs, fx["init"]):
                if c >= 0:
                    f.write("%s\t%d\n" % (b, c))
        args += ["--init-cluster", str(tmp_path / "init.txt")]
    r = subprocess.run(args, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-3000:]
    c0 = gzip.open(str(tmp_path / "out.clust0.samples.gz"), "rt").read().splitlines()
    assert c0[0] == "INT_ID\tBARCODE\tCLUST0"
    assert [int(ln.split("\t")[2]) for ln in c0[1:]] == fx["clust0"]
    if fx["init"] is None:
        # <out>.ldist.gz: the pair rows of the greedy pass in the reference's order (dense, zero rows included)
        ld = gzip.open(str(tmp_path / "out.ldist.gz"), "rt").read().splitlines()
        assert ld[0] == "ID1\tID2\tNSNP\tREAD1\tREAD2\tREADMIN\tLLK0\tLLK2\tLDIFF\tDIFF.SNP"
        want = ["%d\t%d\t%d\t%d\t%d\t%d\t%.2f\t%.2f\t%.2f\t%.4f" % (tuple(int(x) for x in r[:6]) + tuple(float(x) for x in r[6:]))
                for r in fx["ldist_raw"]]
        assert len(ld) - 1 == len(want) > 500
        for a, b in zip(ld[1:], want):
            assert a == b or _numeric_row_match(a, b), (a, b)
        assert sum(a == b for a, b in zip(ld[1:], want)) >= 0.99 * len(want)
    mine = gzip.open(str(tmp_path / "out.clust1.samples.gz"), "rt").read().splitlines(keepends=True)
    ref = fx["samples_text"].splitlines(keepends=True)
    assert len(mine) == len(ref) and mine[0] == ref[0]
    for a, b in zip(mine[1:], ref[1:]):
        assert a == b or _numeric_row_match(a, b), (a, b)
    assert sum(a == b for a, b in zip(mine, ref)) >= 0.95 * len(ref)


@pytest.mark.parametrize("name", refcase.FMX2_CASES)
def test_cli_freemux2_files(built, tmp_path, name):
    """`cramore freemux2` against the reference's own freemux2 run: greedy initial clusters (.clust0), .lmix and the
    final .clust1.samples.gz (geno_error on every iteration, early stop)."""
    fx = refcase.load(name)
    d = refcase.dataset(fx)
    K = fx["spec"]["K"]
    pre = str(tmp_path / "in")
    bcs = write_plp(d, pre)
    args = [CRAMORE, "freemux2", "--plp", pre, "--out", str(tmp_path / "out"), "--nsample", str(K), "--aux-files",
            "--seed", "7"] + fx["spec"]["args"]
    if fx["init"] is not None:
        with open(str(tmp_path / "init.txt"), "w") as f:
            for b, c in zip(bcs, fx["init"]):
                if c >= 0:
                    f.write("%s\t%d\n" % (b, c))
        args += ["--init-cluster", str(tmp_path / "init.txt")]
    r = subprocess.run(args, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-3000:]
    c0 = gzip.open(str(tmp_path / "out.clust0.samples.gz"), "rt").read().splitlines()
    assert [int(ln.split("\t")[2]) for ln in c0[1:]] == fx["clust0"]
    mine = gzip.open(str(tmp_path / "out.clust1.samples.gz"), "rt").read().splitlines(keepends=True)
    ref = fx["samples_text"].splitlines(keepends=True)
    assert len(mine) == len(ref) and mine[0] == ref[0]
    for a, b in zip(mine[1:], ref[1:]):
        assert a == b or _numeric_row_match(a, b), (a, b)
    assert sum(a == b for a, b in zip(mine, ref)) >= 0.95 * len(ref)
    lm, rlm = open(str(tmp_path / "out.lmix")).read().splitlines(), fx["lmix_text"].splitlines()
    assert len(lm) == len(rlm) and lm[0] == rlm[0]
    for a, b in zip(lm[1:], rlm[1:]):
        assert a == b or _numeric_row_match(a, b), (a, b)


def test_cli_demuxlet_sam_input_matches_plp_input(built, tmp_path):
    """`--sam` (BGZF/BAM read by cramore_b200/host/bam_io.h, joined to the VCF on the host) scores the same droplets
    as `--plp` on the same reads.  NUM.SNPS / NUM.READS count the SNPs the VCF filter dropped on the --plp path only
    (the reference keeps them in the pileup store, cmd_cram_demuxlet.cpp:337-455, but never loads them from a BAM)."""
    from cramore_b200.synth import make_dataset, write_bam
    d = make_dataset("C1", nbcs=60, nsnps=300, rbar=150, nv=6, seed=29, with_barcodes=True)
    pre = str(tmp_path / "in")
    write_plp(d, pre)
    write_vcf(d, pre + ".vcf", with_r2=True, lead_decoy=True)
    write_bam(d, pre + ".bam", duplicate_every=7)
    rows = {}
    for mode, src in (("sam", ["--sam", pre + ".bam"]), ("plp", ["--plp", pre])):
        out = str(tmp_path / mode)
        r = subprocess.run([CRAMORE, "demuxlet"] + src + ["--vcf", pre + ".vcf", "--out", out, "--alpha", "0", "--alpha", "0.5"],
                           capture_output=True, text=True)
        assert r.returncode == 0, r.stderr[-3000:]
        rows[mode] = open(out + ".best").read().splitlines()
    assert rows["sam"][0] == rows["plp"][0] and len(rows["sam"]) == len(rows["plp"]) == 61
    for a, b in zip(rows["sam"][1:], rows["plp"][1:]):
        fa, fb = a.split("\t"), b.split("\t")
        assert fa[:2] == fb[:2] and int(fa[2]) <= int(fb[2]) and int(fa[3]) <= int(fb[3])
        assert _numeric_row_match("\t".join(fa[4:]), "\t".join(fb[4:])), (a, b)


REF_BIN = os.path.join(ROOT, "oracle", "_ref", "ref_cramore")
live = pytest.mark.skipif(not os.path.exists(REF_BIN), reason="oracle/_ref/ref_cramore not built")


@live
# CCU_LIVE_DEMUX_CASES / CCU_LIVE_FMX_CASES widen the random sweeps against the compiled reference (default 8 / 4)
@pytest.mark.parametrize("seed", range(int(os.environ.get("CCU_LIVE_DEMUX_CASES", "8"))))
def test_cli_demuxlet_equals_live_reference(built, tmp_path, seed):
    """`cramore demuxlet` on the GPU against the reference's own compiled engine (oracle/_ref/ref_cramore) run on the
    same files, for randomly drawn pool sizes, alpha grids, priors and error settings."""
    from cramore_b200.synth import make_dataset
    from tests.test_reference_live import _complement_free_grid
    rng = np.random.default_rng(3000 + seed)
    nv = int(rng.integers(2, 40))
    d = make_dataset("C1", nbcs=int(rng.integers(10, 60)), nsnps=int(rng.integers(100, 600)), rbar=int(rng.integers(20, 200)),
                     nv=nv, seed=7000 + seed, with_barcodes=True)
    alphas = _complement_free_grid(rng)
    common = ["--doublet-prior", repr(float(rng.choice([0.1, 0.5, 0.8]))), "--geno-error-offset",
              repr(float(rng.choice([0.0, 0.01, 0.1]))), "--geno-error-coeff", repr(float(rng.choice([0.0, 0.5]))),
              "--field", str(rng.choice(["GP", "GT"]))]
    for a in alphas:
        common += ["--alpha", repr(a)]
    pre = str(tmp_path / "in")
    write_plp(d, pre)
    write_vcf(d, pre + ".vcf", with_r2=True)
    r = subprocess.run([REF_BIN, "demuxlet", "--plp", "in", "--vcf", "in.vcf", "--out", "ref"] + common, capture_output=True,
                       text=True, cwd=str(tmp_path))
    assert r.returncode == 0, r.stderr[-2000:]
    r = subprocess.run([CRAMORE, "demuxlet", "--plp", pre, "--vcf", pre + ".vcf", "--out", str(tmp_path / "mine")] + common,
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-2000:]
    _compare_best(open(str(tmp_path / "mine.best")).read().splitlines(keepends=True),
                  open(str(tmp_path / "ref.best")).read().splitlines(keepends=True))


@live
@pytest.mark.parametrize("seed", range(int(os.environ.get("CCU_LIVE_FMX_CASES", "4"))))
def test_cli_freemuxlet_equals_live_reference(built, tmp_path, seed):
    from cramore_b200.synth import make_dataset
    rng = np.random.default_rng(4000 + seed)
    K = int(rng.integers(2, 9))
    d = make_dataset("C4", nbcs=int(rng.integers(40, 120)), nsnps=int(rng.integers(200, 700)), rbar=int(rng.integers(40, 120)),
                     nv=int(rng.integers(2, 7)), seed=8000 + seed, with_barcodes=True)
    pre = str(tmp_path / "in")
    bcs = write_plp(d, pre)
    init = (d["s1"] % K).astype(np.int32)
    init[d["is_dbl"]] = -1
    init[rng.random(len(init)) < 0.3] = -1
    with open(str(tmp_path / "init.txt"), "w") as f:
        for b, c in zip(bcs, init):
            f.write("%s\t%d\n" % (b, c))
    common = ["--nsample", str(K), "--init-cluster", str(tmp_path / "init.txt"), "--iter-init", "0", "--keep-init-missing",
              "--geno-error", repr(float(rng.choice([0.0, 0.05]))), "--doublet-prior", repr(float(rng.choice([0.3, 0.5])))]
    r = subprocess.run([REF_BIN, "freemuxlet", "--plp", "in", "--out", "ref"] + common, capture_output=True, text=True,
                       cwd=str(tmp_path))
    assert r.returncode == 0, r.stderr[-2000:]
    r = subprocess.run([CRAMORE, "freemuxlet", "--plp", pre, "--out", str(tmp_path / "mine")] + common, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-2000:]
    mine = gzip.open(str(tmp_path / "mine.clust1.samples.gz"), "rt").read().splitlines(keepends=True)
    ref = gzip.open(str(tmp_path / "ref.clust1.samples.gz"), "rt").read().splitlines(keepends=True)
    assert len(mine) == len(ref) and mine[0] == ref[0]
    for a, b in zip(mine[1:], ref[1:]):
        assert a == b or _numeric_row_match(a, b), (a, b)
    assert sum(a == b for a, b in zip(mine, ref)) >= 0.9 * len(ref)
    assert open(str(tmp_path / "mine.lmix")).read().splitlines()[0] == open(str(tmp_path / "ref.lmix")).read().splitlines()[0]


@pytest.mark.parametrize("name", ["fmx_k4", "fmx_k6_prior"])
def test_cli_freemuxlet_over_several_handles(built, tmp_path, name):
    """`cramore freemuxlet --gpu-devices a,b,c`: the EM loop sharded over several engine handles (ccu_fmx_run_em_multi,
    peer-memory exchange) writes the files of the single-device run, which equal the reference's."""
    import torch
    fx = refcase.load(name)
    d = refcase.dataset(fx)
    K = fx["spec"]["K"]
    pre = str(tmp_path / "in")
    bcs = write_plp(d, pre)
    with open(str(tmp_path / "init.txt"), "w") as f:
        for b, c in zip(bcs, fx["init"]):
            f.write("%s\t%d\n" % (b, c))
    ndev = torch.cuda.device_count()
    devs = ",".join(str(i % ndev) for i in range(3))
    outs = {}
    for tag, extra in (("one", []), ("many", ["--gpu-devices", devs])):
        out = str(tmp_path / tag)
        r = subprocess.run([CRAMORE, "freemuxlet", "--plp", pre, "--out", out, "--nsample", str(K), "--init-cluster",
                            str(tmp_path / "init.txt"), "--iter-init", "0", "--keep-init-missing"] + fx["spec"]["args"] + extra,
                           capture_output=True, text=True)
        assert r.returncode == 0, r.stderr[-3000:]
        vcf = [ln for ln in gzip.open(out + ".clust1.vcf.gz", "rt").read().splitlines() if not ln.startswith("##fileDate")]
        outs[tag] = (gzip.open(out + ".clust1.samples.gz", "rt").read(), vcf, open(out + ".lmix").read())
    assert outs["one"] == outs["many"]
    ref = fx["samples_text"].splitlines(keepends=True)
    mine = outs["many"][0].splitlines(keepends=True)
    assert len(mine) == len(ref)
    for a, b in zip(mine, ref):
        assert a == b or _numeric_row_match(a, b), (a, b)


REF_GPU_BIN = os.path.join(ROOT, "oracle", "_ref", "ref_cramore_gpu")


@pytest.mark.skipif(not os.path.exists(REF_GPU_BIN), reason="oracle/_ref/ref_cramore_gpu not built (make -C oracle ref_gpu)")
@pytest.mark.parametrize("name", refcase.DEMUX_CASES)
def test_patched_reference_command_runs_on_the_engine(built, tmp_path, name):
    """THE drop-in: the reference's own cmdCramDemuxlet -- its option parser, load_from_plp, VCF join and genotype-error
    mixing, compiled from /root/reference unmodified up to line 426 -- with the likelihood loop replaced by the patch
    of INTEGRATION.md section 1 (oracle/ref_shim/demuxlet_gpu_patch.inc) and linked against libcramore_cuda.so, writes the
    .best file the stock reference writes."""
    fx = refcase.load(name)
    d = refcase.dataset(fx)
    spec = fx["spec"]
    write_plp(d, str(tmp_path / "in"))
    write_vcf(d, str(tmp_path / "full.vcf"), with_r2=spec.get("with_r2", False))
    with open(str(tmp_path / "in.vcf"), "w") as f:
        snp = 0
        for ln in open(str(tmp_path / "full.vcf")):
            if ln.startswith("#"):
                f.write(ln)
                continue
            if not spec.get("drop_every") or snp % spec["drop_every"] != 0:
                f.write(ln)
            snp += 1
    args = [REF_GPU_BIN, "demuxlet", "--plp", "in", "--vcf", "in.vcf", "--field", spec.get("field", "GP"), "--out", "out"] + spec["args"]
    for a in (spec["alphas"] or []):
        args += ["--alpha", repr(float(a))]
    r = subprocess.run(args, capture_output=True, text=True, cwd=str(tmp_path))
    assert r.returncode == 0, r.stderr[-3000:]
    _compare_best(open(str(tmp_path / "out.best")).read().splitlines(keepends=True),
                  fx["best_text"].splitlines(keepends=True))


@pytest.mark.skipif(not os.path.exists(REF_GPU_BIN), reason="oracle/_ref/ref_cramore_gpu not built (make -C oracle ref_gpu)")
@pytest.mark.parametrize("name", refcase.FMX_CASES + refcase.FMX_INIT_CASES)
def test_patched_reference_freemuxlet_runs_on_the_engine(built, tmp_path, name):
    """The reference's own cmdCramFreemuxlet -- options, load_from_plp, --init-cluster file, and its writers of
    .clust1.vcf.gz / .clust1.samples.gz, all compiled from /root/reference -- with lines 105-653 replaced by the engine
    calls of INTEGRATION.md section 2 (oracle/ref_shim/freemuxlet_gpu_patch.inc): same .clust0 / .clust1 files as the
    stock reference."""
    fx = refcase.load(name)
    d = refcase.dataset(fx)
    K = fx["spec"]["K"]
    bcs = write_plp(d, str(tmp_path / "in"))
    args = [REF_GPU_BIN, "freemuxlet", "--plp", "in", "--out", "out", "--nsample", str(K), "--aux-files"] + fx["spec"]["args"]
    if name in refcase.FMX_CASES:
        args += ["--init-cluster", "init.txt", "--iter-init", "0", "--keep-init-missing"]
    elif fx["init"] is not None:
        args += ["--init-cluster", "init.txt"]
    if fx["init"] is not None:
        with open(str(tmp_path / "init.txt"), "w") as f:
            for b, c in zip(bcs, fx["init"]):
                if c >= 0 or name in refcase.FMX_CASES:
                    f.write("%s\t%d\n" % (b, c))
    r = subprocess.run(args, capture_output=True, text=True, cwd=str(tmp_path))
    assert r.returncode == 0, r.stderr[-3000:]
    if "clust0" in fx:
        c0 = gzip.open(str(tmp_path / "out.clust0.samples.gz"), "rt").read().splitlines()
        assert [int(ln.split("\t")[2]) for ln in c0[1:]] == fx["clust0"]
    mine = gzip.open(str(tmp_path / "out.clust1.samples.gz"), "rt").read().splitlines(keepends=True)
    ref = fx["samples_text"].splitlines(keepends=True)
    assert len(mine) == len(ref) and mine[0] == ref[0]
    for a, b in zip(mine[1:], ref[1:]):
        assert a == b or _numeric_row_match(a, b), (a, b)
    assert sum(a == b for a, b in zip(mine, ref)) >= 0.9 * len(ref)
    vcf = [ln for ln in gzip.open(str(tmp_path / "out.clust1.vcf.gz"), "rt").read().splitlines() if not ln.startswith("#")]
    assert len(vcf) >= 40 and all(len(ln.split("\t")) == 9 + K for ln in vcf)


@pytest.mark.skipif(not os.path.exists(REF_GPU_BIN), reason="oracle/_ref/ref_cramore_gpu not built (make -C oracle ref_gpu)")
@pytest.mark.parametrize("name", refcase.FMX2_CASES)
def test_patched_reference_freemux2_runs_on_the_engine(built, tmp_path, name):
    """cmdCramFreemux2 of the reference with lines 106-607 replaced by engine calls (oracle/ref_shim/freemux2_gpu_patch.inc):
    greedy initial clusters, .lmix and the final files equal the stock reference's."""
    fx = refcase.load(name)
    d = refcase.dataset(fx)
    K = fx["spec"]["K"]
    bcs = write_plp(d, str(tmp_path / "in"))
    args = [REF_GPU_BIN, "freemux2", "--plp", "in", "--out", "out", "--nsample", str(K), "--aux-files", "--seed", "7"] + fx["spec"]["args"]
    if fx["init"] is not None:
        with open(str(tmp_path / "init.txt"), "w") as f:
            for b, c in zip(bcs, fx["init"]):
                if c >= 0:
                    f.write("%s\t%d\n" % (b, c))
        args += ["--init-cluster", "init.txt"]
    r = subprocess.run(args, capture_output=True, text=True, cwd=str(tmp_path))
    assert r.returncode == 0, r.stderr[-3000:]
    c0 = gzip.open(str(tmp_path / "out.clust0.samples.gz"), "rt").read().splitlines()
    assert [int(ln.split("\t")[2]) for ln in c0[1:]] == fx["clust0"]
    mine = gzip.open(str(tmp_path / "out.clust1.samples.gz"), "rt").read().splitlines(keepends=True)
    ref = fx["samples_text"].splitlines(keepends=True)
    assert len(mine) == len(ref) and mine[0] == ref[0]
    for a, b in zip(mine[1:], ref[1:]):
        assert a == b or _numeric_row_match(a, b), (a, b)
    lm, rlm = open(str(tmp_path / "out.lmix")).read().splitlines(), fx["lmix_text"].splitlines()
    assert len(lm) == len(rlm) and lm[0] == rlm[0]
    for a, b in zip(lm[1:], rlm[1:]):
        assert a == b or _numeric_row_match(a, b), (a, b)


def test_cli_demuxlet_over_several_devices(built, tmp_path):
    """`cramore demuxlet --gpu-devices a,b,c`: rows split into contiguous ranges, one engine handle and host thread per
    entry of the list (distinct GPUs where the box has them, else handles sharing the GPU), the genotype matrix uploaded
    once and copied between the handles (ccu_demux_copy_genotypes): the same .best, byte for byte, as on one device --
    which in turn is the reference's.  --gpu-timings leaves the machine-readable kernel times."""
    import json
    import torch
    fx = refcase.load("demux_grid_deep")
    d = refcase.dataset(fx)
    spec = fx["spec"]
    pre = str(tmp_path / "in")
    write_plp(d, pre)
    write_vcf(d, pre + ".vcf", with_r2=spec.get("with_r2", False))
    ndev = torch.cuda.device_count()
    common = [CRAMORE, "demuxlet", "--plp", pre, "--vcf", pre + ".vcf", "--field", spec.get("field", "GP")] + spec["args"]
    for a in (spec["alphas"] or []):
        common += ["--alpha", repr(float(a))]
    outs = {}
    for name, devs in (("one", "0"), ("three", ",".join(str(i % ndev) for i in range(3)))):
        r = subprocess.run(common + ["--out", str(tmp_path / name), "--gpu-devices", devs, "--gpu-timings",
                                     str(tmp_path / (name + ".json"))], capture_output=True, text=True)
        assert r.returncode == 0, r.stderr[-3000:]
        outs[name] = open(str(tmp_path / (name + ".best"))).read()
    assert outs["one"] == outs["three"] == fx["best_text"]
    tm = json.load(open(str(tmp_path / "three.json")))
    assert tm["devices"] == 3 and tm["droplets"] == d["nbcs"] and tm["ms_score"] > 0 and tm["kernel_launches"] >= 5


def test_cli_write_single(built, tmp_path):
    """<out>.single (cmd_cram_demuxlet.cpp:516, :552-563; the reference's own writer is commented out, so the check is
    against the oracle's per-sample singlet log-likelihoods llksAB[j][0][0] and a direct evaluation of llks00[0],
    :762-773): header, one row per (droplet, sample), LLK1 / LLK0 to the printed precision, posteriors summing to 1."""
    from oracle import pyoracle
    fx = refcase.load("demux_missing_gp_r2")
    d, alphas, prior, snp_err, has_gp = refcase.demux_inputs(fx)
    spec = fx["spec"]
    pre = str(tmp_path / "in")
    write_plp(d, pre)
    write_vcf(d, str(tmp_path / "full.vcf"), with_r2=spec.get("with_r2", False))
    with open(pre + ".vcf", "w") as f:
        snp = 0
        for ln in open(str(tmp_path / "full.vcf")):
            if ln.startswith("#"):
                f.write(ln)
                continue
            if not spec.get("drop_every") or snp % spec["drop_every"] != 0:
                f.write(ln)
            snp += 1
    args = [CRAMORE, "demuxlet", "--plp", pre, "--vcf", pre + ".vcf", "--out", str(tmp_path / "out"), "--write-single"] + spec["args"]
    for a in (spec["alphas"] or []):
        args += ["--alpha", repr(float(a))]
    r = subprocess.run(args, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-3000:]
    assert open(str(tmp_path / "out.best")).read() == fx["best_text"]
    lines = open(str(tmp_path / "out.single")).read().splitlines()
    assert lines[0] == "BARCODE\tSM_ID\tRD.TOTL\tRD.PASS\tRD.UNIQ\tN.SNP\tLLK1\tLLK0\tPOSTPRB"
    nv = d["nv"]
    assert len(lines) == 1 + d["nbcs"] * nv
    gps = pyoracle.mix_genotypes(d["gp"], snp_err, has_gp)
    _, llk, _ = pyoracle.demux_score(d["cell_ptr"], d["snp_idx"], d["read_ptr"], d["al"], d["bq"], gps, alphas, prior,
                                     has_gp=has_gp, want_llk=True)
    gp0 = gps.sum(axis=1) / nv
    order = sorted(range(d["nbcs"]), key=lambda i: d["barcodes"][i])
    cp, rp = d["cell_ptr"], d["read_ptr"]
    for rank, c in enumerate(order):
        llk0 = 0.0
        for e in range(cp[c], cp[c + 1]):
            s = d["snp_idx"][e]
            if not has_gp[s]:
                continue
            pg = pyoracle.demux_tile(d["al"][rp[e]:rp[e + 1]], d["bq"][rp[e]:rp[e + 1]], alphas)[:9].reshape(3, 3)
            llk0 += np.log(sum(gp0[s][l] * gp0[s][m] * pg[l, m] for l in range(3) for m in range(3)))
        post = 0.0
        for j in range(nv):
            f = lines[1 + rank * nv + j].split("\t")
            assert f[0] == d["barcodes"][c] and f[1] == d["sample_ids"][j] and int(f[5]) == cp[c + 1] - cp[c]
            assert int(f[4]) == rp[cp[c + 1]] - rp[cp[c]]
            assert abs(float(f[6]) - llk[c, j, 0, 0]) <= 1e-5 + 1e-9 * abs(llk[c, j, 0, 0])
            assert abs(float(f[7]) - llk0) <= 1e-5 + 1e-9 * abs(llk0)
            post += float(f[8])
        assert abs(post - 1.0) < 0.02  # %.3lg per value
