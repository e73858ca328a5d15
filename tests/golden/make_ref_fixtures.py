"""Generates tests/golden/ref_*.json: outputs of the REFERENCE's own compiled likelihood code
(oracle/_ref/ref_cramore = /root/reference/cmd_cram_demuxlet.cpp, cmd_cram_freemuxlet.cpp,
sc_drop_seq.cpp, ... compiled unmodified against the zlib I/O shim in oracle/ref_shim/) on synthetic
digital-pileup + VCF inputs written by cramore_b200.synth.  Runs only where /root/reference exists
(`make -C oracle ref` first).  Inputs are not stored: every case is re-created from its seed; the
fixture holds the case parameters and every value the reference passed to hprintf at full precision
(REF_SHIM_RAW=1 side files)."""
import gzip
import json
import os
import subprocess
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from cramore_b200.synth import make_dataset, write_plp, write_vcf  # noqa: E402

REF = os.path.join(ROOT, "oracle", "_ref", "ref_cramore")

DEMUX_CASES = {
    "demux_c1_small": dict(ds=dict(name="C1", nbcs=60, nsnps=400, rbar=60), alphas=None, args=[]),
    "demux_grid_deep": dict(ds=dict(name="C1", nbcs=30, nsnps=200, rbar=90, nv=5, seed=77),
                            alphas=[0, 0.1, 0.2, 0.3, 0.4, 0.5], args=["--doublet-prior", "0.3"]),
    "demux_missing_gp_r2": dict(ds=dict(name="C1", nbcs=40, nsnps=300, rbar=50, nv=6, seed=99), alphas=[0, 0.25, 0.5],
                                args=["--geno-error-offset", "0.05", "--geno-error-coeff", "0.5"], drop_every=7,
                                with_r2=True),
    # hard genotype calls (--field GT: posteriors exactly 0 / 1) without any genotype-error mixing: every likelihood is
    # a single tile element, mismatches cost ~23 nats per read
    "demux_hard_gt": dict(ds=dict(name="C1", nbcs=40, nsnps=300, rbar=70, nv=7, seed=123), alphas=[0, 0.25, 0.5],
                          args=["--geno-error-offset", "0"], field="GT"),
    # a grid that reaches beyond 0.5 and does not contain it: ordered pairs at every alpha.  No alpha together with its
    # complement (0.3 and 0.7: (j,k,0.3) and (k,j,0.7) are one hypothesis, told apart by rounding only, like the two
    # orientations at 0.5) and not 1.0 (the likelihood no longer depends on the first sample of the pair)
    "demux_alpha_beyond_half": dict(ds=dict(name="C1", nbcs=36, nsnps=260, rbar=80, nv=6, seed=321),
                                    alphas=[0, 0.3, 0.6, 0.9], args=["--doublet-prior", "0.2"]),
    # a subset of the VCF's samples (--sm): genotype columns, sample ids and the --min-mac filter follow the subset
    "demux_sample_subset": dict(ds=dict(name="C1", nbcs=36, nsnps=280, rbar=70, nv=9, seed=555), alphas=None,
                                args=["--sm", "SM001", "--sm", "SM004", "--sm", "SM005", "--sm", "SM008"]),
}
FMX_CASES = {
    "fmx_k4": dict(ds=dict(name="C4", nbcs=80, nsnps=600, rbar=80, nv=4), K=4, args=["--geno-error", "0.1"]),
    "fmx_k3_noerr": dict(ds=dict(name="C4", nbcs=50, nsnps=400, rbar=60, nv=3, seed=5), K=3, args=[]),
    # more clusters than the data has samples for (some stay nearly empty), a doublet prior away from 0.5
    "fmx_k6_prior": dict(ds=dict(name="C4", nbcs=120, nsnps=700, rbar=90, nv=5, seed=8), K=6,
                         args=["--geno-error", "0.02", "--doublet-prior", "0.3"]),
}


# initial clustering (cmd_cram_freemuxlet.cpp:166-346): no --init-cluster, or --init-cluster with the sweeps on
FMX_INIT_CASES = {
    "fmx_init_k3": dict(ds=dict(name="C4", nbcs=60, nsnps=300, rbar=80, nv=3, seed=21), K=3, args=[]),
    "fmx_init_k4_partial": dict(ds=dict(name="C4", nbcs=48, nsnps=260, rbar=90, nv=4, seed=31), K=4,
                                args=["--frac-init-clust", "0.7", "--bf-thres", "3.0", "--geno-error", "0.05"]),
    "fmx_init_k3_fromfile": dict(ds=dict(name="C4", nbcs=40, nsnps=240, rbar=80, nv=3, seed=41), K=3,
                                 args=["--keep-init-missing"], init_file=True),
}


# freemux2 (cmd_cram_freemux2.cpp): greedy cluster-pileup initialisation, geno_error on every iteration, early stop
FMX2_CASES = {
    "fmx2_k3_greedy": dict(ds=dict(name="C4", nbcs=60, nsnps=300, rbar=80, nv=3, seed=21), K=3, args=[]),
    "fmx2_k4_greedy_partial": dict(ds=dict(name="C4", nbcs=50, nsnps=280, rbar=90, nv=4, seed=33), K=4,
                                   args=["--frac-init-clust", "0.6", "--geno-error", "0.05", "--doublet-prior", "0.3"]),
    "fmx2_k3_fromfile": dict(ds=dict(name="C4", nbcs=40, nsnps=240, rbar=80, nv=3, seed=41), K=3, args=[], init_file=True),
}


def wanted(cases):
    """(name, spec) of the cases to (re)generate: all, or the ones named on the command line"""
    only = set(sys.argv[1:])
    return [(n, sp) for n, sp in cases.items() if not only or n in only]


def dataset(spec):
    ds = dict(spec["ds"])
    return make_dataset(ds.pop("name"), with_barcodes=True, **ds)


def read_raw(path):
    rows = []
    for line in open(path):
        line = line.rstrip("\n")
        if line:
            rows.append(line.split("\t"))
    return rows


def run(args, cwd):
    env = dict(os.environ, REF_SHIM_RAW="1")
    r = subprocess.run([REF] + args, capture_output=True, text=True, env=env, cwd=cwd)
    if r.returncode != 0:
        raise RuntimeError(r.stderr[-3000:])


def init_clusters(d, K):
    """Deterministic --init-cluster file content: the simulated sample for most droplets."""
    rng = np.random.default_rng(17)
    a = (d["s1"] % K).astype(int)
    a[d["is_dbl"]] = -1
    a[rng.random(len(a)) < 0.25] = -1
    return a


def main():
    if not os.path.exists(REF):
        raise SystemExit("build oracle/_ref first: make -C oracle ref (needs /root/reference)")
    for name, spec in wanted(DEMUX_CASES):
        d = dataset(spec)
        with tempfile.TemporaryDirectory() as tmp:
            write_plp(d, os.path.join(tmp, "in"))
            dv = d
            if spec.get("drop_every"):
                keep = np.arange(d["nsnps"]) % spec["drop_every"] != 0
                lines = open_vcf_lines(d, tmp, spec.get("with_r2", False))
                with open(os.path.join(tmp, "in.vcf"), "w") as f:
                    snp = 0
                    for ln in lines:
                        if ln.startswith("#"):
                            f.write(ln)
                        else:
                            if keep[snp]:
                                f.write(ln)
                            snp += 1
            else:
                write_vcf(dv, os.path.join(tmp, "in.vcf"), with_r2=spec.get("with_r2", False))
            args = ["demuxlet", "--plp", "in", "--vcf", "in.vcf", "--field", spec.get("field", "GP"), "--out", "out"] + spec["args"]
            for a in (spec["alphas"] or []):
                args += ["--alpha", repr(float(a))]
            run(args, tmp)
            best = open(os.path.join(tmp, "out.best")).read()
            raw = read_raw(os.path.join(tmp, "out.best.raw"))
        json.dump({"case": name, "spec": spec, "best_text": best, "best_raw": raw},
                  open(os.path.join(HERE, "ref_%s.json" % name), "w"))
        print(name, len(raw), "rows")
    for name, spec in wanted(FMX_CASES):
        d = dataset(spec)
        K = spec["K"]
        with tempfile.TemporaryDirectory() as tmp:
            bcs = write_plp(d, os.path.join(tmp, "in"))
            init = init_clusters(d, K)
            with open(os.path.join(tmp, "init.txt"), "w") as f:
                for b, c in zip(bcs, init):
                    f.write("%s\t%d\n" % (b, c))
            args = ["freemuxlet", "--plp", "in", "--out", "out", "--nsample", str(K), "--init-cluster", "init.txt",
                    "--iter-init", "0", "--keep-init-missing"] + spec["args"]
            run(args, tmp)
            out = {"case": name, "spec": spec, "init": init.tolist(),
                   "lmix_raw": read_raw(os.path.join(tmp, "out.lmix.raw")),
                   "samples_raw": read_raw(os.path.join(tmp, "out.clust1.samples.gz.raw")),
                   "samples_text": gzip.open(os.path.join(tmp, "out.clust1.samples.gz"), "rt").read(),
                   "vcf_sites": vcf_sites(read_raw(os.path.join(tmp, "out.clust1.vcf.gz.raw")), K)}
        json.dump(out, open(os.path.join(HERE, "ref_%s.json" % name), "w"))
        print(name, len(out["samples_raw"]), "rows")
    for name, spec in wanted(FMX_INIT_CASES):
        d = dataset(spec)
        K = spec["K"]
        with tempfile.TemporaryDirectory() as tmp:
            bcs = write_plp(d, os.path.join(tmp, "in"))
            args = ["freemuxlet", "--plp", "in", "--out", "out", "--nsample", str(K), "--aux-files"] + spec["args"]
            init = None
            if spec.get("init_file"):
                init = init_clusters(d, K)
                with open(os.path.join(tmp, "init.txt"), "w") as f:
                    for b, c in zip(bcs, init):
                        if c >= 0:
                            f.write("%s\t%d\n" % (b, c))
                args += ["--init-cluster", "init.txt"]
            run(args, tmp)
            ldist = read_raw(os.path.join(tmp, "out.ldist.gz.raw")) if init is None else []
            clust0 = [int(r[2]) for r in read_raw(os.path.join(tmp, "out.clust0.samples.gz.raw")) if len(r) == 3
                      and r[0].lstrip("-").isdigit()]
            out = {"case": name, "spec": spec, "init": None if init is None else init.tolist(),
                   "lmix_raw": read_raw(os.path.join(tmp, "out.lmix.raw")),
                   "ldist_raw": [r for r in ldist if r[0].isdigit()],
                   "clust0": clust0,
                   "samples_raw": read_raw(os.path.join(tmp, "out.clust1.samples.gz.raw")),
                   "samples_text": gzip.open(os.path.join(tmp, "out.clust1.samples.gz"), "rt").read()}
        json.dump(out, open(os.path.join(HERE, "ref_%s.json" % name), "w"))
        print(name, len(out["clust0"]), "droplets,", len(out["ldist_raw"]), "pair rows")
    freemux2_cases()


def freemux2_cases():
    for name, spec in wanted(FMX2_CASES):
        d = dataset(spec)
        K = spec["K"]
        with tempfile.TemporaryDirectory() as tmp:
            bcs = write_plp(d, os.path.join(tmp, "in"))
            args = ["freemux2", "--plp", "in", "--out", "out", "--nsample", str(K), "--aux-files", "--seed", "7"] + spec["args"]
            init = None
            if spec.get("init_file"):
                init = init_clusters(d, K)
                with open(os.path.join(tmp, "init.txt"), "w") as f:
                    for b, c in zip(bcs, init):
                        if c >= 0:
                            f.write("%s\t%d\n" % (b, c))
                args += ["--init-cluster", "init.txt"]
            run(args, tmp)
            clust0 = [int(r[2]) for r in read_raw(os.path.join(tmp, "out.clust0.samples.gz.raw")) if len(r) == 3
                      and r[0].lstrip("-").isdigit()]
            out = {"case": name, "spec": spec, "init": None if init is None else init.tolist(),
                   "lmix_raw": read_raw(os.path.join(tmp, "out.lmix.raw")),
                   "lmix_text": open(os.path.join(tmp, "out.lmix")).read(),
                   "clust0": clust0,
                   "samples_raw": read_raw(os.path.join(tmp, "out.clust1.samples.gz.raw")),
                   "samples_text": gzip.open(os.path.join(tmp, "out.clust1.samples.gz"), "rt").read()}
        json.dump(out, open(os.path.join(HERE, "ref_%s.json" % name), "w"))
        print(name, len(out["clust0"]), "droplets")


def vcf_sites(raw, K, limit=60):
    """clust1.vcf.gz as hprintf saw it (cmd_cram_freemuxlet.cpp:670-704): per site one call with
    (chrom, pos, ref, alt, af) followed by K calls with (gt1, gt2, gq, dp, nref, nalt, pl0..2, gp0..2)."""
    sites = []
    i = 0
    while i < len(raw) and len(sites) < limit:
        if len(raw[i]) == 5 and all(len(r) == 12 for r in raw[i + 1:i + 1 + K]):
            sites.append({"pos": int(raw[i][1]), "af": float(raw[i][4]),
                          "clusters": [[float(x) for x in r] for r in raw[i + 1:i + 1 + K]]})
            i += 1 + K
        else:
            i += 1
    return sites


def open_vcf_lines(d, tmp, with_r2):
    p = os.path.join(tmp, "full.vcf")
    write_vcf(d, p, with_r2=with_r2)
    return open(p).readlines()


if __name__ == "__main__":
    main()
