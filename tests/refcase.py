"""Loader for tests/golden/ref_*.json -- outputs of the reference's own compiled engine (see
tests/golden/make_ref_fixtures.py) -- and the matching synthetic inputs, re-created from their seeds."""
import json
import os

import numpy as np

from cramore_b200.synth import make_dataset

HERE = os.path.dirname(os.path.abspath(__file__))
DEMUX_CASES = ["demux_c1_small", "demux_grid_deep", "demux_missing_gp_r2", "demux_hard_gt", "demux_alpha_beyond_half",
               "demux_sample_subset"]
FMX_CASES = ["fmx_k4", "fmx_k3_noerr", "fmx_k6_prior"]
FMX_INIT_CASES = ["fmx_init_k3", "fmx_init_k4_partial", "fmx_init_k3_fromfile"]
FMX2_CASES = ["fmx2_k3_greedy", "fmx2_k4_greedy_partial", "fmx2_k3_fromfile"]

# columns of one .best row as passed to hprintf (cmd_cram_demuxlet.cpp:993-1013)
BEST_COLS = ["int_id", "barcode", "nsnps", "nreads", "type", "best1", "best2", "bestA", "bestLLK", "next1", "next2",
             "nextA", "nextLLK", "diffBestNext", "bestPP", "sngPP", "sBest", "sngBestLLK", "sNext", "sngNextLLK",
             "sngOnlyPP", "dBest1", "dBest2", "dBestAlpha", "dblBestLLK", "diffSngDbl"]
# cmd_cram_freemuxlet.cpp:709-711
FMX_COLS = ["int_id", "barcode", "nsnps", "nreads", "type", "jBest", "kBest", "bestLLK", "jNext", "kNext", "nextLLK",
            "diff", "bestPP", "sngPP", "sBest", "sngBestLLK", "sNext", "sngNextLLK", "sngOnlyPP", "dBest1", "dBest2",
            "dblBestLLK", "diffSngDbl"]


def load(name):
    return json.load(open(os.path.join(HERE, "golden", "ref_%s.json" % name)))


def dataset(fx):
    ds = dict(fx["spec"]["ds"])
    return make_dataset(ds.pop("name"), with_barcodes=True, **ds)


def demux_inputs(fx):
    """(dataset, alphas, doublet_prior, snp_err[nsnps], has_gp[nsnps] or None) as cmdCramDemuxlet would see them."""
    d = dataset(fx)
    spec = fx["spec"]
    args = spec["args"]

    def opt(flag, default):
        return float(args[args.index(flag) + 1]) if flag in args else default

    if "--sm" in args:  # sample subset, in VCF column order (bcf_filtered_reader.cpp:108-127)
        want = {args[i + 1] for i, a in enumerate(args) if a == "--sm"}
        cols = [i for i, sid in enumerate(d["sample_ids"]) if sid in want]
        d["gp"] = np.ascontiguousarray(np.asarray(d["gp"])[:, cols])
        d["geno"] = np.asarray(d["geno"])[:, cols]
        d["sample_ids"] = [d["sample_ids"][i] for i in cols]
        d["nv"] = len(cols)
    if spec.get("field") == "GT":  # parse_posteriors' GT branch with gt_error = 0 (sc_drop_seq.cpp:285): 0 / 1
        gp = np.zeros((d["nsnps"], d["nv"], 3), np.float32)
        np.put_along_axis(gp, np.asarray(d["geno"], np.int64)[:, :, None], np.float32(1.0), axis=2)
        d["gp"] = gp
    alphas = np.asarray(spec["alphas"] if spec["alphas"] else [0.0, 0.5], np.float64)  # cmd_cram_demuxlet.cpp:85-89
    off, coeff = opt("--geno-error-offset", 0.10), opt("--geno-error-coeff", 0.0)
    err = off
    if coeff > 0:  # sc_drop_seq.cpp:300-306 with R2=0.9 read as float
        err += (1 - off) * (1 - float(np.float32(0.9))) * coeff
    # BCFFilteredReader::read drops variants failing --min-mac 1 / --min-callrate 0.5
    # (cmd_cram_demuxlet.cpp:27-28, bcf_filtered_reader.cpp:627-638); load_from_plp then keeps the SNP
    # without genotypes (sc_drop_seq.cpp:278-282), i.e. gps == NULL
    ac = np.asarray(d["geno"], np.int64).sum(axis=1)
    an = 2 * d["nv"]
    has_gp = ((ac >= 1) & (an - ac >= 1)).astype(np.uint8)
    if spec.get("drop_every"):
        has_gp &= (np.arange(d["nsnps"]) % spec["drop_every"] != 0).astype(np.uint8)
    return d, alphas, opt("--doublet-prior", 0.5), np.full(d["nsnps"], err), has_gp


def best_rows(fx):
    """Reference rows keyed by barcode; values converted per BEST_COLS."""
    rows = {}
    for r in fx["best_raw"]:
        if len(r) != len(BEST_COLS):
            continue  # header line (no conversions)
        row = dict(zip(BEST_COLS, r))
        for k in ("bestA", "bestLLK", "nextA", "nextLLK", "diffBestNext", "bestPP", "sngPP", "sngBestLLK", "sngNextLLK",
                  "sngOnlyPP", "dBestAlpha", "dblBestLLK", "diffSngDbl"):
            row[k] = float(row[k])
        for k in ("int_id", "nsnps", "nreads"):
            row[k] = int(row[k])
        rows[row["barcode"]] = row
    return rows


def fmx_rows(fx):
    rows = []
    for r in fx["samples_raw"]:
        if len(r) != len(FMX_COLS):
            continue
        row = dict(zip(FMX_COLS, r))
        for k in ("bestLLK", "nextLLK", "diff", "bestPP", "sngPP", "sngBestLLK", "sngNextLLK", "sngOnlyPP", "dblBestLLK",
                  "diffSngDbl"):
            row[k] = float(row[k])
        for k in ("int_id", "nsnps", "nreads", "jBest", "kBest", "jNext", "kNext", "sBest", "sNext", "dBest1", "dBest2"):
            row[k] = int(row[k])
        rows.append(row)
    return rows


def fmx_init_inputs(fx):
    """Inputs of the initial clustering as cmdCramFreemuxlet sees them: dataset, pileups, AF, singlet scores."""
    from oracle import pyoracle
    d = dataset(fx)
    plp = pyoracle.fmx_pileup(d["read_ptr"], d["al"], d["bq"], 0.5)
    af = np.asarray(["%.5f" % a for a in d["af"]], np.float64)  # AF as it round-trips through .var.gz
    args = fx["spec"]["args"]

    def opt(flag, default):
        return float(args[args.index(flag) + 1]) if flag in args else default

    v2 = fx["case"].startswith("fmx2")
    params = dict(bf_thres=opt("--bf-thres", 5.41), frac_init_clust=opt("--frac-init-clust", 1.0),
                  geno_error=opt("--geno-error", 0.1 if v2 else 0.0), keep_init_missing="--keep-init-missing" in args,
                  doublet_prior=opt("--doublet-prior", 0.5))
    return d, plp, af, params


def sparse_pairs(dense):
    """orc_dropd matrix [nbcs][nbcs] -> the engine's record list (id1 > id2, sorted)."""
    from cramore_b200.engine import PAIR_DIST_DTYPE
    i, j = np.nonzero(dense["nsnps"])
    out = np.zeros(len(i), PAIR_DIST_DTYPE)
    out["id1"], out["id2"] = i, j
    for f in ("nsnps", "nread1", "nread2", "llk0", "llk2"):
        out[f] = dense[f][i, j]
    return out
