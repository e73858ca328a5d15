#!/usr/bin/env python
"""Freemuxlet initial clustering at BASELINE configs[3] size (K=16, 50k barcodes, 300k SNPs): pairwise droplet
distances on the GPU (cmd_cram_freemuxlet.cpp:172-221) and the host votes (:223-346).  Prints one JSON line.

    python tools/bench_fmx_init.py [--nbcs N --nsnps M --iter-init I]"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--nbcs", type=int, default=None)
    ap.add_argument("--nsnps", type=int, default=None)
    ap.add_argument("--iter-init", type=int, default=10)
    args = ap.parse_args()
    from cramore_b200 import FreemuxEngine
    from cramore_b200.engine import initial_clusters
    from cramore_b200.synth import make_dataset

    K = 16
    d = make_dataset("C4", nbcs=args.nbcs, nsnps=args.nsnps)
    eng = FreemuxEngine(0)
    eng.configure_fmx(K, 0.5, 0.0)
    eng.upload_fmx(d["cell_ptr"], d["snp_idx"], d["read_ptr"], d["al"], d["bq"], d["af"])
    l0, l2 = eng.lmix()
    t0 = time.perf_counter()
    eng.pair_distances(download=False)
    pairs = eng.voting_pairs(5.41)  # the pairs that can cast a vote (|llk2 - llk0| > --bf-thres), compacted on the device
    t_pairs = time.perf_counter() - t0
    st = eng.pair_stats
    t0 = time.perf_counter()
    clusts = initial_clusters(l2 - l0, pairs, K, iter_init=args.iter_init)
    t_votes = time.perf_counter() - t0
    voting = len(pairs)
    # bytes the device path moves at least once: pileup diagonals + counts in, 32-byte records out and back in
    nrec = st["records"]
    line = {"metric": "freemuxlet initial clustering", "config": {"workload": "C4: K=%d, %d barcodes, %d SNPs, %d entries" %
                                                                  (K, d["nbcs"], d["nsnps"], int(d["cell_ptr"][-1]))},
            "pairs": st["pairs"], "pair_snp_records": nrec, "gpu_ms": st["ms"],
            "gpu_records_per_s": nrec / (st["ms"] * 1e-3) if st["ms"] > 0 else None,
            "pair_distances_s_incl_download": t_pairs, "host_votes_s": t_votes, "iter_init": args.iter_init,
            "dense_matrix_bytes_of_the_reference": int(d["nbcs"]) * int(d["nbcs"]) // 2 * 32,
            "sparse_bytes_on_device": int(st["pairs"]) * 40, "pairs_passing_bf_thres": voting, "bytes_downloaded": int(pairs.nbytes),
            "clusters_used": int(len(np.unique(clusts[clusts >= 0])))}
    print(json.dumps(line), flush=True)
    eng.close()


if __name__ == "__main__":
    main()
