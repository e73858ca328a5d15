#!/usr/bin/env python
"""Freemuxlet EM benchmark (BASELINE configs[3]: K=16 clusters, 50k barcodes, 300k SNPs) on N GPUs.

    python tools/bench_fmx.py [--nbcs N --nsnps M --iters I --check]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port P tools/bench_fmx.py ...

Every rank holds the full droplet store and pileups (1 GB at C4); the E-step is barcode-sharded, the
M-step SNP-sharded, and the per-cluster per-SNP statistics are sum-all-reduced over NCCL each
iteration (cramore_b200/fmx_dist.py).  --check also runs the single-GPU loop on rank 0 and requires
bit-identical results.  Prints one JSON line from rank 0."""
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
    ap.add_argument("--iters", type=int, default=10)
    ap.add_argument("--check", action="store_true")
    ap.add_argument("--exchange", default="push", choices=["posteriors", "stats", "push"])
    ap.add_argument("--push-barrier", default="kernel", choices=["kernel", "nccl", "none"],
                    help="how the ranks are ordered in the push exchange; none = no ordering at all (WRONG results, timing probe)")
    ap.add_argument("--graph", action="store_true", help="run the whole EM loop as one CUDA graph (captured on the second run)")
    args = ap.parse_args()
    import torch
    import torch.distributed as dist
    from cramore_b200 import FreemuxEngine
    from cramore_b200.fmx_dist import device_views, posterior_view, run_em_distributed, snp_slabs
    from cramore_b200.shard import balanced_ranges
    from cramore_b200.synth import make_dataset

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    # the engine and torch (NCCL) must share ONE real stream: handle 0 (the legacy default stream) means
    # "engine-owned stream" to ccu_set_stream
    torch.cuda.set_stream(torch.cuda.Stream(device=dev))

    K, GE = 16, 0.0
    d = make_dataset("C4", nbcs=args.nbcs, nsnps=args.nsnps)
    rng = np.random.default_rng(11)
    init = (d["s1"] % K).astype(np.int32)  # stands for --init-cluster (B.5: keeps the timed region = the EM loop)
    init[d["is_dbl"]] = -1
    init[rng.random(len(init)) < 0.3] = -1

    eng = FreemuxEngine(local)
    eng.set_stream(torch.cuda.current_stream().cuda_stream)
    eng.configure_fmx(K, 0.5, GE)
    t0 = time.perf_counter()
    eng.upload_fmx(d["cell_ptr"], d["snp_idx"], d["read_ptr"], d["al"], d["bq"], d["af"])
    t_upload = time.perf_counter() - t0
    setup = eng.fmx_timings()
    stats, assign = device_views(eng, dev)
    post = posterior_view(eng, dev)
    cell_ranges = balanced_ranges(d["cell_ptr"], world)
    slabs = snp_slabs(d["snp_idx"], d["nsnps"], world)

    graphs = {} if args.graph else None

    def run():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        tm = {}
        loc, full = run_em_distributed(eng, stats, assign, cell_ranges, slabs, init, niter=args.iters, post=post,
                                       exchange=args.exchange, graphs=graphs, push_barrier=args.push_barrier, timing=tm)
        e1.record()
        torch.cuda.synchronize()
        run.loop_ms = tm["loop_events"][0].elapsed_time(tm["loop_events"][1])
        return loc, full, e0.elapsed_time(e1)

    def breakdown():
        """One extra iteration with a device synchronise around each phase (not part of the timed loop)."""
        out = {}
        def timed(name, fn):
            torch.cuda.synchronize()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            fn()
            b.record()
            torch.cuda.synchronize()
            out[name] = a.elapsed_time(b)
        timed("estep_ms", lambda: eng.estep_fused_dev(False))
        timed("mstep_ms", lambda: eng.mstep_fused_dev(False, False, False))
        timed("estep_unfused_ms", lambda: eng.estep_dev(False))
        timed("mstep_unfused_ms", lambda: eng.mstep_dev())
        if world > 1:
            timed("allreduce_stats_ms", lambda: dist.all_reduce(stats))
            timed("allreduce_posteriors_ms", lambda: dist.all_reduce(post))
            if args.exchange == "push":
                tok = torch.zeros(1, dtype=torch.int32, device=dev)
                timed("rank_barrier_nccl_ms", lambda: dist.all_reduce(tok))
                timed("rank_barrier_kernel_ms", lambda: eng.rank_barrier_dev())
        return out

    run()  # warm-up (NCCL channels, allocations)
    if args.graph:
        run()  # captures the loop, then replays it once
    loc, full, ms = run()
    loop_ms = run.loop_ms
    phases = breakdown()
    t = torch.tensor([ms, loop_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms, loop_ms = float(t[0].item()), float(t[1].item())
    ok = None
    if args.check and rank == 0:
        ref_eng = FreemuxEngine(local)
        ref_eng.configure_fmx(K, 0.5, GE)
        ref_eng.upload_fmx(d["cell_ptr"], d["snp_idx"], d["read_ptr"], d["al"], d["bq"], d["af"])
        ref, _ = ref_eng.run_em(init, niter=args.iters)
        ok = bool(np.array_equal(ref, full))
        ref_eng.close()
    if rank == 0:
        nnz = int(d["cell_ptr"][-1])
        # HBM roofline of the two per-iteration kernels, algorithmic bytes of SURVEY 8(d):
        #   E-step: 84 B pileup + 24*K B gathered cluster posteriors per entry
        #   M-step: 84 B pileup per entry in, 84 B per (cluster, observed SNP) out
        try:
            hbm_peak, peak_src = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"], "MEASURED_PEAKS.json"
        except Exception:
            hbm_peak, peak_src = 6650.0, "fallback of B200_PROFILING.md"
        nobs = int(np.unique(d["snp_idx"]).size)
        e_bytes = (20 + 8 * K) * nnz / world  # single-read entries: (gl[0]/2, b) + SNP id + K mean genotypes
        m_bytes = (92 * nnz + 32 * K * d["nsnps"]) / world  # pileups + droplet ids in, posteriors + mean genotypes out
        roof = []
        for name, key, nbytes in (("fmx_estep_rows_kernel (F2)", "estep_ms", e_bytes),
                                  ("fmx_mstep_kernel (F3)", "mstep_ms", m_bytes)):
            gbs = nbytes / (phases[key] * 1e-3) / 1e9
            roof.append({"kernel": name, "bound": "hbm", "achieved": gbs, "peak": hbm_peak, "unit": "GB/s",
                         "frac": gbs / hbm_peak, "algorithmic_bytes_per_launch": int(nbytes), "ms": phases[key],
                         "peak_source": peak_src})
        line = {"metric": "freemuxlet EM iterations/s", "value": args.iters / (ms * 1e-3), "unit": "iter/s",
                "n_gpus": world, "iters": args.iters, "ms_per_iter": ms / args.iters,
                "ms_per_iter_loop_only": loop_ms / (args.iters + 0.3),
                "config": {"workload": "C4 freemuxlet: K=%d, %d barcodes, %d SNPs, %d entries" %
                                       (K, d["nbcs"], d["nsnps"], nnz)},
                "exchange": args.exchange, "cuda_graph": bool(args.graph), "push_barrier": args.push_barrier, "stats_allreduce_bytes": int(stats.numel() * 8),
                "posterior_allreduce_bytes": int(post.numel() * 8), "setup": setup, "upload_s": t_upload, "phase_ms": phases, "roofline": roof,
                "singlets": int((full["type"] == 0).sum()), "doublets": int((full["type"] == 1).sum()),
                "matches_single_gpu": ok}
        print(json.dumps(line), flush=True)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    if graphs:
        # a captured graph holds NCCL work of the process group: tearing the group down under it has been seen to
        # hang, so the graphs go first and the process leaves without the group's destructor
        graphs.clear()
        torch.cuda.synchronize()
        eng.close()
        sys.stdout.flush()
        sys.stderr.flush()
        os._exit(0)
    eng.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
