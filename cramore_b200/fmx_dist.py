"""Multi-GPU freemuxlet EM driver: one process per GPU.

  E-step  : barcode shards (droplets are independent given the cluster statistics)
  M-step  : SNP slabs.  snp_droplet_pileup::merge (sc_drop_seq.h:77-101) clamps after every droplet, so
            a cluster statistic is an ORDERED fold over droplets and cannot be formed by summing
            per-shard partials.  Every rank therefore folds all droplets, in ascending id, for its own
            SNP slab and leaves zeros elsewhere; a sum-allreduce (NCCL over NVLink, in place on the
            engine's array) then assembles the full array bit-exactly on every rank.
  exchange: the E-step reads the statistics only through the per-(SNP, cluster) genotype posteriors
            (cmd_cram_freemuxlet.cpp:476-489; 3 doubles where the statistics have 12), so each rank derives
            the posteriors of its slab and THAT array is all-reduced every iteration; the statistics array
            itself is all-reduced once, after the last M-step (it is only needed for the cluster VCF).
  per iteration: posteriors of the slab -> all-reduce -> E-step -> all-gather of the assignment vector -> M-step.
  exchange="stats" keeps the heavier form (all-reduce of the 12-double statistics every iteration).
  exchange="push" (the B200 form, and the single-GPU loop): the fused kernels -- the M-step kernel forms the
            posteriors of its slab and stores them into every rank's array over NVLink, the decision kernel does the
            same with the assignments, and the ranks are ordered by epoch flags inside those kernels
            (include/cramore_cuda.h); two launches + the decision per iteration, no collective on the payload.

The engine object only has to provide the methods used below, which lets the CPU test suite drive the
same loop over gloo with a numpy stand-in."""
import numpy as np

from .shard import balanced_ranges, gather_results


class _DevArray:
    """CUDA array interface view of engine-owned device memory (no copy)."""

    def __init__(self, ptr, n, typestr):
        self.__cuda_array_interface__ = {"shape": (int(n),), "typestr": typestr, "data": (int(ptr), False),
                                         "version": 2}


def device_views(eng, device):
    """torch tensors aliasing the engine's statistics (float64) and assignment (int32) arrays."""
    import torch
    sp, sn = eng.stats_dev()
    ap, an = eng.assign_dev()
    stats = torch.as_tensor(_DevArray(sp, sn, "<f8"), device=device)
    assign = torch.as_tensor(_DevArray(ap, an, "<i4"), device=device)
    assert stats.data_ptr() == sp and assign.data_ptr() == ap
    return stats, assign


def posterior_view(eng, device):
    """torch tensor aliasing the engine's per-(SNP, cluster) posterior array [nsnps][K][3] (float64)."""
    import torch
    pp, pn = eng.post_dev()
    post = torch.as_tensor(_DevArray(pp, pn, "<f8"), device=device)
    assert post.data_ptr() == pp
    return post


def snp_slabs(snp_idx, nsnps, world):
    """Contiguous SNP ranges with ~equal numbers of (cell,SNP) entries."""
    counts = np.bincount(np.asarray(snp_idx, np.int64), minlength=nsnps)
    ptr = np.concatenate([[0], np.cumsum(counts)])
    return balanced_ranges(ptr, world)


def run_em_distributed(eng, stats, assign, cell_ranges, snp_ranges, init_assign, niter=10,
                       geno_error_every_iter=False, group=None, sync=None, post=None, exchange=None, graphs=None,
                       push_barrier="kernel", timing=None):
    """cmd_cram_freemuxlet.cpp:359-370 + :457-653 across ranks.  `stats`/`assign` (and `post`) are torch
    tensors aliasing the engine's arrays (device_views / posterior_view, or CPU tensors of a stand-in
    engine).  exchange: "posteriors" (default when `post` is given) or "stats".  Returns this rank's
    per-droplet results and every rank's results gathered in barcode order.

    Stream order: with device tensors the engine is put on torch's CURRENT stream here (`eng.set_stream`), and stays
    there: its kernels, the copies into `assign` / `post` and the NCCL collectives below are then ordered by that one
    stream, which is what the loop relies on (the engine's own default is a separate non-blocking stream, on which a
    collective could read half-written posteriors).  `sync` is only needed by stand-in engines that do not run on the
    stream (CPU tests)."""
    import torch
    import torch.distributed as dist
    if stats.device.type == "cuda":
        eng.set_stream(torch.cuda.current_stream(stats.device).cuda_stream)
    multi = dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1
    rank = dist.get_rank(group) if multi else 0
    cb, ce = cell_ranges[rank]
    sb, se = snp_ranges[rank]
    eng.set_shard(cb, ce, sb, se)
    sync = sync or (lambda: None)
    if exchange is None:
        exchange = "posteriors" if post is not None else "stats"
    if exchange not in ("posteriors", "stats", "push") or (exchange == "posteriors" and post is None):
        raise ValueError("exchange must be 'stats', 'posteriors' (needs the posterior tensor) or 'push'")
    if exchange == "push" and multi:
        # the mappings persist on the engine (they are tied to its arrays, not to one EM run)
        key = (dist.get_world_size(group), rank)
        if getattr(eng, "_peer_world", None) != key:
            mine = eng.peer_handles()
            everyone = [None] * key[0]
            dist.all_gather_object(everyone, mine, group=group)
            eng.peer_open(rank, everyone)
            eng._peer_world = key
        # the 4-byte token of the rank barrier; a captured graph keeps pointing at it, so it lives with the graphs
        token = torch.zeros(1, dtype=torch.int32, device=assign.device)
        if graphs is not None:
            token = graphs.setdefault("_barrier_token", token)

    def rank_barrier():
        # push_barrier == "nccl": a 4-byte all-reduce on the shared stream orders the ranks between the fused kernels
        # (instead of the epoch flags inside them); completes only after every rank's stream has reached it
        sync()
        dist.all_reduce(token, op=dist.ReduceOp.SUM, group=group)

    def exchange_assign():
        # each rank owns assign[cb:ce]; a sum of (value + 1) over zero-filled vectors is an all-gather
        if not multi:
            return
        sync()
        tmp = torch.zeros_like(assign)
        tmp[cb:ce] = assign[cb:ce] + 1
        dist.all_reduce(tmp, op=dist.ReduceOp.SUM, group=group)
        assign.copy_(tmp - 1)

    def exchange_stats():
        if not multi:
            return
        sync()
        dist.all_reduce(stats, op=dist.ReduceOp.SUM, group=group)

    def exchange_post():
        if not multi:
            return
        sync()
        dist.all_reduce(post, op=dist.ReduceOp.SUM, group=group)

    def apply_at(it):  # geno_error for the E-step of iteration `it` (:485 vs cmd_cram_freemux2.cpp:410)
        return bool(geno_error_every_iter) or (it + 1 == niter)

    def em_loop_fused():
        in_kernel = (not multi) or push_barrier == "kernel"

        def order():
            if multi and push_barrier == "nccl":
                rank_barrier()
        eng.mstep_fused_dev(apply_at(0), False, in_kernel)
        order()
        for it in range(niter):
            eng.estep_fused_dev(in_kernel)
            order()
            eng.mstep_fused_dev(apply_at(it + 1), it + 1 == niter, in_kernel)
            order()

    def em_loop():
        if exchange == "push":
            return em_loop_fused()
        eng.mstep_dev()
        if exchange == "stats":
            exchange_stats()
        for it in range(niter):
            apply = apply_at(it)
            if exchange == "stats":
                eng.estep_dev(apply)
            else:
                eng.posteriors_dev(apply)
                exchange_post()
                eng.estep_post_dev()
            exchange_assign()
            eng.mstep_dev()
            if exchange == "stats":
                exchange_stats()

    def finish():
        # once, after the loop: every rank folded only its own SNP slab, the full statistics are needed for the output
        if exchange != "stats":
            exchange_stats()

    assign.copy_(torch.as_tensor(np.ascontiguousarray(init_assign, np.int32)).to(assign.device))
    key = (exchange, niter, bool(geno_error_every_iter), cb, ce, sb, se)
    # graphs: the loop itself (kernels only in the "push" form with in-kernel ordering -- no collective inside the
    # captured region) replayed as one CUDA graph; the final all-reduce of the statistics stays outside
    ev = None
    if timing is not None and stats.device.type == "cuda":
        ev = (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
        ev[0].record()
    if graphs is None or stats.device.type == "cpu":
        em_loop()
    elif key not in graphs:
        graphs[key] = None  # warm-up: eager
        em_loop()
    else:
        if graphs[key] is None:
            g = torch.cuda.CUDAGraph()
            outer = torch.cuda.current_stream()
            with torch.cuda.graph(g):
                eng.set_stream(torch.cuda.current_stream().cuda_stream)  # the capture stream
                em_loop()
            eng.set_stream(outer.cuda_stream)
            graphs[key] = g
        graphs[key].replay()
    if ev is not None:
        ev[1].record()
        timing["loop_events"] = ev
    finish()
    sync()
    local = eng.download_results()
    if exchange == "push" and multi and push_barrier in ("kernel", "none") and eng.rank_barrier_failed():
        raise RuntimeError(eng.last_error())
    if not multi:
        return local, local
    device = None if stats.device.type == "cpu" else stats.device
    return local, gather_results(local, cell_ranges, device=device, group=group)
