"""Barcode sharding across GPUs: contiguous barcode ranges balanced by the number
of (cell,SNP) entries -- the engine's cost driver -- one shard per rank, genotypes replicated.
This is the multi-process analogue of the reference's --group-list recipe
(cmd_cram_demuxlet.cpp:75,118-131).  Demuxlet needs no data-path collective: the only exchange is
the final gather of the per-droplet result records."""
import numpy as np


def balanced_ranges(cell_ptr, world):
    """[(begin, end)] * world: contiguous barcode ranges with ~equal entry counts."""
    cell_ptr = np.asarray(cell_ptr, np.int64)
    nbcs = len(cell_ptr) - 1
    total = int(cell_ptr[-1])
    cuts = [0]
    for r in range(1, world):
        target = total * r / world
        c = int(np.searchsorted(cell_ptr, target, side="left"))
        c = min(max(c, cuts[-1]), nbcs)
        cuts.append(c)
    cuts.append(nbcs)
    return [(cuts[i], cuts[i + 1]) for i in range(world)]


def slice_csr(cell_ptr, snp_idx, read_ptr, al, bq, begin, end):
    """CSR of barcodes [begin, end) re-based to zero."""
    cell_ptr = np.asarray(cell_ptr, np.int64)
    read_ptr = np.asarray(read_ptr, np.int64)
    e0, e1 = int(cell_ptr[begin]), int(cell_ptr[end])
    r0, r1 = int(read_ptr[e0]), int(read_ptr[e1])
    return (cell_ptr[begin:end + 1] - e0, np.asarray(snp_idx)[e0:e1], read_ptr[e0:e1 + 1] - r0,
            np.asarray(al)[r0:r1], np.asarray(bq)[r0:r1])


def gather_results(local, ranges, device=None, group=None):
    """All-gather of the per-rank result records (any structured dtype) into barcode order.
    `ranges` is balanced_ranges(...) of the full store; works over gloo (CPU tensors) and nccl
    (pass device="cuda")."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    assert len(ranges) == world and len(local) == ranges[rank][1] - ranges[rank][0]
    item = local.dtype.itemsize
    cap = max(e - b for b, e in ranges) * item
    buf = np.zeros(max(cap, 1), np.uint8)
    buf[:len(local) * item] = np.ascontiguousarray(local).view(np.uint8).reshape(-1)
    t = torch.from_numpy(buf)
    if device is not None and dist.get_backend(group) != "gloo":  # gloo gathers host tensors only
        t = t.to(device)
    outs = [torch.empty_like(t) for _ in range(world)]
    dist.all_gather(outs, t, group=group)
    parts = []
    for (b, e), o in zip(ranges, outs):
        parts.append(o.cpu().numpy()[:(e - b) * item].view(local.dtype))
    return np.concatenate(parts) if parts else local[:0]


def score_sharded(score_fn, cell_ptr, snp_idx, read_ptr, al, bq, device=None, group=None):
    """Each rank scores its own barcode shard with score_fn(cell_ptr, snp_idx, read_ptr, al, bq)
    (normally DemuxEngine.score) and every rank receives the full result array."""
    import torch.distributed as dist
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    ranges = balanced_ranges(cell_ptr, world)
    b, e = ranges[rank]
    local = score_fn(*slice_csr(cell_ptr, snp_idx, read_ptr, al, bq, b, e))
    return gather_results(local, ranges, device=device, group=group)
