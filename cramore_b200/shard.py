"""Barcode sharding across GPUs: contiguous barcode ranges balanced by the number
of (cell,SNP) entries -- the engine's cost driver -- one shard per rank, genotypes replicated.
This is the multi-process analogue of the reference's --group-list recipe
(cmd_cram_demuxlet.cpp:75,118-131)."""
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
