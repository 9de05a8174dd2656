"""Multi-GPU plumbing: one process per GPU over torch.distributed (NCCL on GPUs,
gloo in the CPU tests).

The classification path shards embarrassingly: reads are independent units
(`Classifier::Query` is pure w.r.t. the index), so every rank
opens its own replica of the index on its GPU and classifies a contiguous shard
of the reads with no data-path collective.  The only exchanges are
  * the final SUM all-reduce of the per-taxon assignment counters, and
  * (optional) gathering the per-shard TSV text on rank 0 in input order.
"""
import numpy as np


def shard_bounds(n_items, rank, world):
    """Contiguous, balanced [lo, hi) of `n_items` for `rank` (first n % world ranks get one extra)."""
    if world <= 0 or not (0 <= rank < world):
        raise ValueError("bad rank/world")
    base, extra = divmod(int(n_items), int(world))
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def shard_packed(seq, off, rank, world):
    """Slice one packed read buffer (bytes, offsets[n+1]) to this rank's shard (offsets re-based)."""
    n = len(off) - 1
    lo, hi = shard_bounds(n, rank, world)
    o = np.asarray(off[lo:hi + 1], dtype=np.uint64)
    s = np.asarray(seq[int(o[0]):int(o[-1])])
    return s, (o - o[0]).astype(np.uint64), lo, hi


class _DevVec:
    """Zero-copy torch view of a uint64 device vector owned by libcfrb200.so (as int64:
    SUM over non-negative counters is bit-identical)."""

    def __init__(self, ptr, n):
        self.__cuda_array_interface__ = {"shape": (int(n),), "typestr": "<i8", "data": (int(ptr), False),
                                         "version": 2}


def device_counts_tensor(clf):
    """torch int64 CUDA tensor aliasing the classifier's per-taxon counters in HBM."""
    import torch
    ptr, n = clf.taxon_counts_device()
    return torch.as_tensor(_DevVec(ptr, n), device="cuda")


def final_counts(clf, group=None):
    """The path's one collective: SUM all-reduce of the per-taxon assignment counters, once, after the
    last batch.  The library's live counter vector is snapshotted first (after a device
    synchronize, so no scoring kernel is still adding to it) and the COPY is reduced: the live
    vector stays this rank's own cumulative count and can be reduced again later without
    double counting."""
    import torch
    torch.cuda.synchronize()
    snap = device_counts_tensor(clf).clone()
    return allreduce_counts(snap, group=group)


def allreduce_counts(t, group=None):
    """SUM all-reduce of a counter tensor (NCCL for CUDA tensors, gloo for CPU tensors)."""
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    return t


def gather_in_order(obj, dst=0, group=None):
    """Gather one picklable object per rank on `dst`, ordered by rank (None elsewhere)."""
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return [obj]
    world = dist.get_world_size(group)
    out = [None] * world if dist.get_rank(group) == dst else None
    dist.gather_object(obj, out, dst=dst, group=group)
    return out
