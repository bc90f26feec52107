"""Multi-GPU plumbing: replica sharding and the one collective of the path.

The reference's only parallelism is "one independent Markov chain per MPI rank, then MPI_Allreduce(SUM) of the
accumulators" (mcrg.cpp:101-103).  Here: one process per GPU, the global replica list is cut into contiguous
blocks (the global replica id enters every Philox counter, so the union of all ranks' chains is the same set of
chains whatever the world size), and ONE int64 all-reduce combines the accumulator totals.  The totals travel as
32-bit limbs in int64 containers, so the sum is exact and order independent: 1, 2, 4 and 8 GPUs give
bit-identical totals.
"""
import numpy as np


def shard_replicas(n_total, world_size, rank):
    """-> (first_global_replica, count) of this rank: contiguous blocks, remainder to the low ranks."""
    if world_size < 1 or not (0 <= rank < world_size):
        raise ValueError("bad rank/world_size")
    q, r = divmod(n_total, world_size)
    count = q + (1 if rank < r else 0)
    first = rank * q + min(rank, r)
    return first, count


def ints_to_limbs(values):
    """exact ints (|v| < 2^127) -> int64 array [n, 4] of 32-bit limbs, top limb signed (what the device emits)."""
    out = np.zeros((len(values), 4), np.int64)
    for i, v in enumerate(values):
        v = int(v)
        lo = v & ((1 << 64) - 1)
        hi = v >> 64  # arithmetic
        out[i, 0] = lo & 0xFFFFFFFF
        out[i, 1] = lo >> 32
        out[i, 2] = hi & 0xFFFFFFFF
        out[i, 3] = hi >> 32
    return out


def limbs_to_ints(limbs):
    """int64 array [n, 4] (possibly summed over ranks, limbs then exceed 32 bits) -> list of exact ints."""
    limbs = np.asarray(limbs, np.int64).reshape(-1, 4)
    return [int(l[0]) + (int(l[1]) << 32) + (int(l[2]) << 64) + (int(l[3]) << 96) for l in limbs]


def allreduce_limbs(limbs_tensor, group=None):
    """In-place SUM all-reduce of a torch int64 tensor of limbs (NCCL on GPUs, gloo in the CPU tests)."""
    import torch.distributed as dist

    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(limbs_tensor, op=dist.ReduceOp.SUM, group=group)
    return limbs_tensor
