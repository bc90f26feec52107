"""world_size-2 `gloo` test of the N>1 path's host logic (CPU): replica sharding + the exact limb all-reduce.

The device side of a rank (its replicas' accumulators) is stood in for by exact integers computed from the oracle;
what is under test is that two ranks with disjoint replica blocks reduce to the same totals as one rank with all
replicas — bit for bit, independent of the world size."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as tmp

import _libs

ROOT = _libs.ROOT
N_REPLICAS, L = 6, 16


def replica_values(replica):
    """A deterministic stand-in for one replica's accumulator row: S sums and products of a hot-start pyramid."""
    o = _libs.oracle()
    s = np.zeros((L, L), np.int32)
    o.orc_hot_start(L, 99, replica, s)
    S = _libs.pyramid(L, s, 99, replica, 0)
    vals = [int(x) for x in S.ravel()]
    vals += [int(S[1, 0]) * int(S[0, 0]) * (1 << 40), -int(S[0, 1]) * (1 << 70) - replica]  # exercise the high limbs
    return vals


def _worker(rank, world, port, out_dir):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from mcrg_b200 import dist as mdist

    first, count = mdist.shard_replicas(N_REPLICAS, world, rank)
    local = [replica_values(r) for r in range(first, first + count)]
    totals = [sum(col) for col in zip(*local)]
    limbs = torch.from_numpy(mdist.ints_to_limbs(totals))
    mdist.allreduce_limbs(limbs)
    np.save(os.path.join(out_dir, f"rank{rank}.npy"), limbs.numpy())
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_two_rank_allreduce_is_exact_and_world_size_independent(tmp_path, world):
    from mcrg_b200 import dist as mdist

    port = 29500 + (os.getpid() % 2000) + world
    tmp.spawn(_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    want = [sum(col) for col in zip(*[replica_values(r) for r in range(N_REPLICAS)])]
    for rank in range(world):
        got = mdist.limbs_to_ints(np.load(tmp_path / f"rank{rank}.npy"))
        assert got == want
