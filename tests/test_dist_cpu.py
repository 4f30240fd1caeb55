"""World-size-2 gloo test of the host logic of the multi-GPU (replica) run."""
import os

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from deformablelka_b200 import dist as dd
    assert dd.env_world() == (world, rank, rank)
    lo, hi = dd.shard_batch(16, world, rank)
    assert (lo, hi) == (rank * 8, rank * 8 + 8)
    # rank 1 is slower: the job's time is the max over ranks, the value counts both ranks' voxels
    val, ms = dd.aggregate_throughput(1000, 10, 10.0 * (rank + 1))
    out[rank] = (val, ms)
    dist.barrier()
    dist.destroy_process_group()


def test_replica_aggregation_gloo_world2():
    mgr = mp.Manager()
    out = mgr.dict()
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(2, port, out), nprocs=2, join=True)
    for r in (0, 1):
        val, ms = out[r]
        assert ms == pytest.approx(20.0)
        assert val == pytest.approx(2 * 1000 * 10 / 20e-3)


def test_shard_batch_rejects_uneven_split():
    from deformablelka_b200 import dist as dd
    with pytest.raises(ValueError):
        dd.shard_batch(10, 4, 0)
    assert dd.max_over_ranks(3.5) == 3.5
