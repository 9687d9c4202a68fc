"""CPU, world_size 2 over gloo: the N>1 host logic of bench.py -- cost-balanced contiguous pair shards, every pair
processed exactly once, results re-assembled in pair order, max-over-ranks reduction of the timed region."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from equidock_public_b200 import sharding


def test_shard_bounds_cover_and_balance():
    rng = np.random.default_rng(0)
    sizes = [(int(a), int(b)) for a, b in rng.integers(40, 1200, size=(57, 2))]
    costs = [sharding.pair_cost(a, b, 10 * a, 10 * b) for a, b in sizes]
    for world in (1, 2, 4, 8):
        b = sharding.shard_bounds(costs, world)
        assert b[0][0] == 0 and b[-1][1] == len(costs)
        assert all(b[i][1] == b[i + 1][0] for i in range(world - 1))
        loads = [sum(costs[lo:hi]) for lo, hi in b]
        assert max(loads) <= sum(costs) / world + max(costs)        # within one pair of the ideal
    few = sharding.shard_bounds([1.0, 1.0], 4)                       # fewer pairs than ranks: still a partition
    assert few[0][0] == 0 and few[-1][1] == 2 and sum(hi - lo for lo, hi in few) == 2


def _worker(rank, world, port, out_dir):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    rng = np.random.default_rng(3)
    sizes = rng.integers(20, 400, size=(23, 2))
    costs = [sharding.pair_cost(int(a), int(b), 10 * int(a), 10 * int(b)) for a, b in sizes]
    lo, hi = sharding.my_shard(costs, world, rank)
    # stand-in for the per-pair engine output: a deterministic function of the pair alone
    local = torch.tensor([[i, float(sizes[i, 0] * 3 + sizes[i, 1])] for i in range(lo, hi)], dtype=torch.float64).reshape(-1, 2)
    counts = [torch.zeros(1, dtype=torch.int64) for _ in range(world)]
    dist.all_gather(counts, torch.tensor([local.shape[0]], dtype=torch.int64))
    padded = torch.zeros(23, 2, dtype=torch.float64)
    padded[:local.shape[0]] = local
    gathered = [torch.zeros(23, 2, dtype=torch.float64) for _ in range(world)]
    dist.all_gather(gathered, padded)
    full = torch.cat([g[:int(c)] for g, c in zip(gathered, counts)])
    t = torch.tensor([1.0 + rank], dtype=torch.float64)              # "elapsed" differs per rank
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0:
        torch.save({'full': full, 'tmax': t, 'sizes': torch.from_numpy(sizes)}, os.path.join(out_dir, 'r0.pt'))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(120)
def test_two_rank_gloo_sharded_run_equals_single(tmp_path):
    with socket.socket() as s:
        s.bind(('127.0.0.1', 0))
        port = s.getsockname()[1]
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    res = torch.load(os.path.join(str(tmp_path), 'r0.pt'))
    sizes = res['sizes'].numpy()
    assert res['full'][:, 0].tolist() == list(range(23))            # every pair exactly once, in order
    assert np.allclose(res['full'][:, 1].numpy(), sizes[:, 0] * 3 + sizes[:, 1])
    assert float(res['tmax']) == 2.0                                  # max over ranks
