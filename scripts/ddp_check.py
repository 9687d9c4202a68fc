#!/usr/bin/env python
"""Multi-GPU check of the data-parallel trainer (run with torchrun on N GPUs):
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 scripts/ddp_check.py
Every rank trains ONE step on its shard of a 2N-pair ragged batch (bucketed NCCL all-reduce overlapped with backward, fused
clip + Adam); rank 0 then repeats the step single-process on the whole batch from the same initial weights and compares the
updated parameters: data-parallel == global batch."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, 'tests'), os.path.join(ROOT, 'oracle')):
    sys.path.insert(0, p)
import golden_io as gio  # noqa: E402
from equidock_public_b200 import hetero_graph as hg, synthetic  # noqa: E402
from equidock_public_b200.losses import PocketBatch  # noqa: E402
from equidock_public_b200.training import DataParallelTrainer  # noqa: E402


def main():
    rank, world, lr = int(os.environ['RANK']), int(os.environ['WORLD_SIZE']), int(os.environ['LOCAL_RANK'])
    torch.cuda.set_device(lr)
    dev = torch.device('cuda', lr)
    dist.init_process_group('nccl', device_id=dev)
    per = 2
    rng = np.random.default_rng(5)
    sizes = [(int(a), int(b)) for a, b in rng.integers(40, 180, size=(per * world, 2))]
    pairs = [synthetic.synthetic_pair(np.random.default_rng([3, i]), a, b, 10) for i, (a, b) in enumerate(sizes)]

    def batch(idx):
        g = hg.batch_pairs(synthetic.to_torch_pairs([pairs[i] for i in idx])).to(dev)
        bl = [torch.from_numpy(pairs[i][0]['x']) for i in idx]
        br = [torch.from_numpy(pairs[i][1]['x'] + 8.0) for i in idx]
        pk = [torch.from_numpy((0.5 * (pairs[i][0]['x'][:11] + pairs[i][1]['x'][:11] + 8.0)).astype(np.float32)) for i in idx]
        return g, PocketBatch(bl, br, pk, pk, dev)

    model = gio.build_model('db5', dev)
    tr = DataParallelTrainer(model, lr=1e-3, weight_decay=1e-4, clip=100.0, world=world)
    g, t = batch(range(rank * per, (rank + 1) * per))
    out = tr.step(g, t)
    torch.cuda.synchronize()
    w_dp = tr.flat_w.clone()
    g_dp = tr.flat_g.clone()          # the all-reduced, averaged (and clipped) gradient the optimiser consumed
    gathered = [torch.zeros_like(w_dp) for _ in range(world)]
    dist.all_gather(gathered, w_dp)
    if rank == 0:
        for r in range(1, world):
            assert torch.equal(gathered[0], gathered[r]), f'rank {r} diverged from rank 0'
        ref = DataParallelTrainer(gio.build_model('db5', dev), lr=1e-3, weight_decay=1e-4, clip=100.0, world=1)
        g, t = batch(range(per * world))
        ref.step(g, t)
        torch.cuda.synchronize()
        gmax = ref.flat_g.abs().max().item()
        gdiff = (ref.flat_g - g_dp).abs().max().item()
        wdiff = (ref.flat_w - w_dp).abs().max().item()
        print(f'ddp_check: world {world}: max |g_dp - g_global_batch| = {gdiff:.3e} (max |g| = {gmax:.3e}); '
              f'max |w_dp - w_global_batch| = {wdiff:.3e} (Adam step 1 moves every weight by ~lr = 1e-3 whatever its '
              f'gradient, so near-zero gradients amplify rounding)', flush=True)
        assert gdiff <= 2e-5 * gmax, (gdiff, gmax)
        assert wdiff < 5e-4, wdiff
        print('ddp_check: OK', flush=True)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == '__main__':
    main()
