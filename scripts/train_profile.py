#!/usr/bin/env python
"""Phase timing of one training step of the `train` bench workload (run on the GPU box):
    python scripts/train_profile.py                 # wall-clock per phase with synchronisation between phases
    ncu --profile-from-start off --metrics gpu__time_duration.sum --csv --log-file gpurun_out/train_launches.csv \
        python scripts/train_profile.py --ncu       # per-kernel durations of ONE step
"""
import argparse
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, 'tests'), os.path.join(ROOT, 'oracle')):
    sys.path.insert(0, p)
import bench  # noqa: E402
import bench_train  # noqa: E402
import golden_io as gio  # noqa: E402
from equidock_public_b200 import hetero_graph as hg, synthetic  # noqa: E402
from equidock_public_b200.losses import PocketBatch, device_losses  # noqa: E402
from equidock_public_b200.training import DataParallelTrainer  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--ncu', action='store_true')
    ap.add_argument('--pairs', type=int, default=32)
    a = ap.parse_args()
    args = argparse.Namespace(workload='train', pairs_per_gpu=a.pairs, seed=0)
    dev = torch.device('cuda:0')
    triples, _, _ = bench_train.make_train_pairs(args, 0, 1, bench)
    model = gio.build_model('db5', dev)
    tr = DataParallelTrainer(model, lr=1e-4, weight_decay=1e-4)
    g = hg.batch_pairs(synthetic.to_torch_pairs([(t[0], t[1]) for t in triples])).to(dev)
    tl = lambda k: [torch.from_numpy(t[2][k]) for t in triples]
    tgt = PocketBatch(tl('bound_lig'), tl('bound_rec'), tl('pocket_lig'), tl('pocket_rec'), dev)
    print('pairs', len(triples), 'nodes', g.num_nodes(), 'edges', g.num_edges(), 'pockets', [int(t[2]['pocket_lig'].shape[0]) for t in triples])
    for _ in range(2):
        tr.step(g, tgt)
    torch.cuda.synchronize()
    if a.ncu:
        torch.cuda.profiler.start()
        tr.step(g, tgt)
        torch.cuda.synchronize()
        torch.cuda.profiler.stop()
        return
    sync = torch.cuda.synchronize
    eng = tr.engine
    for rep in range(3):
        t = [time.perf_counter()]
        fwd = eng.forward(g); sync(); t.append(time.perf_counter())
        res = device_losses(fwd['plan'], fwd['ligand_coors'], fwd['keypts'], tgt, *tr.loss_args); sync(); t.append(time.perf_counter())
        tr.flat_g.zero_()
        eng.backward(fwd, res['dcoors'], res['dkeypts'], flat=tr.flat_g); sync(); t.append(time.perf_counter())
        tr.invalidate_packed(); t.append(time.perf_counter())
        t0 = time.perf_counter(); tr.step(g, tgt); sync(); whole = time.perf_counter() - t0
        if rep == 0:
            st = res['parts'][:, 3].cpu().numpy()
            pk = [int(t_[2]['pocket_lig'].shape[0]) for t_ in triples]
            print('EMD solver: (pocket, augmentations, sinks settled):', [(p_, int(v), int(round((v - int(v)) * 1e9))) for p_, v in zip(pk, st)])
        d = np.diff(t) * 1e3
        print(f'rep {rep}: forward {d[0]:.2f} ms | losses {d[1]:.2f} ms | backward {d[2]:.2f} ms | invalidate {d[3]:.2f} ms | whole step {whole * 1e3:.2f} ms')


if __name__ == '__main__':
    main()
