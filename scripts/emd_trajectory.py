"""GPU: a short training trajectory on the `train` bench batch; per step: wall time (synchronised), EMD augmentations / rounds."""
import os, sys, argparse, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import numpy as np, torch
import bench, bench_train
import golden_io as gio
from equidock_public_b200 import hetero_graph as hg, synthetic
from equidock_public_b200.losses import PocketBatch
from equidock_public_b200.training import DataParallelTrainer
args = argparse.Namespace(workload='train', pairs_per_gpu=bench.WORKLOADS['train']['pairs_per_gpu'], seed=0, gpus=1)
dev = torch.device('cuda:0')
triples, (lo, hi), sizes = bench_train.make_train_pairs(args, 0, 1, bench)
wl = bench.WORKLOADS['train']
sd, margs = gio.load_checkpoint(wl['ckpt']), gio.load_args(wl['ckpt'])
model = gio.build_model(wl['ckpt'], dev, sd=sd, args=margs)
trainer = DataParallelTrainer(model, lr=1e-4, weight_decay=1e-4, clip=100.0, world=1)
batch = hg.batch_pairs(synthetic.to_torch_pairs([(t[0], t[1]) for t in triples])).to(dev)
tl = lambda key: [torch.from_numpy(t[2][key]) for t in triples]
tgt = PocketBatch(tl('bound_lig'), tl('bound_rec'), tl('pocket_lig'), tl('pocket_rec'), dev)
npk = np.array([t[2]['pocket_lig'].shape[0] for t in triples])
for it in range(int(sys.argv[1]) if len(sys.argv) > 1 else 40):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    out = trainer.step(batch, tgt)
    torch.cuda.synchronize(); dt = (time.perf_counter() - t0) * 1e3
    p = out['parts'].cpu().numpy()
    aug = np.floor(p[:, 3]); pops = (p[:, 3] - aug) * 1e9
    j = int(np.argmax(pops))
    print(f'step {it:3d}: {dt:7.1f} ms  loss {float(out["loss"][0]):9.3f}  ot {p[:, 1].sum():10.3f}  aug total {int(aug.sum()):6d} max {int(aug.max()):5d}  '
          f'rounds total {int(pops.sum()):7d} max {int(round(pops.max())):6d} (pocket {npk[j]})', flush=True)
