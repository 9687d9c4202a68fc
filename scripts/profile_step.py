"""A short run for ncu: `--iters` forwards of the bench workload (device-resident inputs)."""
import argparse, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import torch
import golden_io as gio
from equidock_public_b200 import hetero_graph as hg, synthetic

ap = argparse.ArgumentParser()
ap.add_argument('--pairs', type=int, default=256)
ap.add_argument('--iters', type=int, default=2)
ap.add_argument('--nres', type=int, default=200)
a = ap.parse_args()
dev = torch.device('cuda:0')
model = gio.build_model('dips', dev)
batch = hg.batch_pairs(synthetic.to_torch_pairs(synthetic.synthetic_batch(a.pairs, a.nres, a.nres, 10, seed=0))).to(dev)
for _ in range(a.iters):
    model(batch, 0)
torch.cuda.synchronize()
print('done')
