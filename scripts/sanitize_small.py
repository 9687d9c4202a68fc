"""A small multi-tile forward for compute-sanitizer (synccheck / racecheck / memcheck)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import torch
import golden_io as gio
from equidock_public_b200 import hetero_graph as hg, synthetic
dev = torch.device('cuda:0')
model = gio.build_model('dips', dev)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 180
batch = hg.batch_pairs(synthetic.to_torch_pairs(synthetic.synthetic_batch(n, 200, 200, 10, seed=0))).to(dev)
out = model(batch, 0)
torch.cuda.synchronize()
print('done', float(torch.cat(out[0]).double().sum()))
