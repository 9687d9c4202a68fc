"""GPU: run the on-device graph construction of the bench batch (256 x (200+200) residues, all-atom inputs) a few times;
under `ncu --metrics gpu__time_duration.sum` this gives the per-kernel times of the f2 stage."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import numpy as np, torch
from equidock_public_b200 import synthetic, graph_build as gb
dev = torch.device('cuda:0')
rng = np.random.default_rng(0)
pairs = [synthetic.synthetic_residue_pair(rng, 200, 200) for _ in range(256)]
rb = gb.ResidueBatch(pairs)
inputs = {k: v.to(dev) for k, v in rb.t.items()}
for _ in range(int(sys.argv[1]) if len(sys.argv) > 1 else 3):
    g = gb.build_graphs(rb, dev, sync_sizes=False, dev_inputs=inputs)
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(5):
    g = gb.build_graphs(rb, dev, sync_sizes=False, dev_inputs=inputs)
b.record(); b.synchronize()
print('graph build of 256 pairs: %.3f ms' % (a.elapsed_time(b) / 5))
