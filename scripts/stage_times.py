"""Per-stage CUDA-event times of the bench workload (B=256), for quick A/B of kernel changes."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import torch, numpy as np
import golden_io as gio
from equidock_public_b200 import hetero_graph as hg, synthetic
from equidock_public_b200.engine import IEGMNEngine
import bench
dev = torch.device('cuda:0')
model = gio.build_model('dips', dev)
batch = hg.batch_pairs(synthetic.to_torch_pairs(synthetic.synthetic_batch(int(os.environ.get('PAIRS', 256))))).to(dev)
for _ in range(3): model(batch, 0)
timer = bench.StageTimer(torch)
orig = IEGMNEngine.forward
IEGMNEngine.forward = lambda self, *a, **k: orig(self, *a, stage_timer=timer, **k)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
torch.cuda.synchronize(); e0.record()
for _ in range(5): model(batch, 0)
e1.record(); torch.cuda.synchronize()
print('step ms %.3f  edge %.1f us  node %.1f us' % (e0.elapsed_time(e1) / 5, timer.mean_ms('edge_stage') * 1e3, timer.mean_ms('node_stage') * 1e3))
