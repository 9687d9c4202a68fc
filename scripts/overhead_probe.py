import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import torch, numpy as np
import golden_io as gio
from equidock_public_b200 import hetero_graph as hg, synthetic
from equidock_public_b200.engine import IEGMNEngine
dev = torch.device('cuda:0')
model = gio.build_model('dips', dev)
for B in (256, 16):
    batch = hg.batch_pairs(synthetic.to_torch_pairs(synthetic.synthetic_batch(B))).to(dev)
    for _ in range(3): model(batch, 0)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(10): model(batch, 0)
    torch.cuda.synchronize(); t_sync = (time.perf_counter() - t0) / 10
    # async: no status check
    orig = IEGMNEngine.forward
    IEGMNEngine.forward = lambda self, *a, **k: orig(self, *a[:10], False, *a[11:], **k) if len(a) > 10 else orig(self, *a, **{**k, 'check_status': False})
    iegmn = model.iegmn_original
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter(); e0.record()
    for _ in range(10): iegmn.run_engine(batch)
    t_launch = (time.perf_counter() - t0) / 10
    e1.record(); torch.cuda.synchronize()
    t_gpu = e0.elapsed_time(e1) / 10
    IEGMNEngine.forward = orig
    print(f'B={B}: synced step {t_sync*1e3:.3f} ms | async: CPU launch {t_launch*1e3:.3f} ms/step, GPU {t_gpu:.3f} ms/step', flush=True)
import cProfile, pstats
batch = hg.batch_pairs(synthetic.to_torch_pairs(synthetic.synthetic_batch(16))).to(dev)
pr = cProfile.Profile(); pr.enable()
for _ in range(20): model(batch, 0)
pr.disable()
pstats.Stats(pr).sort_stats('cumulative').print_stats(18)
