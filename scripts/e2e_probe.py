import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import torch
import golden_io as gio
from equidock_public_b200 import hetero_graph as hg, synthetic
from equidock_public_b200.serving import PipelinedInference
from equidock_public_b200.engine import GraphPlan
dev = torch.device('cuda:0')
model = gio.build_model('dips', dev)
B = 256
host = hg.batch_pairs(synthetic.to_torch_pairs(synthetic.synthetic_batch(B))).pin_memory()
devb = host.to(dev)
for _ in range(3): model(devb, 0)
torch.cuda.synchronize()
def T(fn, n=10):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(n): fn()
    torch.cuda.synchronize(); return (time.perf_counter() - t0) / n * 1e3
print('H2D only (to, non_blocking): %.2f ms' % T(lambda: host.to(dev, non_blocking=True)))
print('plan build: %.2f ms' % T(lambda: GraphPlan.from_graph(devb, dev, 10)))
print('forward sync (cached plan): %.2f ms' % T(lambda: model(devb, 0)))
def fresh():
    g = host.to(dev, non_blocking=True); model(g, 0)
print('H2D + plan + forward, serial: %.2f ms' % T(fresh))
pipe = PipelinedInference(model, dev)
def piped():
    last = None
    for r in pipe.run(host for _ in range(10)): last = r
    last['_event'].synchronize()
t = T(piped, 2) / 10
print('pipelined: %.2f ms/step -> %.0f pairs/s' % (t, B / t * 1e3))
pend = None
def asyncfw():
    global pend
    nx = model.forward_async(devb, 0)
    if pend is not None: pend.result()
    pend = nx
print('forward_async lagged (device-resident): %.2f ms' % T(asyncfw, 20))
import subprocess; print(subprocess.run(['nvidia-smi', '--query-gpu=clocks.sm,clocks.max.sm,power.draw,pcie.link.gen.current,pcie.link.width.current', '--format=csv,noheader'], capture_output=True, text=True).stdout)
