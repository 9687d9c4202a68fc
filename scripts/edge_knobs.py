"""GPU: time the tensor-core edge stage alone (one 64-wide layer of the dips checkpoint, the bench's 256 x (200+200) batch)
for a list of EQD_EDGE_STAGGER values (start offset of tile group 1, SM cycles).  The inputs (he ~123 MB, proj ~131 MB) are
larger than L2; 30 launches per setting, CUDA events on the launching stream, median."""
import os, sys, ctypes as C
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import numpy as np, torch
import golden_io as gio
from equidock_public_b200 import _native as nat, synthetic
from equidock_public_b200.engine import GraphPlan
dev = torch.device('cuda:0')
lib = nat.load()
model = gio.build_model('dips', dev)
pairs = synthetic.synthetic_batch(256)
g = gio.make_batch(pairs, dev)
plan = GraphPlan.from_graph(g, dev, 10)
lay = model.iegmn_original.iegmn_layers[1].packed(dev)
N = plan.N
torch.manual_seed(0)
proj = torch.randn(N, 128 + 3 * 64, device=dev)
x = (torch.randn(N, 3, device=dev, dtype=torch.float64) * 5)
aggr = torch.zeros(N, 64, device=dev); xo = torch.zeros(N, 3, device=dev, dtype=torch.float64)
st = torch.zeros(plan.n_pairs + 1, dtype=torch.int32, device=dev)
stream = torch.cuda.current_stream().cuda_stream
def run():
    rc = lib.eqd_edge_stage(C.byref(plan.struct), C.byref(lay.struct), nat.ptr(proj), nat.ptr(x), nat.ptr(x), nat.ptr(aggr),
                            nat.ptr(xo), nat.ptr(st), stream)
    assert rc == 0
ref = None
for val in [int(v) for v in (sys.argv[1:] or ['0', '2000', '4000', '6000', '8000', '12000', '0'])]:
    os.environ['EQD_EDGE_STAGGER'] = str(val)
    for _ in range(5):
        run()
    torch.cuda.synchronize()
    ts = []
    for _ in range(30):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); run(); b.record(); b.synchronize()
        ts.append(a.elapsed_time(b) * 1e3)
    if ref is None:
        ref = (aggr.clone(), xo.clone())
    same = torch.equal(ref[0], aggr) and torch.equal(ref[1], xo)
    print(f'stagger {val:6d} cycles: median {np.median(ts):8.1f} us  min {min(ts):8.1f}  max {max(ts):8.1f}  '
          f'nodes {N} edges {plan.E}  bitwise-same-as-first {same}', flush=True)
