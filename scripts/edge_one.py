"""GPU: launch one build variant of the edge stage a few times on the bench batch (for `ncu -k regex:edge_stage_tc`)."""
import os, sys, ctypes as C
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import torch
import golden_io as gio
from equidock_public_b200 import _native as nat, synthetic
from equidock_public_b200.engine import GraphPlan
dev = torch.device('cuda:0')
nat.load()
lib = C.CDLL(os.path.join(ROOT, 'variants', f'libedge_{sys.argv[1]}.so'))
lib.eqd_edge_stage.restype = C.c_int
lib.eqd_edge_stage.argtypes = [C.c_void_p] * 9
model = gio.build_model('dips', dev)
g = gio.make_batch(synthetic.synthetic_batch(256), dev)
plan = GraphPlan.from_graph(g, dev, 10)
lay = model.iegmn_original.iegmn_layers[1].packed(dev)
N = plan.N
torch.manual_seed(0)
proj = torch.randn(N, 128 + 3 * 64, device=dev)
x = (torch.randn(N, 3, device=dev, dtype=torch.float64) * 5)
st = torch.zeros(plan.n_pairs + 1, dtype=torch.int32, device=dev)
aggr = torch.zeros(N, 64, device=dev); xo = torch.zeros(N, 3, device=dev, dtype=torch.float64)
for _ in range(int(sys.argv[2]) if len(sys.argv) > 2 else 5):
    rc = lib.eqd_edge_stage(C.addressof(plan.struct), C.addressof(lay.struct), nat.ptr(proj), nat.ptr(x), nat.ptr(x), nat.ptr(aggr),
                            nat.ptr(xo), nat.ptr(st), None)
    assert rc == 0
torch.cuda.synchronize()
print('ok', float(aggr.abs().sum()), float(xo.abs().sum()))
