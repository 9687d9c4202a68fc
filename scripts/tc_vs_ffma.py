"""GPU: tensor-core edge stage vs its FFMA twin on the same inputs (one layer), then the sanity run."""
import os, sys, ctypes as C
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import numpy as np, torch
import golden_io as gio
from equidock_public_b200 import _native as nat, hetero_graph as hg, synthetic
from equidock_public_b200.engine import GraphPlan
dev = torch.device('cuda:0')
lib = nat.load()
model = gio.build_model('dips', dev)
for desc, pairs in (('golden dips', [gio.load_pairs('dips')[1][n] for n in gio.load_pairs('dips')[0]]),
                    ('synthetic 16x(200+200)', synthetic.synthetic_batch(16))):
    g = gio.make_batch(pairs, dev)
    plan = GraphPlan.from_graph(g, dev, 10)
    lay = model.iegmn_original.iegmn_layers[1].packed(dev)
    N = plan.N
    torch.manual_seed(0)
    proj = torch.randn(N, 128 + 3 * 64, device=dev)
    x = (torch.randn(N, 3, device=dev, dtype=torch.float64) * 5)
    outs = []
    for fn in (lib.eqd_edge_stage_ffma, lib.eqd_edge_stage):
        aggr = torch.zeros(N, 64, device=dev); xo = torch.zeros(N, 3, device=dev, dtype=torch.float64)
        st = torch.zeros(plan.n_pairs + 1, dtype=torch.int32, device=dev)
        rc = fn(C.byref(plan.struct), C.byref(lay.struct), nat.ptr(proj), nat.ptr(x), nat.ptr(x), nat.ptr(aggr), nat.ptr(xo), nat.ptr(st), None)
        torch.cuda.synchronize()
        outs.append((rc, aggr, xo, st))
    (r0, a0, x0, s0), (r1, a1, x1, s1) = outs
    print(desc, 'rc', r0, r1, 'status', s0[-1].item(), s1[-1].item(),
          'aggr max|diff|', (a0 - a1).abs().max().item(), 'max|aggr|', a0.abs().max().item(),
          'x max|diff|', (x0 - x1).abs().max().item(), flush=True)
