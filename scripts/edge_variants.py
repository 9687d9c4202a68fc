"""GPU: time build variants of the tensor-core edge stage against each other on the bench batch (256 x (200+200), one 64-wide
dips layer) and check that their outputs are bitwise those of variant 0.  Variant libraries: variants/libedge_<v>.so, built
from csrc/edge_stage_tc.cu with -DEDGE_OPT=<v> (only eqd_edge_stage is used from them).  Rounds are interleaved (v0 v1 .. v0 v1 ..)
so that clock / thermal drift hits all variants alike; 5 rounds x 20 launches, CUDA events, median of all."""
import os, sys, glob, ctypes as C
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import numpy as np, torch
import golden_io as gio
from equidock_public_b200 import _native as nat, synthetic
from equidock_public_b200.engine import GraphPlan
dev = torch.device('cuda:0')
nat.load()
model = gio.build_model('dips', dev)
g = gio.make_batch(synthetic.synthetic_batch(256), dev)
plan = GraphPlan.from_graph(g, dev, 10)
lay = model.iegmn_original.iegmn_layers[1].packed(dev)
N = plan.N
torch.manual_seed(0)
proj = torch.randn(N, 128 + 3 * 64, device=dev)
x = (torch.randn(N, 3, device=dev, dtype=torch.float64) * 5)
st = torch.zeros(plan.n_pairs + 1, dtype=torch.int32, device=dev)
stream = torch.cuda.current_stream().cuda_stream
libs = {}
for path in sorted(glob.glob(os.path.join(ROOT, 'variants', 'libedge_*.so'))):
    name = os.path.basename(path)[8:-3]
    lib = C.CDLL(path)
    lib.eqd_edge_stage.restype = C.c_int
    lib.eqd_edge_stage.argtypes = [C.c_void_p] * 9
    libs[name] = lib
order = sorted(libs, key=lambda s: (len(s), s))
outs, times = {}, {k: [] for k in order}
def run(lib, aggr, xo):
    rc = lib.eqd_edge_stage(C.addressof(plan.struct), C.addressof(lay.struct), nat.ptr(proj), nat.ptr(x), nat.ptr(x), nat.ptr(aggr),
                            nat.ptr(xo), nat.ptr(st), stream)
    assert rc == 0, rc
for k in order:
    aggr = torch.zeros(N, 64, device=dev); xo = torch.zeros(N, 3, device=dev, dtype=torch.float64)
    for _ in range(3):
        run(libs[k], aggr, xo)
    torch.cuda.synchronize()
    outs[k] = (aggr, xo)
for rnd in range(5):
    for k in order:
        aggr, xo = outs[k]
        for _ in range(20):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); run(libs[k], aggr, xo); b.record(); b.synchronize()
            times[k].append(a.elapsed_time(b) * 1e3)
base = order[0]
for k in order:
    same = torch.equal(outs[k][0], outs[base][0]) and torch.equal(outs[k][1], outs[base][1])
    t = np.array(times[k])
    print(f'variant {k:>4s}: median {np.median(t):7.1f} us  p10 {np.percentile(t, 10):7.1f}  p90 {np.percentile(t, 90):7.1f}  '
          f'vs {base}: {np.median(t) / np.median(times[base]):.4f}  bitwise == variant {base}: {same}', flush=True)
