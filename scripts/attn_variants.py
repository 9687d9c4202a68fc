"""GPU: time build variants of the tensor-core attention kernel (64-wide layers) on the bench batch and check their outputs bitwise
against the first one.  Variant libraries: variants/libattn_<v>.so built from csrc/attn_tc.cu (-DATTN_OPT=<v>); libattn_prof.so
(-DATTN_PROF) additionally reports the SM cycles thread 0 of CTA 0 spends in each phase of the tile loop."""
import os, sys, glob, ctypes as C
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import numpy as np, torch
import golden_io as gio
from equidock_public_b200 import _native as nat, synthetic
from equidock_public_b200.engine import GraphPlan
dev = torch.device('cuda:0')
main = nat.load()
g = gio.make_batch(synthetic.synthetic_batch(256), dev)
plan = GraphPlan.from_graph(g, dev, 10)
N = plan.N
torch.manual_seed(0)
proj = torch.randn(N, 320, device=dev) * 0.5
kv = torch.zeros(main.eqd_kv_blocks_bytes(N), dtype=torch.uint8, device=dev)
nat.check(main.eqd_kv_blocks(C.byref(plan.struct), nat.ptr(proj), 320, 192, 256, nat.ptr(kv), None), 'kv_blocks')
torch.cuda.synchronize()
stream = torch.cuda.current_stream().cuda_stream
libs = {}
for path in sorted(glob.glob(os.path.join(ROOT, 'variants', 'libattn_*.so'))):
    name = os.path.basename(path)[8:-3]
    lib = C.CDLL(path)
    lib.eqd_attention_tc.restype = C.c_int
    lib.eqd_attention_tc.argtypes = [C.c_void_p] * 5
    libs[name] = lib
order = sorted(libs, key=lambda s: (s == 'prof', len(s), s))
outs, times = {}, {k: [] for k in order}
def run(lib, mu):
    rc = lib.eqd_attention_tc(C.addressof(plan.struct), nat.ptr(proj), nat.ptr(kv), nat.ptr(mu), stream)
    assert rc == 0, rc
for k in order:
    mu = torch.zeros(N, 64, device=dev)
    for _ in range(3):
        run(libs[k], mu)
    torch.cuda.synchronize()
    outs[k] = mu
for rnd in range(5):
    for k in order:
        for _ in range(20):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); run(libs[k], outs[k]); b.record(); b.synchronize()
            times[k].append(a.elapsed_time(b) * 1e3)
base = order[0]
for k in order:
    t = np.array(times[k])
    print(f'variant {k:>5s}: median {np.median(t):7.1f} us  p10 {np.percentile(t, 10):7.1f}  p90 {np.percentile(t, 90):7.1f}  '
          f'vs {base}: {np.median(t) / np.median(times[base]):.4f}  bitwise == {base}: {torch.equal(outs[k], outs[base])}  '
          f'max|diff| {float((outs[k] - outs[base]).abs().max()):.3e}', flush=True)
if 'prof' in libs:
    lib = libs['prof']
    run(lib, outs['prof']); torch.cuda.synchronize()
    buf = (C.c_longlong * 16)()
    lib.eqd_attn_prof_read.argtypes = [C.c_void_p]
    assert lib.eqd_attn_prof_read(C.addressof(buf)) == 0
    v = np.array(list(buf), dtype=np.float64)
    names = ['metadata', 'Q->TMEM+bar', 'p1 K wait', 'p1 issue+MMA wait', 'p1 ld+max+bar', 'p2 K wait', 'p2 issue+S MMA wait', 'p2 ld+exp',
             'p2 bar1', 'p2 P store+bar2', 'p2 V wait+PV MMA wait', 'p2 O ld+acc+bar3', 'mu/l store+bar', '-', '-', 'loop']
    print('phase cycles of thread 0 / CTA 0 over its tiles (total %.0f):' % v.sum())
    for nme, c in zip(names, v):
        if c: print(f'  {nme:24s} {c:10.0f}  {100 * c / v.sum():5.1f}%')
