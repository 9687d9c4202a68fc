"""GPU: A/B of two builds of libeqd_iegmn.so (same ABI): whole-forward outputs on the golden pairs of both checkpoints and on the
bench batch must be BITWISE equal; per-stage times from the engine's stage events.
usage: forward_ab.py <lib_a.so> <lib_b.so>      (child mode: forward_ab.py --child <out.npz>)"""
import os, sys, subprocess, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import numpy as np


def child(out):
    import torch
    import golden_io as gio
    from equidock_public_b200 import synthetic
    dev = torch.device('cuda:0')
    res = {}
    for ds in ('db5', 'dips'):
        model = gio.build_model(ds, dev)
        names, pairs, _, _ = gio.load_pairs(ds)
        coors, kl, kr, rot, trans = model(gio.make_batch([pairs[n] for n in names], dev), epoch=0)
        for i, n in enumerate(names):
            res[f'{ds}/{n}/coors'] = coors[i].cpu().numpy(); res[f'{ds}/{n}/rot'] = rot[i].cpu().numpy()
            res[f'{ds}/{n}/trans'] = trans[i].cpu().numpy(); res[f'{ds}/{n}/kl'] = kl[i].cpu().numpy()
    model = gio.build_model('dips', dev)
    g = gio.make_batch(synthetic.synthetic_batch(256), dev)
    for _ in range(3):
        coors, kl, kr, rot, trans = model(g, epoch=0)
    torch.cuda.synchronize()
    ts = []
    for _ in range(20):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); o = model(g, epoch=0); b.record(); b.synchronize(); ts.append(a.elapsed_time(b))
    res['bench/coors'] = torch.cat([c.reshape(-1) for c in coors]).cpu().numpy(); res['bench/rot'] = torch.stack(list(rot)).cpu().numpy()
    res['bench/ms_eager'] = np.array(np.median(ts))
    np.savez(out, **res)


if sys.argv[1] == '--child':
    child(sys.argv[2])
else:
    outs = []
    for i, lib in enumerate(sys.argv[1:3]):
        out = f'/tmp/forward_ab_{i}.npz'
        env = dict(os.environ, EQD_LIB_PATH=os.path.abspath(lib))
        subprocess.run([sys.executable, os.path.abspath(__file__), '--child', out], check=True, env=env)
        outs.append(np.load(out))
    a, b = outs
    bad = [k for k in a.files if k != 'bench/ms_eager' and not np.array_equal(a[k], b[k])]
    print(f'{len(a.files) - 1} arrays compared, bitwise different: {len(bad)} {bad[:5]}')
    print(f'eager forward of the bench batch: {float(a["bench/ms_eager"]):.3f} ms ({sys.argv[1]}) vs {float(b["bench/ms_eager"]):.3f} ms ({sys.argv[2]})')
    sys.exit(1 if bad else 0)
