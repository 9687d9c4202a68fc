"""Quick end-to-end check on a GPU box: engine vs the fp64 reference fixtures."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests'))
import numpy as np, torch
import golden_io as gio

dev = torch.device('cuda:0')
for ds in ('db5', 'dips'):
    names, pairs, outs, _ = gio.load_pairs(ds)
    model = gio.build_model(ds, dev)
    for n in names:
        g = gio.make_batch([pairs[n]], dev)
        t0 = time.time()
        coors, kl, kr, rot, tr = model(g, epoch=0)
        torch.cuda.synchronize()
        dt = time.time() - t0
        r64, r32 = outs[n]['ref64'], outs[n]['ref32']
        e = np.abs(coors[0].cpu().numpy() - r64['ligand_coors']).max()
        y = np.abs(r32['ligand_coors'] - r64['ligand_coors']).max()
        eh = np.abs(g.nodes['ligand'].data['hv_iegmn_out'].cpu().numpy() - r64['h_out_ligand']).max()
        ex = np.abs(g.nodes['receptor'].data['x_iegmn_out'].cpu().numpy() - r64['x_out_receptor']).max()
        ek = np.abs(kl[0].cpu().numpy() - r64['keypts_ligand']).max()
        er = np.abs(rot[0].cpu().numpy() - r64['rotation']).max()
        print(f'{ds} {n}: |coors-ref64|={e:.3e} (ref32 yardstick {y:.3e}) h={eh:.3e} x_rec={ex:.3e} keyp={ek:.3e} R={er:.3e} {dt*1e3:.1f} ms', flush=True)
    # ragged batch of all pairs
    g = gio.make_batch([pairs[n] for n in names], dev)
    coors, kl, kr, rot, tr = model(g, epoch=0)
    for i, n in enumerate(names):
        e = np.abs(coors[i].cpu().numpy() - outs[n]['ref64']['ligand_coors']).max()
        print(f'  batched {n}: {e:.3e}')
