#!/usr/bin/env python
"""Per-pair error table of the shipped build on every golden fixture (run on the GPU box):
max|coords - reference fp64|, the pair's yardstick |reference fp32 - reference fp64|, rotation / translation errors.
Writes gpurun_out/parity_table.txt (copied to profiles/ per round)."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, 'tests'), os.path.join(ROOT, 'oracle')):
    sys.path.insert(0, p)
import golden_io as gio  # noqa: E402


def main():
    dev = torch.device('cuda:0')
    rows = ['ds     pair                       N_l   N_r   err_coords  yard(ref32-ref64)  err/ max(1e-4,yard)  err_R      err_t']
    worst = 0.0
    for ds in ('db5', 'dips'):
        model = gio.build_model(ds, dev)
        names, pairs, outs, _ = gio.load_pairs(ds)
        for n in names:
            coors, kl, kr, rot, tr = model(gio.make_batch([pairs[n]], dev), epoch=0)
            r64, r32 = outs[n]['ref64'], outs[n]['ref32']
            yard = float(np.abs(r32['ligand_coors'] - r64['ligand_coors']).max())
            err = float(np.abs(coors[0].cpu().numpy() - r64['ligand_coors']).max())
            er = float(np.abs(rot[0].cpu().numpy() - r64['rotation']).max())
            et = float(np.abs(tr[0].cpu().numpy() - r64['translation']).max())
            ratio = err / max(1e-4, yard)
            worst = max(worst, ratio)
            rows.append(f'{ds:6s} {n:26s} {pairs[n][0]["x"].shape[0]:5d} {pairs[n][1]["x"].shape[0]:5d}  {err:.3e}   {yard:.3e}'
                        f'          {ratio:5.2f}                {er:.2e}  {et:.2e}')
    rows.append(f'worst err / max(1e-4, yard) = {worst:.3f}  (test bound: 1.0)')
    out = os.path.join(ROOT, 'gpurun_out')
    os.makedirs(out, exist_ok=True)
    with open(os.path.join(out, 'parity_table.txt'), 'w') as fh:
        fh.write('\n'.join(rows) + '\n')
    print('\n'.join(rows))


if __name__ == '__main__':
    main()
