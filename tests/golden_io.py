"""Loaders for the committed fixtures under tests/golden/ (written by oracle/make_golden.py)."""
from __future__ import annotations

import json
import os

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')
IN_KEYS_L = ['src', 'dst', 'he', 'res_feat', 'x', 'new_x', 'mu_r_norm']
IN_KEYS_R = ['src', 'dst', 'he', 'res_feat', 'x', 'mu_r_norm']
OUT_KEYS = ['ligand_coors', 'keypts_ligand', 'keypts_receptor', 'rotation', 'translation']


def load_args(ds):
    with open(os.path.join(GOLDEN, f'{ds}_args.json')) as fh:
        a = json.load(fh)
    a['debug'] = False
    return a


def load_checkpoint(ds):
    return dict(np.load(os.path.join(GOLDEN, f'{ds}_checkpoint.npz')))


_CACHE = {}


def load_pairs(ds):
    """-> (names, {name: (lig_dict, rec_dict)}, {name: {tag: {key: array}}}, raw npz)."""
    if ds in _CACHE:
        return _CACHE[ds]
    z = np.load(os.path.join(GOLDEN, f'{ds}_pairs.npz'))
    names = [str(n) for n in z['names']]
    pairs, outs = {}, {}
    for n in names:
        pairs[n] = ({k: z[f'{n}/lig/{k}'] for k in IN_KEYS_L}, {k: z[f'{n}/rec/{k}'] for k in IN_KEYS_R})
        outs[n] = {tag: {k: z[f'{n}/{tag}/{k}'] for k in OUT_KEYS + ['x_out_ligand', 'x_out_receptor',
                                                                     'h_out_ligand', 'h_out_receptor']}
                   for tag in ('ref32', 'ref64')}
        outs[n]['pdb'] = {'rotation': z[f'{n}/pdb/rotation'], 'translation': z[f'{n}/pdb/translation']}
        if f'{n}/ca/ligand_in' in z:
            outs[n]['ca'] = {k: z[f'{n}/ca/{k}'] for k in ('ligand_in', 'ligand_gt', 'receptor_gt')}
    _CACHE[ds] = (names, pairs, outs, z)
    return _CACHE[ds]


def summary():
    with open(os.path.join(GOLDEN, 'summary.json')) as fh:
        return json.load(fh)


def make_batch(pair_list, device=None):
    """[(lig_dict, rec_dict) numpy] -> PairGraphBatch (torch tensors, optionally moved to device)."""
    import torch
    from equidock_public_b200 import hetero_graph as hg
    tp = [tuple({k: torch.from_numpy(np.ascontiguousarray(v)) for k, v in d.items()} for d in p) for p in pair_list]
    g = hg.batch_pairs(tp)
    return g.to(device) if device is not None else g


def build_model(ds, device, sd=None, args=None):
    import torch
    from equidock_public_b200.rigid_docking_model import Rigid_Body_Docking_Net
    args = dict(args or load_args(ds))
    args['device'] = device
    model = Rigid_Body_Docking_Net(args, log=print)
    sd = sd if sd is not None else load_checkpoint(ds)
    model.load_state_dict({k: torch.from_numpy(np.asarray(v)) for k, v in sd.items()}, strict=True)
    return model.to(device).eval()


_ALL = {}


def load_all(ds):
    """Compact fixtures of ALL shipped test pairs of a set (tests/golden/{ds}_all.npz, oracle/make_golden_all.py):
    -> (names, {name: {'lig': protein, 'rec': protein, 'ref64': {...}, 'ref32': {...}, 'pdb': {...}, 'ca': {...}, 'yard': float}})
    where protein = {'atoms', 'atom_ptr', 'nca_c', 'res_feat'} (the format of oracle/graph_oracle.py)."""
    if ds in _ALL:
        return _ALL[ds]
    z = np.load(os.path.join(GOLDEN, f'{ds}_all.npz'))
    names = [str(n) for n in z['names']]
    out = {}
    for n in names:
        e = {}
        for side in ('lig', 'rec'):
            e[side] = {'atoms': z[f'{n}/{side}/atoms'], 'atom_ptr': z[f'{n}/{side}/atom_ptr'], 'nca_c': z[f'{n}/{side}/nca_c'],
                       'res_feat': z[f'{n}/{side}/res_feat'].astype(np.float32).reshape(-1, 1)}
            e[side]['bound_ca'] = e[side]['nca_c'][:, 1].copy()
        for tag in ('ref64', 'ref32', 'pdb'):
            e[tag] = {'rotation': z[f'{n}/{tag}/rotation'], 'translation': z[f'{n}/{tag}/translation']}
        e['ca'] = {k: z[f'{n}/ca/{k}'] for k in ('ligand_in', 'ligand_gt', 'receptor_gt')}
        e['yard'] = float(z[f'{n}/yard'])
        out[n] = e
    _ALL[ds] = (names, out)
    return _ALL[ds]
