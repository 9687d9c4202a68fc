"""TEST INFRASTRUCTURE -- drives the reference's *unmodified* hot path in this container.

Only usable where ``/root/reference`` is mounted (the authoring container; NOT the GPU box).
It puts ``oracle/dgl_shim`` (clean-room stand-ins for the un-vendored dgl / dgllife / biopandas /
ot dependencies, ``requirements.txt:5-8``) and the reference root on ``sys.path`` and imports

* ``src.model.rigid_docking_model``  (the hot path, all 696 lines),
* ``src.utils.protein_utils``        (PDB residues -> k-NN graphs, used to rebuild test inputs),
* ``src.utils.train_utils``          (``hetero_graph_from_sg_l_r_pair``, ``create_model``).

``src.utils.args`` / ``src.train`` / ``src.inference_rigid`` are never imported (argparse and
file-system side effects at import, ``src/utils/args.py:119``, ``src/train.py:22-24``); the few
lines of ``inference_rigid.py:98-205`` that matter are re-driven by :func:`run_inference_pair`.

Nothing under the product package imports this file.
"""
from __future__ import annotations

import contextlib
import glob
import io
import os
import sys

import numpy as np
import torch

REFERENCE_ROOT = os.environ.get('EQD_REFERENCE_ROOT', '/root/reference')
_SHIM = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'dgl_shim')

CHECKPOINT_GLOB = {
    # selection logic of src/inference_rigid.py:89-94
    'dips': 'checkpts/oct20_Wdec_0.0001#*Nlay_8#shrdLay_F#*/dips_model_best.pth',
    'db5': 'checkpts/oct20_Wdec_0.001#*Nlay_5#shrdLay_T#*/db5_model_best.pth',
}


def reference_available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, 'src', 'model', 'rigid_docking_model.py'))


def import_reference():
    """Returns (rigid_docking_model, protein_utils, train_utils) modules of the reference."""
    if not reference_available():
        raise RuntimeError(f'reference not mounted at {REFERENCE_ROOT}')
    for p in (_SHIM, REFERENCE_ROOT):
        if p not in sys.path:
            sys.path.insert(0, p)
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter('ignore')  # SyntaxWarning: invalid escape sequence (rigid_docking_model.py:270)
        import src.model.rigid_docking_model as rdm
        import src.utils.protein_utils as pu
        import src.utils.train_utils as tu
    return rdm, pu, tu


def load_checkpoint(dataset: str):
    (path,) = glob.glob(os.path.join(REFERENCE_ROOT, CHECKPOINT_GLOB[dataset]))
    ckpt = torch.load(path, map_location='cpu', weights_only=False)
    args = dict(ckpt['args'])
    args['debug'] = False
    args['device'] = torch.device('cpu')
    return args, ckpt['state_dict']


def build_reference_model(args, state_dict, dtype=torch.float32):
    """``create_model`` + ``load_state_dict`` + ``eval`` (inference_rigid.py:108-112).  For fp64 the
    default dtype must be switched *before* construction: the reference hard-codes default-dtype
    tensors (rigid_docking_model.py:71, 574, 578, 586)."""
    _, _, tu = import_reference()
    prev = torch.get_default_dtype()
    torch.set_default_dtype(dtype)
    try:
        model = tu.create_model(args, log=lambda *a, **k: None)
        model.load_state_dict(state_dict)
        model = model.to(dtype)
        model.eval()
    finally:
        torch.set_default_dtype(prev)
    return model


def _get_residues(pdb_filename):
    """inference_rigid.py:77-82 verbatim in behaviour (groupby sorts by key, not file order)."""
    from biopandas.pdb import PandasPdb
    df = PandasPdb().read_pdb(pdb_filename).df['ATOM']
    df = df.rename(columns={'chain_id': 'chain', 'residue_number': 'residue', 'residue_name': 'resname',
                            'x_coord': 'x', 'y_coord': 'y', 'z_coord': 'z', 'element_symbol': 'element'})
    return list(df.groupby(['chain', 'residue', 'resname']))


def test_pair_files(dataset: str, name: str):
    base = os.path.join(REFERENCE_ROOT, 'test_sets_pdb', f'{dataset}_test_random_transformed')
    return (os.path.join(base, 'random_transformed', f'{name}_l_b.pdb'),
            os.path.join(base, 'complexes', f'{name}_r_b_COMPLEX.pdb'),
            os.path.join(REFERENCE_ROOT, 'test_sets_pdb', f'{dataset}_equidock_results', f'{name}_l_b_EQUIDOCK.pdb'))


def list_test_pairs(dataset: str):
    d = os.path.join(REFERENCE_ROOT, 'test_sets_pdb', f'{dataset}_test_random_transformed', 'random_transformed')
    return sorted(f[:-len('_l_b.pdb')] for f in os.listdir(d) if f.endswith('_l_b.pdb'))


def build_pair_graphs(dataset: str, name: str, args):
    """PDB -> (ligand_graph, receptor_graph) with the reference's own preprocessing
    (inference_rigid.py:161-186)."""
    _, pu, _ = import_reference()
    lig_file, rec_file, _ = test_pair_files(dataset, name)
    with contextlib.redirect_stdout(io.StringIO()):
        ul, ur, bl, br = pu.preprocess_unbound_bound(_get_residues(lig_file), _get_residues(rec_file),
                                                     graph_nodes=args['graph_nodes'],
                                                     pos_cutoff=args['pocket_cutoff'], inference=True)
        lg, rg = pu.protein_to_graph_unbound_bound(ul, ur, bl, br, graph_nodes=args['graph_nodes'],
                                                   cutoff=args['graph_cutoff'],
                                                   max_neighbor=args['graph_max_neighbor'], one_hot=False,
                                                   residue_loc_is_alphaC=args['graph_residue_loc_is_alphaC'])
    lg.ndata['new_x'] = lg.ndata['x']
    return lg, rg


def graph_to_dict(g, with_new_x: bool):
    """Flat tensors of one protein graph (the fixture format, see tests/golden/README.md)."""
    src, dst = g.edges()
    d = {'src': src.to(torch.int32), 'dst': dst.to(torch.int32), 'he': g.edata['he'],
         'res_feat': g.ndata['res_feat'], 'x': g.ndata['x'], 'mu_r_norm': g.ndata['mu_r_norm']}
    if with_new_x:
        d['new_x'] = g.ndata['new_x']
    return {k: v.clone() for k, v in d.items()}


def dicts_to_reference_batch(pairs, dtype=torch.float32):
    """[(ligand_dict, receptor_dict), ...] -> batched shim heterograph via the reference's own
    ``hetero_graph_from_sg_l_r_pair`` + ``dgl.batch`` (train_utils.py:61-100)."""
    import dgl
    _, _, tu = import_reference()
    hs = []
    for lig, rec in pairs:
        gs = []
        for d, is_l in ((lig, True), (rec, False)):
            g = dgl.graph(([], []), idtype=torch.int32)
            g.add_nodes(int(d['x'].shape[0]))
            g.add_edges(d['src'], d['dst'])
            g.ndata['res_feat'] = d['res_feat'].to(dtype)
            g.ndata['x'] = d['x'].to(dtype)
            g.ndata['mu_r_norm'] = d['mu_r_norm'].to(dtype)
            if is_l:
                g.ndata['new_x'] = d['new_x'].to(dtype)
            g.edata['he'] = d['he'].to(dtype)
            gs.append(g)
        hs.append(tu.hetero_graph_from_sg_l_r_pair(gs[0], gs[1]))
    return dgl.batch(hs)


@torch.no_grad()
def run_reference(model, pairs, dtype=torch.float32):
    """Reference ``Rigid_Body_Docking_Net.forward`` on a list of pair dicts.  Returns a dict of
    lists of numpy arrays plus the last-layer node states the model writes into the graph
    (rigid_docking_model.py:507-510)."""
    prev = torch.get_default_dtype()
    torch.set_default_dtype(dtype)
    try:
        bg = dicts_to_reference_batch(pairs, dtype)
        coors, kp_l, kp_r, rot, trans = model(bg, epoch=0)
    finally:
        torch.set_default_dtype(prev)
    f = lambda lst: [t.detach().cpu().numpy() for t in lst]
    return {'ligand_coors': f(coors), 'keypts_ligand': f(kp_l), 'keypts_receptor': f(kp_r),
            'rotation': f(rot), 'translation': f(trans),
            'x_out_ligand': bg.nodes['ligand'].data['x_iegmn_out'].cpu().numpy(),
            'x_out_receptor': bg.nodes['receptor'].data['x_iegmn_out'].cpu().numpy(),
            'h_out_ligand': bg.nodes['ligand'].data['hv_iegmn_out'].cpu().numpy(),
            'h_out_receptor': bg.nodes['receptor'].data['hv_iegmn_out'].cpu().numpy()}


def read_all_atoms(pdb_path) -> np.ndarray:
    from biopandas.pdb import PandasPdb
    df = PandasPdb().read_pdb(pdb_path).df['ATOM']
    return df[['x_coord', 'y_coord', 'z_coord']].to_numpy().astype(np.float64)


def read_ca_atoms(pdb_path) -> np.ndarray:
    from biopandas.pdb import PandasPdb
    df = PandasPdb().read_pdb(pdb_path).df['ATOM']
    return df[df['atom_name'] == 'CA'][['x_coord', 'y_coord', 'z_coord']].to_numpy().astype(np.float64)


def golden_rigid_from_pdbs(dataset: str, name: str):
    """(R*, t*) of the reference's shipped output PDB w.r.t. its input PDB, by all-atom Kabsch;
    returns (R, t, residual_max)."""
    lig_file, _, out_file = test_pair_files(dataset, name)
    P, Q = read_all_atoms(lig_file), read_all_atoms(out_file)
    pc, qc = P.mean(0), Q.mean(0)
    H = (P - pc).T @ (Q - qc)
    U, _, Vt = np.linalg.svd(H)
    D = np.diag([1., 1., np.sign(np.linalg.det(Vt.T @ U.T))])
    R = Vt.T @ D @ U.T
    t = qc - R @ pc
    return R, t, float(np.abs((R @ P.T).T + t - Q).max())
