"""TEST INFRASTRUCTURE -- compact fixtures for ALL 125 shipped test pairs (25 DB5.5 + 100 DIPS), authoring container only:

    python oracle/make_golden_all.py [db5|dips]

For every pair of ``test_sets_pdb/{db5,dips}_test_random_transformed`` it
  1. runs the reference's own, unmodified preprocessing + graph construction (src/utils/protein_utils.py via
     oracle/reference_runner.py) and extracts the COMPACT inputs the graph builder needs (all-atom coordinates per residue,
     N/CA/C, residue type) -- ~50 KB per pair instead of ~600 KB of edge features;
  2. asserts that ``oracle/graph_oracle.py`` rebuilds the reference's graphs from them (identical edges; he / mu_r_norm / x
     to fp32 rounding) -- this pins the graph oracle on all 125 pairs;
  3. runs the reference's unmodified model with the shipped checkpoint in fp32 and fp64 and stores (R, t) of both plus the
     per-pair yardstick max|coords32 - coords64|;
  4. stores the golden (R*, t*) recovered from the shipped output PDB and the C-alpha traces needed by the reference's
     C-RMSD / I-RMSD metric (src/test_all_methods/eval_pdb_outputset.py:71-109).
Output: tests/golden/{ds}_all.npz, tests/golden/summary_all.json.
"""
from __future__ import annotations

import contextlib
import io
import json
import os
import sys
import time

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import graph_oracle as go  # noqa: E402
import reference_runner as rr  # noqa: E402

OUT = os.path.join(os.path.dirname(HERE), 'tests', 'golden')


def compact_protein(residues, graph):
    atoms, ptr, nca_c = [], [0], []
    for _, df in residues:
        c = df[['x', 'y', 'z']].to_numpy().astype(np.float32)
        atoms.append(c)
        ptr.append(ptr[-1] + c.shape[0])
        pick = lambda nm: df[df['atom_name'] == nm][['x', 'y', 'z']].to_numpy().squeeze().astype(np.float32)
        nca_c.append(np.stack([pick('N'), pick('CA'), pick('C')]))
    nca_c = np.stack(nca_c)
    return {'atoms': np.concatenate(atoms), 'atom_ptr': np.asarray(ptr, np.int32), 'nca_c': nca_c,
            'res_feat': graph.ndata['res_feat'].numpy().astype(np.float32), 'bound_ca': nca_c[:, 1].copy()}


def main(which):
    _, pu, _ = rr.import_reference()
    summary_path = os.path.join(OUT, 'summary_all.json')
    summary = json.load(open(summary_path)) if os.path.isfile(summary_path) else {}
    for ds in which:
        args, sd = rr.load_checkpoint(ds)
        m32 = rr.build_reference_model(args, sd, torch.float32)
        m64 = rr.build_reference_model(args, sd, torch.float64)
        names = rr.list_test_pairs(ds)
        blob, summ = {'names': np.array(names)}, {}
        for name in names:
            t0 = time.time()
            lig_file, rec_file, out_file = rr.test_pair_files(ds, name)
            with contextlib.redirect_stdout(io.StringIO()):
                ul, ur, bl, br = pu.preprocess_unbound_bound(rr._get_residues(lig_file), rr._get_residues(rec_file),
                                                             graph_nodes=args['graph_nodes'], pos_cutoff=args['pocket_cutoff'],
                                                             inference=True)
                lg, rg = pu.protein_to_graph_unbound_bound(ul, ur, bl, br, graph_nodes=args['graph_nodes'],
                                                           cutoff=args['graph_cutoff'], max_neighbor=args['graph_max_neighbor'],
                                                           one_hot=False, residue_loc_is_alphaC=args['graph_residue_loc_is_alphaC'])
            lg.ndata['new_x'] = lg.ndata['x']
            dev = {'he': 0.0, 'mu': 0.0, 'x': 0.0}
            for side, res, g in (('lig', ul, lg), ('rec', ur, rg)):
                cp = compact_protein(res, g)
                mine = go.build_graph(cp, cutoff=args['graph_cutoff'], max_neighbor=args['graph_max_neighbor'])
                src, dst = g.edges()
                assert np.array_equal(mine['src'], src.numpy()) and np.array_equal(mine['dst'], dst.numpy()), (name, side)
                dev['he'] = max(dev['he'], float(np.abs(mine['he'] - g.edata['he'].numpy()).max()))
                dev['mu'] = max(dev['mu'], float(np.abs(mine['mu_r_norm'] - g.ndata['mu_r_norm'].numpy()).max()))
                dev['x'] = max(dev['x'], float(np.abs(mine['x'] - g.ndata['x'].numpy()).max()))
                assert np.array_equal(mine['res_feat'], g.ndata['res_feat'].numpy())
                for k in ('atoms', 'atom_ptr', 'nca_c'):
                    blob[f'{name}/{side}/{k}'] = cp[k]
                blob[f'{name}/{side}/res_feat'] = cp['res_feat'].astype(np.uint8)
            assert dev['he'] < 5e-6 and dev['mu'] < 5e-6 and dev['x'] < 1e-5, (name, dev)
            pair = (rr.graph_to_dict(lg, True), rr.graph_to_dict(rg, False))
            o32 = rr.run_reference(m32, [pair], torch.float32)
            o64 = rr.run_reference(m64, [pair], torch.float64)
            R, t, resid = rr.golden_rigid_from_pdbs(ds, name)
            blob[f'{name}/ref64/rotation'] = o64['rotation'][0].astype(np.float64)
            blob[f'{name}/ref64/translation'] = o64['translation'][0].astype(np.float64)
            blob[f'{name}/ref32/rotation'] = o32['rotation'][0].astype(np.float32)
            blob[f'{name}/ref32/translation'] = o32['translation'][0].astype(np.float32)
            blob[f'{name}/pdb/rotation'], blob[f'{name}/pdb/translation'] = R, t
            gt_lig_file = rec_file.replace('_r_b_COMPLEX', '_l_b_COMPLEX')
            blob[f'{name}/ca/ligand_in'] = rr.read_ca_atoms(lig_file).astype(np.float32)
            blob[f'{name}/ca/ligand_gt'] = rr.read_ca_atoms(gt_lig_file).astype(np.float32)
            blob[f'{name}/ca/receptor_gt'] = rr.read_ca_atoms(rec_file).astype(np.float32)
            yard = float(np.abs(o32['ligand_coors'][0].astype(np.float64) - o64['ligand_coors'][0]).max())
            blob[f'{name}/yard'] = np.float64(yard)
            summ[name] = {'n_ligand': int(pair[0]['x'].shape[0]), 'n_receptor': int(pair[1]['x'].shape[0]),
                          'graph_oracle_vs_reference': dev, 'ref_fp32_vs_fp64_coors_max_abs': yard,
                          'pdb_rigid_fit_residual': resid}
            print(ds, name, summ[name]['n_ligand'], summ[name]['n_receptor'], dev, f'yard {yard:.2e}', f'{time.time() - t0:.1f}s', flush=True)
        np.savez_compressed(os.path.join(OUT, f'{ds}_all.npz'), **blob)
        summary[ds] = summ
        with open(summary_path, 'w') as fh:
            json.dump(summary, fh, indent=1, sort_keys=True)
        print(ds, 'file', os.path.getsize(os.path.join(OUT, f'{ds}_all.npz')) // 1024, 'KB')


if __name__ == '__main__':
    main(sys.argv[1:] or ['db5', 'dips'])
