"""CPU: the numpy graph-construction restatement (oracle/graph_oracle.py) rebuilds, from the compact all-atom fixtures,
exactly the graphs the reference's own preprocessing produced for the nine full-graph fixtures (tests/golden/*_pairs.npz,
written by oracle/make_golden.py from the reference).  oracle/make_golden_all.py asserts the same on all 125 pairs at
generation time (tests/golden/summary_all.json)."""
import json
import os

import numpy as np
import pytest

import golden_io as gio
import graph_oracle as go


@pytest.mark.parametrize('ds', ['db5', 'dips'])
def test_graph_oracle_reproduces_reference_graphs(ds):
    names, pairs, _, _ = gio.load_pairs(ds)
    _, allp = gio.load_all(ds)
    for n in names:
        if ds == 'db5' and n == '1N2C':
            continue                       # 548 + 2000 residues: ~20 s of numpy; covered at generation time
        for side, ref in (('lig', pairs[n][0]), ('rec', pairs[n][1])):
            g = go.build_graph(allp[n][side])
            assert np.array_equal(g['src'], ref['src']) and np.array_equal(g['dst'], ref['dst']), (n, side)
            assert np.abs(g['he'] - ref['he']).max() < 5e-6
            assert np.abs(g['mu_r_norm'] - ref['mu_r_norm']).max() < 5e-6
            assert np.abs(g['x'] - ref['x']).max() < 1e-5
            assert np.array_equal(g['res_feat'], ref['res_feat'])


def test_all_125_pairs_were_pinned_at_generation_time():
    with open(os.path.join(gio.GOLDEN, 'summary_all.json')) as fh:
        s = json.load(fh)
    assert len(s['db5']) == 25 and len(s['dips']) == 100
    for ds in s:
        for n, e in s[ds].items():
            d = e['graph_oracle_vs_reference']
            assert d['he'] < 5e-6 and d['mu'] < 5e-6 and d['x'] < 1e-5, (ds, n, d)


@pytest.mark.parametrize('ds,name', [('db5', '1AVX'), ('dips', 'aq_4aqa.pdb1_0.dill')])
def test_centroid_bounds_of_the_pruned_neighbour_search(ds, name):
    """csrc/graph_build.cu evaluates the mean all-atom distance only for residues that the bounds
    |c_i - c_j| <= D_ij <= |c_i - c_j| + rho_i + rho_j (c = residue centroid, rho = mean atom-to-centroid distance) cannot
    exclude from the 10 nearest.  The bounds, and the pruning rule built on them (everything selected by the reference has a
    lower bound below the 10th smallest upper bound), hold on real residues."""
    _, allp = gio.load_all(ds)
    for side in ('lig', 'rec'):
        p = allp[name][side]
        atoms, ptr = np.asarray(p['atoms'], np.float64), np.asarray(p['atom_ptr'], np.int64)
        n = len(ptr) - 1
        cen = np.stack([atoms[ptr[i]:ptr[i + 1]].mean(0) for i in range(n)])
        rho = np.array([np.linalg.norm(atoms[ptr[i]:ptr[i + 1]] - cen[i], axis=1).mean() for i in range(n)])
        D = go.residue_distance_matrix_fast(np.asarray(p['atoms'], np.float32), ptr)
        dc = np.linalg.norm(cen[:, None, :] - cen[None, :, :], axis=-1)
        off = ~np.eye(n, dtype=bool)
        assert (dc[off] <= D[off] + 1e-9).all()
        assert (D[off] <= (dc + rho[:, None] + rho[None, :])[off] + 1e-9).all()
        g = go.build_graph(p)
        ub = dc + rho[:, None] + rho[None, :]
        ub[~off] = np.inf
        for i in range(n):
            nb = g['src'][g['dst'] == i]
            certain = int((ub[i] < 30.0).sum())
            thr = np.sort(ub[i])[9] if certain > 10 else 30.0
            assert (dc[i, nb] <= thr + 1e-9).all(), (side, i)
