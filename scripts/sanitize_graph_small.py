"""Small on-device graph construction (incl. a protein with fewer than 11 residues and two chains far apart) for compute-sanitizer."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, 'tests')):
    sys.path.insert(0, p)
import numpy as np
import torch
from equidock_public_b200 import synthetic, graph_build as gb
dev = torch.device('cuda:0')
rng = np.random.default_rng(1)
pairs = [synthetic.synthetic_residue_pair(rng, a, b) for a, b in ((7, 40), (150, 131), (260, 12))]
far = synthetic.synthetic_residue_pair(rng, 60, 60)
far[0]['atoms'][far[0]['atom_ptr'][30]:] += 100.0          # second half of the chain 100 A away: few residues inside the cutoff
far[0]['nca_c'][30:] += 100.0
if 'bound_ca' in far[0]: far[0]['bound_ca'][30:] += 100.0
pairs.append(far)
rb = gb.ResidueBatch(pairs)
g = gb.build_graphs(rb, dev, sync_sizes=True)
torch.cuda.synchronize()
print('done', g.num_nodes(), g.num_edges())
