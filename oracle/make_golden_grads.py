"""TEST INFRASTRUCTURE -- golden GRADIENTS from the reference's own autograd (authoring container only).

    python oracle/make_golden_grads.py

For one small shipped test pair per checkpoint it runs the reference's unmodified ``Rigid_Body_Docking_Net`` in fp64
WITH autograd, evaluates ``iegmn_oracle_torch.probe_loss`` (MSE on the predicted ligand coordinates + a constant-plan
transport cost on both keypoint sets: every gradient path of src/train.py:112-150 -- coordinates -> (T, b) -> 3x3 SVD
backward -> keypoints -> attention -> all IEGMN layers) and calls ``backward()`` (train.py:154).  It stores, per
parameter, the gradient's norm and its projection on a seeded random direction (fp64), and the full gradient of the
first / last layer and the small head tensors (fp32), in ``tests/golden/{ds}_grads.npz``.  These pin the backward
oracle (``TorchOracle.forward_pair_grad`` + torch.autograd, fp64): tests/test_oracle_golden.py.
"""
from __future__ import annotations

import os
import sys
import zlib

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import reference_runner as rr  # noqa: E402
from iegmn_oracle_torch import probe_loss  # noqa: E402

OUT = os.path.join(os.path.dirname(HERE), 'tests', 'golden')
PAIR = {'db5': '1QA9', 'dips': 'kq_1kq1.pdb1_2.dill'}
FULL_LIMIT = 20000   # parameters up to this many elements of the first / last layer and the head are stored in full


def direction(name, shape):
    """The seeded random direction a gradient is projected on (same function in the test)."""
    rng = np.random.default_rng(zlib.crc32(name.encode()))
    return rng.standard_normal(shape)


def targets(n_lig, coors64, seed):
    rng = np.random.default_rng(seed)
    return {'coors': coors64 + rng.normal(0, 3.0, coors64.shape), 'p_l': rng.normal(0, 20.0, (50, 3)),
            'p_r': rng.normal(0, 20.0, (50, 3)), 'w_l': rng.uniform(0.2, 1.0, 50), 'w_r': rng.uniform(0.2, 1.0, 50)}


def main():
    sys.path.insert(0, os.path.join(os.path.dirname(HERE), 'tests'))
    import golden_io as gio
    for ds, name in PAIR.items():
        args, sd = rr.load_checkpoint(ds)
        model = rr.build_reference_model(args, sd, torch.float64)
        names, pairs, outs, _ = gio.load_pairs(ds)
        pair = tuple({k: torch.as_tensor(v) for k, v in side.items()} for side in pairs[name])
        tgt = targets(pair[0]['x'].shape[0], outs[name]['ref64']['ligand_coors'].astype(np.float64), seed=7)
        prev = torch.get_default_dtype()
        torch.set_default_dtype(torch.float64)
        try:
            bg = rr.dicts_to_reference_batch([pair], torch.float64)
            for p in model.parameters():
                p.grad = None
            coors, kp_l, kp_r, rot, trans = model(bg, epoch=0)
            loss = probe_loss(coors[0], kp_l[0], kp_r[0], tgt)
            loss.backward()
        finally:
            torch.set_default_dtype(prev)
        blob = {'loss': np.float64(loss.item())}
        for k, v in tgt.items():
            blob['target/' + k] = v
        n_layers = int(args['iegmn_n_lays'])
        for pname, p in model.named_parameters():
            g = p.grad.detach().numpy().astype(np.float64) if p.grad is not None else np.zeros(tuple(p.shape))
            blob['norm/' + pname] = np.float64(np.linalg.norm(g))
            blob['proj/' + pname] = np.float64((g * direction(pname, g.shape)).sum())
            first_last = ('.iegmn_layers.0.' in pname or f'.iegmn_layers.{n_layers - 1}.' in pname
                          or '.iegmn_layers.' not in pname)
            if first_last and g.size <= FULL_LIMIT:
                blob['full/' + pname] = g.astype(np.float32)
        path = os.path.join(OUT, f'{ds}_grads.npz')
        np.savez_compressed(path, **blob)
        print(ds, name, 'loss', loss.item(), 'params', sum(1 for k in blob if k.startswith('norm/')),
              'file', os.path.getsize(path) // 1024, 'KB')


if __name__ == '__main__':
    main()
