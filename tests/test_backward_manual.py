"""CPU: the hand-derived backward (oracle/backward_manual.py, the stage-by-stage checker of the CUDA backward kernels)
against torch.autograd on the torch restatement in fp64 (pinned against the reference's own autograd by
tests/golden/*_grads.npz), for both checkpoints (5 shared layers / 8 layers) incl. the 3x3 SVD backward."""
import numpy as np
import pytest
import torch

import backward_manual as bm
import golden_io as gio
import iegmn_oracle as orc
import iegmn_oracle_torch as ot

PAIR = {'db5': '1QA9', 'dips': 'kq_1kq1.pdb1_2.dill'}


@pytest.mark.parametrize('ds', ['db5', 'dips'])
def test_manual_backward_equals_autograd(ds):
    names, pairs, outs, _ = gio.load_pairs(ds)
    lig, rec = pairs[PAIR[ds]]
    sd, args = gio.load_checkpoint(ds), gio.load_args(ds)
    cfg = orc.OracleConfig.from_args(args)
    z = np.load(gio.GOLDEN + f'/{ds}_grads.npz')
    tgt = {k[len('target/'):]: z[k] for k in z.files if k.startswith('target/')}

    def loss_grads(out):
        n = out['ligand_coors'].shape[0]
        dco = 2.0 * (out['ligand_coors'] - tgt['coors']) / (3 * n)
        dyl = 2.0 * tgt['w_l'][:, None] * (out['keypts_ligand'] - tgt['p_l']) / 50
        dyr = 2.0 * tgt['w_r'][:, None] * (out['keypts_receptor'] - tgt['p_r']) / 50
        return dco, dyl, dyr

    grads, out = bm.full_backward(sd, cfg, lig, rec, loss_grads, bool(args['shared_layers']))
    model = ot.TorchOracle(sd, cfg.n_layers, cfg.skip_weight_h, cfg.x_connection_init, cfg.slope, cfg.num_att_heads,
                           dtype=torch.float64)
    psd = model.parameters_for_grad()
    o = model.forward_pair_grad(lig, rec)
    ot.probe_loss(o['ligand_coors'], o['keypts_ligand'], o['keypts_receptor'], tgt).backward()
    assert np.abs(out['ligand_coors'] - o['ligand_coors'].detach().numpy()).max() < 1e-8
    shared = bool(args['shared_layers'])
    for name, t in psd.items():
        if not t.is_floating_point():
            continue
        ref = t.grad.numpy() if t.grad is not None else np.zeros(tuple(t.shape))
        if shared and '.iegmn_layers.' in name and int(name.split('.iegmn_layers.')[1].split('.')[0]) >= 1:
            # the torch restatement keeps one leaf per layer index; the shared module's gradient is their sum
            suffix = name.split('.iegmn_layers.')[1].split('.', 1)[1]
            ref = sum(psd[f'iegmn_original.iegmn_layers.{j}.{suffix}'].grad.numpy() for j in range(1, cfg.n_layers))
        got = grads[name]
        scale = max(np.abs(ref).max(), 1e-12)
        assert np.abs(got.reshape(ref.shape) - ref).max() <= 1e-7 * scale + 1e-12, (name, np.abs(got.reshape(ref.shape) - ref).max(), scale)
    # and against the golden gradients of the unmodified reference (norm + seeded projection)
    import zlib
    for k in z.files:
        if k.startswith('norm/'):
            pname = k[5:]
            g = grads[pname]
            assert abs(np.linalg.norm(g) - float(z[k])) <= 1e-6 * max(float(z[k]), 1e-9), pname
            d = np.random.default_rng(zlib.crc32(pname.encode())).standard_normal(g.shape)
            assert abs((g * d).sum() - float(z['proj/' + pname])) <= 1e-6 * max(float(z[k]), 1e-9), pname
