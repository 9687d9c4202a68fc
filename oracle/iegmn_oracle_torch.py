"""TEST INFRASTRUCTURE -- the CPU *reference arm* of the bench: a PyTorch restatement of the
reference's hot path that runs the reference's own operator sequence (``nn.functional`` linear /
layer_norm / softmax / ``torch.linalg.svd`` on the host cores, fp32 by default, all host threads),
because the reference module itself needs DGL and ``/root/reference`` is not present on the GPU box.

Checked against ``iegmn_oracle.py`` (and hence against the unmodified reference) in
``tests/test_oracle_golden.py``.  Only ``bench.py`` (``cpu_baseline`` / ``--impl reference``) and
``tests/`` import it.  Line numbers cite ``/root/reference/src/model/rigid_docking_model.py``.

Like the reference it processes a batch the way ``inference_rigid.py`` does -- one pair per call
(B=1 is the reference's fastest configuration: its batched path builds a dense
(sum N_l x sum N_r) mask and gets slower with B).
"""
from __future__ import annotations

import math

import torch
import torch.nn.functional as F

SIGMAS = [1.5 ** s for s in range(15)]  # :116


def _segment_mean(values, dst, n):
    out = torch.zeros((n,) + tuple(values.shape[1:]), dtype=values.dtype)
    out.index_add_(0, dst, values)
    deg = torch.zeros(n, dtype=values.dtype).index_add_(0, dst, torch.ones(dst.shape[0], dtype=values.dtype))
    return out / deg.clamp(min=1).unsqueeze(1)  # update_all(copy_edge, mean) :274-283


class TorchOracle:
    def __init__(self, state_dict, n_layers, skip_weight_h, x_connection_init=0.0, slope=0.01, heads=50,
                 dtype=torch.float32):
        self.sd = {k: torch.as_tensor(v).to(dtype) for k, v in state_dict.items()}
        self.L, self.sk, self.eta, self.slope, self.K, self.dtype = n_layers, skip_weight_h, x_connection_init, slope, heads, dtype

    def _layer(self, li, sides):
        p = lambda k: self.sd[f'iegmn_original.iegmn_layers.{li}.{k}']
        lr = lambda t: F.leaky_relu(t, self.slope)
        q = [lr(F.linear(s['h'], p('att_mlp_Q.0.weight'))) for s in sides]          # :247-256
        k = [lr(F.linear(s['h'], p('att_mlp_K.0.weight'))) for s in sides]
        v = [F.linear(s['h'], p('att_mlp_V.0.weight')) for s in sides]
        new = []
        for i, s in enumerate(sides):
            o = 1 - i
            x, h, src, dst = s['x'], s['h'], s['src'], s['dst']
            n = x.shape[0]
            x_rel = x[src] - x[dst]                                                  # :204-205
            d2 = (x_rel ** 2).sum(1, keepdim=True)                                   # :208-209
            rbf = torch.cat([torch.exp(-d2 / sg) for sg in SIGMAS], dim=-1)          # :210
            cat = torch.cat([h[src], h[dst], s['he'], rbf], dim=-1)                  # :186, 229-231
            a = lr(F.linear(cat, p('edge_mlp.0.weight'), p('edge_mlp.0.bias')))
            a = F.layer_norm(a, (a.shape[1],), p('edge_mlp.3.weight'), p('edge_mlp.3.bias'))
            msg = F.linear(a, p('edge_mlp.4.weight'), p('edge_mlp.4.bias'))          # :236
            mu = torch.softmax(q[i] @ k[o].t(), dim=1) @ v[o]                        # :61-63 (pair block)
            c = lr(F.linear(msg, p('coors_mlp.0.weight'), p('coors_mlp.0.bias')))
            coef = F.linear(c, p('coors_mlp.4.weight'), p('coors_mlp.4.bias'))       # :263
            x_new = self.eta * s['x0'] + (1. - self.eta) * x + _segment_mean(x_rel * coef, dst, n)  # :286-292
            inp = torch.cat([h, _segment_mean(msg, dst, n), mu, s['h0']], dim=-1)    # :319-323
            hid = lr(F.linear(inp, p('node_mlp.0.weight'), p('node_mlp.0.bias')))
            hid = F.layer_norm(hid, (hid.shape[1],), p('node_mlp.3.weight'), p('node_mlp.3.bias'))
            h_new = F.linear(hid, p('node_mlp.4.weight'), p('node_mlp.4.bias'))
            if h_new.shape[1] == h.shape[1]:                                         # :332-334
                h_new = self.sk * h_new + (1. - self.sk) * h
            new.append((x_new, h_new))
        for s, (x_new, h_new) in zip(sides, new):
            s['x'], s['h'] = x_new, h_new

    @torch.no_grad()
    def forward_pair(self, lig, rec):
        return self.forward_pair_grad(lig, rec)

    def parameters_for_grad(self):
        """Turns the weights into autograd leaves (the backward ORACLE of SURVEY 8a 'backward map': parity of any
        future backward kernels is defined against torch.autograd on this restatement in fp64, itself pinned against
        the unmodified reference's autograd by tests/golden/*_grads.npz)."""
        for v in self.sd.values():
            if v.is_floating_point():
                v.requires_grad_(True)
        return self.sd

    def forward_pair_grad(self, lig, rec):
        """forward_pair without the no_grad guard: outputs stay attached to the weights' autograd graph."""
        sd, dt = self.sd, self.dtype
        emb = sd['iegmn_original.residue_emb_layer.weight']
        sides = []
        for s, ck in ((lig, 'new_x'), (rec, 'x')):
            idx = torch.as_tensor(s['res_feat']).reshape(-1).long()                  # :460
            h0 = torch.cat([emb[idx], torch.log(torch.as_tensor(s['mu_r_norm']).to(dt))], dim=1)  # :468-471
            x0 = torch.as_tensor(s[ck]).to(dt)
            sides.append({'x': x0, 'x0': x0, 'h': h0, 'h0': h0, 'he': torch.as_tensor(s['he']).to(dt),
                          'src': torch.as_tensor(s['src']).long(), 'dst': torch.as_tensor(s['dst']).long()})
        for li in range(self.L):
            self._layer(li, sides)
        l, r = sides
        g = lambda k: sd['iegmn_original.' + k]
        d = l['h'].shape[1]
        mean = lambda h: F.leaky_relu(F.linear(h, g('mlp_h_mean_ROT.0.weight'), g('mlp_h_mean_ROT.0.bias')),
                                      self.slope).mean(0, keepdim=True)              # :525, :529
        m_l, m_r = mean(l['h']), mean(r['h'])

        def keypts(hk, mq, z):                                                       # :542-548
            keys = F.linear(hk, g('att_mlp_key_ROT.0.weight')).view(-1, self.K, d).transpose(0, 1)
            qry = F.linear(mq, g('att_mlp_query_ROT.0.weight')).view(1, self.K, d).transpose(0, 1).transpose(1, 2)
            att = torch.softmax(keys @ qry / math.sqrt(d), dim=1).view(self.K, -1)
            return att @ z

        y_r, y_l = keypts(r['h'], m_l, r['x']), keypts(l['h'], m_r, l['x'])
        yr_m, yl_m = y_r.mean(0, keepdim=True), y_l.mean(0, keepdim=True)
        A = (y_r - yr_m).t() @ (y_l - yl_m)                                          # :567
        U, S, Vt = torch.linalg.svd(A)                                               # :571
        corr = torch.diag(torch.tensor([1., 1., float(torch.sign(torch.det(A.detach())))], dtype=dt))  # :586
        T = (U @ corr) @ Vt
        b = yr_m - (T @ yl_m.t()).t()                                                # :589
        return {'ligand_coors': (T @ l['x0'].t()).t() + b, 'rotation': T, 'translation': b,
                'keypts_ligand': y_l, 'keypts_receptor': y_r}


def probe_loss(coors, y_l, y_r, tgt):
    """A scalar that exercises every gradient path of the training losses (src/train.py:112-150) without POT: the MSE of
    the predicted ligand coordinates against a fixed target (train.py:114) plus a transport-like cost of both keypoint
    sets against fixed points with a CONSTANT plan (the reference detaches its OT plan, ot_utils.py:27), here the
    seeded weights `w_l`, `w_r`."""
    t = lambda k: torch.as_tensor(tgt[k]).to(coors.dtype)
    mse = ((coors - t('coors')) ** 2).mean()
    ot_l = (t('w_l') * ((y_l - t('p_l')) ** 2).sum(1)).sum() / y_l.shape[0]
    ot_r = (t('w_r') * ((y_r - t('p_r')) ** 2).sum(1)).sum() / y_r.shape[0]
    return mse + ot_l + ot_r
