"""TEST INFRASTRUCTURE -- the reference's TRAINING STEP on the CPU for one pair (src/train.py:98-154): forward of the
torch port, the three losses (MSE, exact EMD through the LP of loss_oracle.ot_emd -- POT's network simplex is not in this
image, so the EMD leg here is slower than the reference's; its share is reported by bench.py --, body intersection) and
``loss.backward()`` through torch.autograd.  Used by bench.py's CPU arm of the `train` workload and by the tests."""
from __future__ import annotations

import time

import numpy as np
import torch

import loss_oracle as lo

EMD_SECONDS = [0.0]


def reference_train_pair(model, pair, w_ot=1.0, w_int=10.0, sigma=25.0, ct=10.0):
    """pair = (lig, rec, targets) with targets = {'bound_lig', 'bound_rec', 'pocket_lig', 'pocket_rec'} numpy arrays."""
    lig, rec, tgt = pair
    for v in model.sd.values():
        if v.is_floating_point() and v.grad is not None:
            v.grad = None
    out = model.forward_pair_grad(lig, rec)
    dt = out['ligand_coors'].dtype
    t = lambda a: torch.as_tensor(np.asarray(a)).to(dt)
    coors, yl, yr = out['ligand_coors'], out['keypts_ligand'], out['keypts_receptor']
    mse = ((coors - t(tgt['bound_lig'])) ** 2).mean()
    cost = ((t(tgt['pocket_lig'])[:, None] - yl[None]) ** 2).sum(2) + ((t(tgt['pocket_rec'])[:, None] - yr[None]) ** 2).sum(2)
    t0 = time.perf_counter()
    _, plan, _ = lo.ot_emd(cost.detach().double().numpy())          # ot_utils.py:23-27: plan detached
    EMD_SECONDS[0] += time.perf_counter() - t0
    ot = (t(plan) * cost).sum()
    recb = t(tgt['bound_rec'])
    G = lambda prot, x: -sigma * torch.log(1e-3 + torch.exp(-((prot[None] - x[:, None]) ** 2).sum(2) / sigma).sum(1))
    inter = torch.clamp(ct - G(recb, coors), min=0).mean() + torch.clamp(ct - G(coors, recb), min=0).mean()
    loss = mse + w_ot * ot + w_int * inter
    loss.backward()
    return float(loss.item())
