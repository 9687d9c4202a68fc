"""Device-side evaluation metric: a drop-in for the reference's ``Meter_Unbound_Bound`` (src/utils/eval.py) whose
``update_rmsd`` of a whole batch is ONE kernel launch (csrc/head.cu ``rmsd_meter_kernel``) instead of a numpy SVD per pair
on the host (src/train.py:136-140 calls it from the training loop)."""
from __future__ import annotations

import ctypes as C
from typing import List, Sequence

import numpy as np
import torch

from . import _native as nat
from .engine import GraphPlan


class Meter_Unbound_Bound(object):
    def __init__(self):
        self.complex_rmsd_list: List[float] = []
        self.ligand_rmsd_list: List[float] = []
        self.receptor_rmsd_list: List[float] = []

    def update_rmsd_batch(self, plan: GraphPlan, ligand_coors_pred, receptor_coors_pred, ligand_coors_true, receptor_coors_true):
        """All pairs of a batch at once.  Arguments: per-pair lists (as the model returns them) or concatenated (N,3)
        tensors in batch order, on the plan's CUDA device.  Returns the (B,3) fp64 tensor [complex, ligand, receptor]."""
        dev = plan.device
        cat = lambda a: (torch.cat(list(a)) if isinstance(a, (list, tuple)) else a).detach().to(device=dev, dtype=torch.float32).contiguous()
        lp, rp, lt, rt = map(cat, (ligand_coors_pred, receptor_coors_pred, ligand_coors_true, receptor_coors_true))
        assert lp.shape == (plan.N_l, 3) and lt.shape == (plan.N_l, 3) and rp.shape == (plan.N_r, 3) and rt.shape == (plan.N_r, 3)
        out = torch.empty(plan.n_pairs, 3, dtype=torch.float64, device=dev)
        with torch.cuda.device(dev):
            st = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
            nat.check(nat.load().eqd_rmsd_meter(C.byref(plan.struct), nat.ptr(lp), nat.ptr(rp), nat.ptr(lt), nat.ptr(rt),
                                                nat.ptr(out), st), 'eqd_rmsd_meter')
        host = out.cpu().numpy()
        self.complex_rmsd_list += host[:, 0].tolist()
        self.ligand_rmsd_list += host[:, 1].tolist()
        self.receptor_rmsd_list += host[:, 2].tolist()
        return out

    def update_rmsd(self, ligand_coors_pred, receptor_coors_pred, ligand_coors_true, receptor_coors_true):
        """The reference's per-pair signature (eval.py:19): one pair = a batch of one."""
        dev = ligand_coors_pred.device
        n_l, n_r = int(ligand_coors_pred.shape[0]), int(receptor_coors_pred.shape[0])
        z = torch.zeros(0, dtype=torch.int32, device=dev)
        he = torch.zeros(0, 27, device=dev)
        plan = GraphPlan([n_l], [n_r], z, z, z, z, he, he, dev)
        return float(self.update_rmsd_batch(plan, ligand_coors_pred, receptor_coors_pred, ligand_coors_true, receptor_coors_true)[0, 0].item())

    def summarize(self, reduction_rmsd='median'):
        f = np.mean if reduction_rmsd == 'mean' else np.median
        if reduction_rmsd not in ('mean', 'median'):
            raise ValueError('Meter_Unbound_Bound: reduction_rmsd mis specified!')
        # (ligand, receptor, complex): the reference's return order (eval.py:67)
        return f(np.array(self.ligand_rmsd_list)), f(np.array(self.receptor_rmsd_list)), f(np.array(self.complex_rmsd_list))

    def summarize_with_std(self, reduction_rmsd='median'):
        _, _, c = self.summarize(reduction_rmsd)
        return c, np.std(np.array(self.complex_rmsd_list))
