"""CPU, world_size 2 over gloo: the N>1 host logic of the data-parallel trainer -- the flat gradient layout and its
contiguous completion-order buckets, bucket-wise all-reduce == one all-reduce of the flat buffer, and 'average of the
per-rank gradients of per-rank mean losses == gradient of the global-batch mean loss' (src/train.py:143-150 normalises
by the batch length), checked with the fp64 torch restatement as the per-rank gradient engine."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import golden_io as gio


def test_param_layout_buckets_are_a_partition_in_completion_order():
    from equidock_public_b200.rigid_docking_model import Rigid_Body_Docking_Net
    from equidock_public_b200.training import ParamLayout
    for ds in ('db5', 'dips'):
        args = gio.load_args(ds)
        args['device'] = 'cpu'
        model = Rigid_Body_Docking_Net(args)
        lo = ParamLayout(model)
        assert lo.n_param_elements == sum(p.numel() for p in model.parameters())
        assert lo.buckets[0][0] == 'head' and lo.buckets[-1][0] == 'emb'
        assert lo.buckets[0][1] == 0 and lo.buckets[-1][2] == lo.total
        assert all(lo.buckets[i][2] == lo.buckets[i + 1][1] for i in range(len(lo.buckets) - 1))
        labels = [b[0] for b in lo.buckets[1:-1]]
        assert labels == sorted(labels, key=lambda s: -int(s[5:]))          # layers last to first
        offs = sorted((lo.offset[id(p)], p.numel()) for p in lo.params)
        assert all(o % 64 == 0 for o, _ in offs)                            # 256-byte aligned parameters
        assert all(offs[i][0] + offs[i][1] <= offs[i + 1][0] for i in range(len(offs) - 1))
        flat = torch.zeros(lo.total)
        for v, p in zip(lo.views(flat), lo.params):
            assert v.shape == p.shape


def _worker(rank, world, port, out_dir):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    import iegmn_oracle_torch as ot
    from equidock_public_b200.training import allreduce_buckets
    names, pairs, outs, _ = gio.load_pairs('dips')
    sd, args = gio.load_checkpoint('dips'), gio.load_args('dips')
    sel = [names[0], names[1]]

    def grads_of(pair_names):
        m = ot.TorchOracle(sd, args['iegmn_n_lays'], args['skip_weight_h'], args['x_connection_init'], args['leakyrelu_neg_slope'],
                           args['num_att_heads'], dtype=torch.float64)
        psd = m.parameters_for_grad()
        loss = 0.0
        for n in pair_names:
            o = m.forward_pair_grad(*pairs[n])
            loss = loss + ((o['ligand_coors'] - torch.from_numpy(outs[n]['ref64']['ligand_coors']) - 1.0) ** 2).mean() \
                + (o['keypts_ligand'] ** 2).mean() * 1e-3
        (loss / len(pair_names)).backward()                    # mean over the (local) batch, train.py:143
        keys = sorted(k for k, v in psd.items() if v.is_floating_point())
        return keys, torch.cat([(psd[k].grad if psd[k].grad is not None else torch.zeros_like(psd[k])).reshape(-1) for k in keys])

    keys, local = grads_of([sel[rank]])
    n = local.numel()
    cuts = [0, n // 3, n // 2, n]
    buckets = [(f'b{i}', cuts[i], cuts[i + 1]) for i in range(3)]
    a = local.clone()
    allreduce_buckets(a, buckets, world)
    b = local.clone()
    dist.all_reduce(b)
    assert torch.equal(a, b)                                   # bucket-wise == whole-buffer all-reduce
    a /= world                                                 # the trainer's scale_extra = 1 / world
    if rank == 0:
        _, full = grads_of(sel)                                # single process, global batch of 2
        torch.save({'dp': a, 'full': full}, os.path.join(out_dir, 'r0.pt'))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_two_rank_gradient_average_equals_global_batch_gradient(tmp_path):
    with socket.socket() as s:
        s.bind(('127.0.0.1', 0))
        port = s.getsockname()[1]
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    r = torch.load(os.path.join(str(tmp_path), 'r0.pt'))
    assert (r['dp'] - r['full']).abs().max().item() <= 1e-10 * max(1.0, r['full'].abs().max().item())
