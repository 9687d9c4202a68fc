"""bench.py --workload train: one full training step of the reference (BASELINE configs 3 / 4; src/train.py:88-169) per
"step": forward with stash, device losses (MSE + exact-EMD pocket OT + body intersection), CUDA backward, per-bucket NCCL
all-reduce of the flat gradient overlapped with the tail of backward, fused clip + Adam.  32 DIPS-shaped ragged pairs per
GPU, 5-layer shared IEGMN (DB5 checkpoint weights as the starting point), weak scaling (config 4 = 8 x 32 = 256 pairs)."""
from __future__ import annotations

import json
import time

import numpy as np


def make_targets(pair, rng):
    """Synthetic training labels of one pair with the reference's shapes (src/utils/db5_data.py:185-192): bound ligand /
    receptor C-alpha coordinates and N_pocket pocket points = midpoints of ligand-receptor residue pairs, the same array
    for both sides; N_pocket ~ the DIPS test distribution (7..398, median 48)."""
    lig, rec = pair
    n_p = int(np.clip(np.exp(rng.normal(np.log(48), 0.8)), 7, 398))
    bl = lig['x'].astype(np.float32)
    br = (rec['x'] + np.float32(8.0)).astype(np.float32)          # a docked pose next to the ligand
    i = rng.integers(0, bl.shape[0], n_p)
    j = rng.integers(0, br.shape[0], n_p)
    mid = (0.5 * (bl[i] + br[j])).astype(np.float32)
    return {'bound_lig': bl, 'bound_rec': br, 'pocket_lig': mid, 'pocket_rec': mid.copy()}


def make_train_pairs(args, rank, world, bench):
    pairs, (lo, hi), sizes = bench.make_pairs(args, rank, world)
    out = []
    for k, p in enumerate(pairs):
        out.append((p[0], p[1], make_targets(p, np.random.default_rng([args.seed, 7, lo + k]))))
    return out, (lo, hi), sizes


def run(args, rank, local_rank, world, bench):
    numa = bench.bind_to_gpu_numa(local_rank) if not args.no_numa_bind else None
    import torch
    import golden_io as gio
    from equidock_public_b200 import hetero_graph as hg
    from equidock_public_b200 import synthetic
    from equidock_public_b200.losses import PocketBatch, check_loss_status
    from equidock_public_b200.training import DataParallelTrainer

    torch.cuda.set_device(local_rank)
    dev = torch.device('cuda', local_rank)
    D = bench.Dist(world, dev, torch)
    wl = bench.WORKLOADS['train']
    triples, (lo, hi), sizes = make_train_pairs(args, rank, world, bench)
    B, total_pairs = len(triples), len(sizes)
    sd, margs = gio.load_checkpoint(wl['ckpt']), gio.load_args(wl['ckpt'])
    model = gio.build_model(wl['ckpt'], dev, sd=sd, args=margs)
    trainer = DataParallelTrainer(model, lr=1e-4, weight_decay=1e-4, clip=100.0, world=world,
                                  pocket_ot_loss_weight=float(margs.get('pocket_ot_loss_weight', 1.0)),
                                  intersection_loss_weight=float(margs.get('intersection_loss_weight', 10.0)),
                                  intersection_sigma=float(margs.get('intersection_sigma', 25.0)),
                                  intersection_surface_ct=float(margs.get('intersection_surface_ct', 10.0)))
    host_batch = hg.batch_pairs(synthetic.to_torch_pairs([(t[0], t[1]) for t in triples])).pin_memory()
    tl = lambda key: [torch.from_numpy(t[2][key]) for t in triples]
    host_tgt = {k: [a.pin_memory() for a in tl(k)] for k in ('bound_lig', 'bound_rec', 'pocket_lig', 'pocket_rec')}
    dev_batch = host_batch.to(dev)
    dev_tgt = PocketBatch(host_tgt['bound_lig'], host_tgt['bound_rec'], host_tgt['pocket_lig'], host_tgt['pocket_rec'], dev)
    K, W, R = args.steps, max(args.warmup, 3), args.reps
    losses = []

    def body():
        last = None
        for _ in range(K):
            last = trainer.step(dev_batch, dev_tgt)
        losses.append(float(last['loss'][0].item()))          # D2H of the loss: the step's host-visible result
        check_loss_status(last)

    for _ in range(W):
        trainer.step(dev_batch, dev_tgt)
    sampler = bench.ClockSampler(local_rank)
    sampler.start()
    D.barrier()
    t_lead = time.perf_counter()
    while time.perf_counter() - t_lead < 1.0:
        trainer.step(dev_batch, dev_tgt)
    sampler.mark()
    rep_ms, med, per_rank_ms = bench.timed_reps(torch, D, R, body)
    ms_total = rep_ms[med]
    value = total_pairs * K / (ms_total * 1e-3)

    # e2e: every step also copies the batch (graph + labels) from pinned host memory and reads the loss back
    h2d = host_batch.nbytes() + sum(int(a.numel() * 4) for v in host_tgt.values() for a in v)

    def e2e_body():
        last = None
        for _ in range(K):
            g = host_batch.to(dev, non_blocking=True)
            tgt = PocketBatch(host_tgt['bound_lig'], host_tgt['bound_rec'], host_tgt['pocket_lig'], host_tgt['pocket_rec'], dev)
            last = trainer.step(g, tgt)
            float(last['loss'][0].item())
    e2e_body()
    e2e_rep = []
    for _ in range(R):
        D.barrier()
        t0 = time.perf_counter()
        e2e_body()
        D.barrier()
        e2e_rep.append(D.max(time.perf_counter() - t0))
    clocks = sampler.stop()
    e2e_val = total_pairs * K / float(np.median(e2e_rep))
    rank_clocks = D.gather(clocks.get('sm_mhz') or 0.0)
    if rank != 0:
        return D
    n_layers = wl['n_layers']
    launches_fwd = 2 + 7 + sum(1 + 2 + (0 if li == n_layers - 1 else 1) for li in range(n_layers))
    launches_bwd = 11 + n_layers * (1 + 1 + 2 + 1 + 1 + 1 + 2 * 11) + 1     # head (9 + TN + reduce), per layer, embed
    line = {
        'metric': bench.METRIC['train'], 'value': value, 'unit': 'pairs/s', 'n_gpus': world, 'steps': K, 'warmup': W,
        'ms_per_step': ms_total / K, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32',
        'data': 'synthetic', 'config': bench.workload_config(args, world),
        'notes': {'step': 'forward(stash) + device losses (MSE, exact EMD, intersection) + CUDA backward + bucketed NCCL '
                          'all-reduce of the flat gradient overlapped with backward + fused clip/Adam; weights are updated every step',
                  'coords_and_head_dtype': 'f64', 'numa': numa, 'shard': [lo, hi], 'pairs_total': total_pairs,
                  'nodes_per_rank': host_batch.num_nodes(), 'edges_per_rank': host_batch.num_edges(),
                  'loss_first_last': [losses[0], losses[-1]] if losses else None,
                  'value_protocol': f'median of {R} repetitions of the {K}-step loop, barrier+sync on both sides, CUDA events, max over ranks'},
        'rep_ms': rep_ms, 'per_rank_ms_per_step': [m / K for m in per_rank_ms],
        'e2e': {'value': e2e_val, 'unit': 'pairs/s', 'h2d_bytes_per_step': h2d, 'd2h_bytes_per_step': 8, 'rep_s': e2e_rep},
        'gpu_launches': (launches_fwd + launches_bwd + 5) * K * R,
        'clocks': {**clocks, 'per_rank_sm_mhz': rank_clocks},
        'roofline': None,
    }
    if world == 1 and not args.no_cpu_baseline:
        line['cpu_baseline'] = cpu_baseline(triples, args.cpu_seconds, bench)
    print(json.dumps(line), flush=True)
    return D


def cpu_baseline(triples, budget_s, bench):
    cores = bench.effective_cores()
    pool = bench.ReferencePool(cores, 'train')
    pool.run(triples[:pool.workers])
    n = max(pool.workers, min(len(triples), 2 * pool.workers))
    t0 = time.perf_counter()
    done = pool.run(triples[:n])
    dt = time.perf_counter() - t0
    pool.close()
    return {'value': done / dt, 'unit': 'pairs/s', 'cores': cores, 'kind': 'port',
            'sample': f'first {done} pairs of the step batch: torch fp32 port forward + losses + autograd backward per pair '
                      f'(EMD by HiGHS LP: POT is not in this image), {pool.workers} processes x {pool.threads} threads'}
