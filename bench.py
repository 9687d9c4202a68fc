#!/usr/bin/env python
"""bench.py -- protein pairs/sec of the IEGMN forward hot path (IEGMN layers + keypoints + Kabsch).

    python bench.py --gpus N --steps K --warmup W            # B200 engine (this repo)
    python bench.py --impl reference --gpus N ...            # CPU reference arm (oracle port, rank 0)

Workload (BASELINE.json north_star / configs[1] shape): synthetic DB5.5-shaped residue graphs,
200+200 residues, k=10, 8-layer IEGMN with the shipped DIPS checkpoint's weights, batched inference,
`--pairs-per-gpu` pairs per step per GPU (weak scaling; pairs shard with no data-path collective).
One step = one forward of the whole batch.  Prints ONE JSON line on rank 0.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))

N_LIG = N_REC = 200
KNN = 10
N_LAYERS = 8
# algorithmic work per 200+200 pair, 8 layers, reference formulation
FLOP_PER_PAIR = 1.781e9
BYTES_PER_PAIR = 6.20e6


def edge_stage_algorithmic_bytes(n_nodes: int, n_edges: int) -> float:
    """Compulsory HBM bytes of ONE edge-stage launch (one layer): he (27 fp32 / edge), CSR ids
    (4(E+N+2)), coordinates in + out (12 B / node each) -- the SURVEY 8(d) per-layer terms that
    flow through this kernel (DESIGN.md 'kernels')."""
    return 4.0 * 27 * n_edges + 4.0 * (n_edges + n_nodes + 2) + 12.0 * 2 * n_nodes


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ('index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,'
         'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,'
         'clocks_event_reasons.sw_power_cap')

    def __init__(self, gpu_index: int):
        self.gpu, self.rows, self.proc = gpu_index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(['nvidia-smi', f'--query-gpu={self.Q}', '--format=csv,noheader,nounits',
                                          '-lms', '100', '-i', str(self.gpu)], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except OSError:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable']}
        time.sleep(0.15)
        self.proc.terminate()
        sm, smax, reasons = [], [], set()
        for r in self.rows:
            f = [c.strip() for c in r.split(',')]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); smax.append(float(f[2]))
            except ValueError:
                continue
            for name, val in zip(('hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap'), f[5:9]):
                if val.lower().startswith('active'):
                    reasons.add(name)
        return {'sm_mhz': float(np.median(sm)) if sm else None, 'sm_max_mhz': max(smax) if smax else None,
                'reasons': sorted(reasons), 'samples': len(sm)}


class StageTimer:
    """CUDA events around each edge/node stage launch, on the launching (current) stream."""

    def __init__(self, torch):
        self.torch, self.ev, self.open = torch, {}, {}

    def begin(self, name, li):
        e = self.torch.cuda.Event(enable_timing=True)
        e.record()
        self.open[(name, li)] = e

    def end(self, name, li):
        e = self.torch.cuda.Event(enable_timing=True)
        e.record()
        self.ev.setdefault(name, []).append((self.open.pop((name, li)), e))

    def mean_ms(self, name):
        v = [a.elapsed_time(b) for a, b in self.ev.get(name, [])]
        return float(np.mean(v)) if v else None

    def total_ms(self, name):
        return float(sum(a.elapsed_time(b) for a, b in self.ev.get(name, [])))


def workload_config(pairs_per_gpu: int, world: int):
    return {'workload': f'synthetic DB5.5-shaped {N_LIG}+{N_REC} residues k={KNN}, {N_LAYERS}-layer IEGMN '
                        f'(DIPS checkpoint weights), batched inference, {pairs_per_gpu} pairs/step/GPU',
            'pairs_per_gpu': pairs_per_gpu, 'parallelism': f'dp{world} (pairs sharded, no data-path collective)'}


def make_workload(pairs_per_gpu: int, seed: int):
    from equidock_public_b200 import synthetic
    import golden_io as gio
    pairs = synthetic.synthetic_batch(pairs_per_gpu, N_LIG, N_REC, KNN, seed=seed)
    return pairs, gio.load_checkpoint('dips'), gio.load_args('dips')


def measured_peaks():
    """(HBM GB/s, dense bf16 TFLOP/s sustained, source).  The edge stage is timed inside a long step, so the sustained
    tensor figure is the denominator."""
    p = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.isfile(p):
        with open(p) as fh:
            d = json.load(fh)
        return (float(d['hbm_gbs']), float(d.get('bf16_tflops_sustained', d.get('bf16_tflops', 1380.0))),
                'measured (MEASURED_PEAKS.json)')
    return 6650.0, 1380.0, 'fallback (B200_PROFILING.md)'


def edge_stage_algorithmic_flops(n_edges: int, dh: int = 64) -> float:
    """fp32 FLOPs of ONE edge-stage launch in the reference formulation (SURVEY 8(d) per-layer edge terms, MAC = 2):
    edge_mlp.0 on cat[h_src, h_dst, he, rbf] (2E(2 dh + 42) 64), edge_mlp.4 and coors_mlp.0 (2E 64 64 each),
    coors_mlp.4 (2E 64)."""
    return n_edges * (2.0 * (2 * dh + 42) * 64 + 2 * 2.0 * 64 * 64 + 2.0 * 64)


# bf16 FLOPs the tensor-core edge stage really issues per edge: (K 48 x N 64 + K 64 x N 128) MACs x 6 split products
EDGE_TC_BF16_FLOP_PER_EDGE = 2.0 * (48 * 64 + 64 * 128) * 6
# dram__bytes_read.sum + dram__bytes_write.sum of one edge_stage_tc_kernel launch of this workload (ncu --set full,
# profiles/r01_v3_edge_stage_tc_ncu_summary.txt): 176.5 MB + 18.0 MB
EDGE_TC_NCU_TRAFFIC_BYTES = 194.5e6


def effective_cores() -> int:
    """Host cores this process may really use: min(affinity, cgroup CPU quota)."""
    n = len(os.sched_getaffinity(0)) if hasattr(os, 'sched_getaffinity') else (os.cpu_count() or 1)
    try:
        quota, period = open('/sys/fs/cgroup/cpu.max').read().split()
        if quota != 'max':
            n = min(n, max(1, int(int(quota) / int(period))))
    except (OSError, ValueError):
        pass
    return max(1, n)


_REF = {}


def _ref_init(threads):
    import torch
    sys.path.insert(0, os.path.join(ROOT, 'oracle'))
    import golden_io as gio
    import iegmn_oracle_torch as ot
    torch.set_num_threads(threads)
    sd, margs = gio.load_checkpoint('dips'), gio.load_args('dips')
    _REF['model'] = ot.TorchOracle(sd, N_LAYERS, margs['skip_weight_h'], margs['x_connection_init'],
                                   margs['leakyrelu_neg_slope'], margs['num_att_heads'])


def _ref_run(pairs):
    for p in pairs:
        _REF['model'].forward_pair(*p)
    return len(pairs)


class ReferencePool:
    """The CPU reference arm on ALL usable host cores: `workers` processes x `threads` torch threads, each
    running the oracle's torch port one pair per call (pairs are independent, like the GPU shards)."""

    def __init__(self, cores: int, threads: int = 4):
        import multiprocessing as mp
        self.threads = min(threads, cores)
        self.workers = max(1, cores // self.threads)
        self.pool = mp.get_context('spawn').Pool(self.workers, initializer=_ref_init, initargs=(self.threads,))

    def run(self, pairs):
        chunks = [pairs[i::self.workers] for i in range(self.workers)]
        return sum(self.pool.map(_ref_run, [c for c in chunks if c]))

    def close(self):
        self.pool.close()
        self.pool.join()


def run_reference(args, rank, world):
    """CPU reference arm: the oracle's PyTorch port (the reference's own op sequence, fp32, all host
    threads), one pair per call like src/inference_rigid.py, on a bounded sample of the workload."""
    if rank != 0:
        return
    cores = effective_cores()
    sample = min(args.ref_sample, args.pairs_per_gpu)
    pairs, sd, margs = make_workload(sample, seed=0)
    pool = ReferencePool(cores)
    for _ in range(max(1, args.warmup)):
        pool.run(pairs[:pool.workers])
    t0 = time.perf_counter()
    for _ in range(args.steps):
        pool.run(pairs)
    dt = time.perf_counter() - t0
    pool.close()
    val = args.steps * sample / dt
    line = {'impl': 'reference', 'metric': 'protein_pairs_per_sec_iegmn_fwd_kabsch', 'value': val, 'unit': 'pairs/s',
            'n_gpus': args.gpus, 'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': dt / args.steps * 1e3,
            'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
            'config': workload_config(args.pairs_per_gpu, world),
            'cpu_baseline': {'value': val, 'unit': 'pairs/s', 'cores': cores, 'kind': 'port',
                             'sample': f'each step = the first {sample} pairs of the {args.pairs_per_gpu}-pair batch, one pair '
                                       f'per call like src/inference_rigid.py; torch fp32 port of the reference op '
                                       f'sequence, {pool.workers} processes x {pool.threads} threads = {cores} usable host cores'},
            'e2e': {'value': val, 'unit': 'pairs/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
            'gpu_launches': 0}
    print(json.dumps(line), flush=True)


def run_engine(args, rank, local_rank, world):
    import torch
    import torch.distributed as dist
    import golden_io as gio
    from equidock_public_b200 import hetero_graph as hg
    from equidock_public_b200 import synthetic
    from equidock_public_b200.engine import IEGMNEngine

    torch.cuda.set_device(local_rank)
    dev = torch.device('cuda', local_rank)
    if world > 1:
        dist.init_process_group('nccl', device_id=dev)
    B = args.pairs_per_gpu
    pairs, sd, margs = make_workload(B, seed=rank)   # every rank owns its own shard of pairs
    model = gio.build_model('dips', dev, sd=sd, args=margs)
    host_batch = hg.batch_pairs(synthetic.to_torch_pairs(pairs)).pin_memory()
    dev_batch = host_batch.to(dev)
    n_nodes, n_edges = B * (N_LIG + N_REC), host_batch.num_edges()
    iegmn = model.iegmn_original

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def maxr(v):
        if world == 1:
            return v
        t = torch.tensor([v], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---- device-resident throughput ("value") -------------------------------------------------------
    for _ in range(max(args.warmup, 3)):
        model(dev_batch, 0)
    from equidock_public_b200 import engine as engine_mod
    timer = StageTimer(torch) if engine_mod._PY_FORWARD else engine_mod.NativeStageTimer()
    orig_forward = IEGMNEngine.forward
    IEGMNEngine.forward = lambda self, *a, **k: orig_forward(self, *a, stage_timer=timer, **k)
    sampler = ClockSampler(local_rank)
    barrier()
    if rank == 0:
        sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    pending = None
    for _ in range(args.steps):          # two steps in flight: step k is launched before step k-1's status words are read
        nxt = model.forward_async(dev_batch, 0)
        if pending is not None:
            pending.result()
        pending = nxt
    pending.result()
    e1.record()
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    ms_total = maxr(e0.elapsed_time(e1))
    IEGMNEngine.forward = orig_forward
    value = world * B * args.steps / (ms_total * 1e-3)

    # ---- end to end through the public API with HOST buffers ("e2e") -------------------------------------------------
    # every step: H2D of that step's pinned inputs, the module's forward, D2H of coordinates / R / t into pinned memory;
    # equidock_public_b200.serving.PipelinedInference overlaps the copy of batch k+1 with the kernels of batch k.
    from equidock_public_b200.serving import PipelinedInference
    pipe = PipelinedInference(model, dev)
    d2h_bytes = 0

    def drain(n_steps):
        nonlocal d2h_bytes
        last = None
        for res in pipe.run(host_batch for _ in range(n_steps)):
            last = res
        last['_event'].synchronize()
        d2h_bytes = sum(int(v.numel() * v.element_size()) for k, v in last.items() if k != '_event')

    drain(3)
    barrier()
    t0 = time.perf_counter()
    drain(args.steps)
    barrier()
    e2e_s = maxr(time.perf_counter() - t0)
    e2e_val = world * B * args.steps / e2e_s
    h2d_bytes = host_batch.nbytes()

    if rank != 0:
        return
    hbm_peak, tc_peak, peak_src = measured_peaks()
    edge_ms = timer.mean_ms('edge_stage')
    node_ms = timer.mean_ms('node_stage')
    alg = edge_stage_algorithmic_bytes(n_nodes, n_edges)
    ach = alg / (edge_ms * 1e-3) / 1e9
    alg_flops = edge_stage_algorithmic_flops(n_edges)
    step_ms = ms_total / args.steps
    sm_mhz = (clocks or {}).get('sm_mhz') or 1965.0
    fp32_peak = 148 * 128 * 2 * sm_mhz * 1e6 / 1e12
    line = {
        'metric': 'protein_pairs_per_sec_iegmn_fwd_kabsch', 'value': value, 'unit': 'pairs/s', 'n_gpus': world,
        'steps': args.steps, 'warmup': max(args.warmup, 3), 'ms_per_step': step_ms, 'higher_is_better': True,
        'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
        'config': {**workload_config(B, world),
                   'l2': f'per-step working set {(n_edges * 108 + n_nodes * 3880) / 1e6:.0f} MB > 126 MB L2, no flush needed',
                   'coords_and_head_dtype': 'f64'},
        'e2e': {'value': e2e_val, 'unit': 'pairs/s', 'h2d_bytes_per_step': h2d_bytes, 'd2h_bytes_per_step': d2h_bytes},
        'gpu_launches': IEGMNEngine.launches_per_forward(N_LAYERS) * args.steps,
        'clocks': clocks,
        'roofline': {'kernel': 'edge_stage_tc_kernel', 'bound': 'tensor', 'achieved': alg_flops / (edge_ms * 1e-3) / 1e12,
                     'peak': tc_peak, 'unit': 'TFLOP/s', 'frac': alg_flops / (edge_ms * 1e-3) / 1e12 / tc_peak,
                     'traffic': EDGE_TC_NCU_TRAFFIC_BYTES if B == 256 else None, 'peak_source': peak_src + ', sustained bf16',
                     'algorithmic_flops_per_launch': alg_flops, 'launch_ms': edge_ms,
                     'issued_bf16_tflops': n_edges * EDGE_TC_BF16_FLOP_PER_EDGE / (edge_ms * 1e-3) / 1e12,
                     'issued_bf16_frac': n_edges * EDGE_TC_BF16_FLOP_PER_EDGE / (edge_ms * 1e-3) / 1e12 / tc_peak,
                     'hbm': {'achieved': ach, 'peak': hbm_peak, 'unit': 'GB/s', 'frac': ach / hbm_peak,
                             'algorithmic_bytes_per_launch': alg},
                     'share_of_step': timer.total_ms('edge_stage') / ms_total,
                     'note': 'fp32-accurate GEMMs as 6 bf16 split products on tcgen05 (bf16x6): the tensor ceiling in '
                             'algorithmic fp32 FLOPs is peak x 38272 / 135168 = 0.283 x peak; AI ~290 FLOP/B, so the HBM '
                             'fraction (north star) is small by construction'},
        'kernels_ms': {'edge_stage': edge_ms, 'node_stage': node_ms,
                       'edge_share': timer.total_ms('edge_stage') / ms_total,
                       'node_share': timer.total_ms('node_stage') / ms_total},
        'step_roofline': {'hbm_frac': value / world * BYTES_PER_PAIR / 1e9 / hbm_peak,
                          'fp32_tflops': value / world * FLOP_PER_PAIR / 1e12, 'fp32_peak_tflops': fp32_peak,
                          'fp32_frac': value / world * FLOP_PER_PAIR / 1e12 / fp32_peak,
                          'algorithmic': 'SURVEY 8(d): 1.781 GFLOP, 6.20 MB per pair (reference formulation)'},
    }
    if world == 1 and not args.no_cpu_baseline:
        line['cpu_baseline'] = cpu_baseline(pairs, args.cpu_seconds)
    print(json.dumps(line), flush=True)
    if world > 1:
        pass


def cpu_baseline(pairs, budget_s):
    """Oracle port timed on this box's usable host cores on a bounded sample of the same workload."""
    cores = effective_cores()
    pool = ReferencePool(cores)
    pool.run(pairs[:pool.workers])                       # warm-up: imports, weights, first-call allocations
    t0 = time.perf_counter()
    pool.run(pairs[:2 * pool.workers])
    rate = 2 * pool.workers / (time.perf_counter() - t0)
    n = int(min(len(pairs), max(2 * pool.workers, rate * budget_s)))
    t0 = time.perf_counter()
    done = pool.run(pairs[:n])
    dt = time.perf_counter() - t0
    pool.close()
    return {'value': done / dt, 'unit': 'pairs/s', 'cores': cores, 'kind': 'port',
            'sample': f'first {done} pairs of the step batch, one pair per call, torch fp32 port of the reference op '
                      f'sequence, {pool.workers} processes x {pool.threads} threads'}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=10)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='b200', choices=['b200', 'reference'])
    ap.add_argument('--pairs-per-gpu', type=int, default=256)
    ap.add_argument('--ref-sample', type=int, default=128, help='pairs per step of the CPU reference arm')
    ap.add_argument('--cpu-seconds', type=float, default=15.0)
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--watchdog-seconds', type=int, default=1500,
                    help='abort (with a stack dump) instead of stalling forever if the run has not finished by then')
    args = ap.parse_args()
    if args.watchdog_seconds > 0:
        import faulthandler
        faulthandler.dump_traceback_later(args.watchdog_seconds, exit=True)
    rank = int(os.environ.get('RANK', '0'))
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    if args.impl == 'reference':
        run_reference(args, rank, world)
        return
    run_engine(args, rank, local_rank, world)
    if world > 1:
        import torch.distributed as dist
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
