import os, time, sys
print('cpu_count', os.cpu_count(), 'affinity', len(os.sched_getaffinity(0)))
for f in ('/sys/fs/cgroup/cpu.max', '/sys/fs/cgroup/cpu/cpu.cfs_quota_us', '/sys/fs/cgroup/cpu/cpu.cfs_period_us'):
    try: print(f, open(f).read().strip())
    except Exception as e: print(f, 'n/a')
sys.path.insert(0, '.'); sys.path.insert(0, 'tests'); sys.path.insert(0, 'oracle')
import torch, golden_io as gio, iegmn_oracle_torch as ot
from equidock_public_b200 import synthetic
sd, a = gio.load_checkpoint('dips'), gio.load_args('dips')
pairs = synthetic.synthetic_batch(4)
m = ot.TorchOracle(sd, 8, a['skip_weight_h'])
for t in (1, 2, 4, 8, 16, 32):
    torch.set_num_threads(t)
    m.forward_pair(*pairs[0])
    t0 = time.perf_counter()
    for p in pairs: m.forward_pair(*p)
    print('threads', t, 'pairs/s', 4 / (time.perf_counter() - t0), flush=True)
