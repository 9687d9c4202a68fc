import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import torch
print('threads', torch.get_num_threads(), flush=True)
import golden_io as gio
from equidock_public_b200 import hetero_graph as hg, synthetic
from equidock_public_b200.serving import PipelinedInference
dev = torch.device('cuda:0')
model = gio.build_model('dips', dev)
B = 256
host = hg.batch_pairs(synthetic.to_torch_pairs(synthetic.synthetic_batch(B))).pin_memory()
pipe = PipelinedInference(model, dev)
for rep in range(3):
    ts = []
    t0 = time.perf_counter()
    for r in pipe.run(host for _ in range(10)):
        ts.append(time.perf_counter() - t0)
    r['_event'].synchronize(); ts.append(time.perf_counter() - t0)
    print('rep', rep, ' '.join(f'{x*1e3:.1f}' for x in ts), flush=True)
import cProfile, pstats
pr = cProfile.Profile(); pr.enable()
for r in pipe.run(host for _ in range(10)): pass
r['_event'].synchronize()
pr.disable()
pstats.Stats(pr).sort_stats('tottime').print_stats(14)
