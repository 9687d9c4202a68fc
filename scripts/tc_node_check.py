"""GPU: each tensor-core node-stage kernel vs a torch fp64 evaluation of the same formula."""
import os, sys, ctypes as C
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import numpy as np, torch
import golden_io as gio
from equidock_public_b200 import _native as nat, hetero_graph as hg, synthetic
from equidock_public_b200.engine import GraphPlan
dev = torch.device('cuda:0')
lib = nat.load()
model = gio.build_model('dips', dev)
sd = {k: v.double() for k, v in model.state_dict().items()}
li = 2
pre = f'iegmn_original.iegmn_layers.{li}.'
lay = model.iegmn_original.iegmn_layers[li].packed(dev)
names, pairs, _, _ = gio.load_pairs('dips')
for desc, prs in (('golden dips x4', [pairs[n] for n in names]), ('synthetic 6x(200+200)+ragged', synthetic.synthetic_batch(6) + synthetic.synthetic_batch(2, 37, 301, 10, seed=5))):
    g = gio.make_batch(prs, dev)
    plan = GraphPlan.from_graph(g, dev, 10)
    N = plan.N
    torch.manual_seed(1)
    h = torch.randn(N, 64, device=dev) * 0.7
    G = C.byref(plan.struct); L = C.byref(lay.struct)
    # ---- projection
    proj_ff = torch.zeros(N, 320, device=dev); proj_tc = torch.zeros(N, 320, device=dev)
    kvb = lib.eqd_kv_blocks_bytes(N)
    kv = torch.zeros(kvb, dtype=torch.uint8, device=dev)
    assert lib.eqd_project(G, L, nat.ptr(h), 64, nat.ptr(proj_ff), None) == 0
    assert lib.eqd_project_tc(G, L, nat.ptr(h), nat.ptr(proj_tc), None, None) == 0      # all 5 groups as fp32
    assert lib.eqd_project_tc(G, L, nat.ptr(h), nat.ptr(proj_tc), nat.ptr(kv), None) == 0  # + K/V blocks
    torch.cuda.synchronize()
    w1 = sd[pre + 'edge_mlp.0.weight']; b1 = sd[pre + 'edge_mlp.0.bias']
    lr = lambda t: torch.nn.functional.leaky_relu(t, 0.01)
    hd = h.double()
    ref = torch.cat([hd @ w1[:, :64].t(), hd @ w1[:, 64:128].t() + b1, lr(hd @ sd[pre + 'att_mlp_Q.0.weight'].t()),
                     lr(hd @ sd[pre + 'att_mlp_K.0.weight'].t()), hd @ sd[pre + 'att_mlp_V.0.weight'].t()], 1)
    print(desc, 'proj: ffma err %.2e  tc err %.2e' % ((proj_ff.double() - ref).abs().max().item(), (proj_tc.double() - ref).abs().max().item()))
    # kv blocks round trip
    ng = (N + 7) // 8 + 8
    blocks = kv.view(torch.bfloat16).view(2, 3, ng, 8, 8, 8).float().sum(1)      # [which][ng][d/8][n%8][d%8]
    rec = blocks.permute(0, 1, 3, 2, 4).reshape(2, ng * 8, 64)[:, :N]
    print('   kv blocks: K err %.2e V err %.2e' % ((rec[0] - proj_tc[:, 192:256]).abs().max().item(), (rec[1] - proj_tc[:, 256:320]).abs().max().item()))
    # ---- attention
    mu = torch.zeros(N, 64, device=dev)
    assert lib.eqd_attention_tc(G, nat.ptr(proj_tc), nat.ptr(kv), nat.ptr(mu), None) == 0
    torch.cuda.synchronize()
    seg = plan.seg_ptr_host; B = plan.n_pairs
    mu_ref = torch.zeros(N, 64, dtype=torch.float64, device=dev)
    P = proj_tc.double()
    for s in range(2 * B):
        p_ = s + B if s < B else s - B
        qs = P[seg[s]:seg[s + 1], 128:192]; ks = P[seg[p_]:seg[p_ + 1], 192:256]; vs = P[seg[p_]:seg[p_ + 1], 256:320]
        mu_ref[seg[s]:seg[s + 1]] = torch.softmax(qs @ ks.t(), 1) @ vs
    print('   attention: err %.2e (max|mu| %.2f, max logit %.1f)' % ((mu.double() - mu_ref).abs().max().item(), mu_ref.abs().max().item(), (P[:, 128:192].abs().max() * P[:, 192:256].abs().max() * 64).item()))
    # ---- node MLP
    aggr = torch.randn(N, 64, device=dev) * 0.3
    h0 = torch.zeros(N, 72, device=dev); h0[:, :69] = torch.randn(N, 69, device=dev)
    hout = torch.zeros(N, 64, device=dev)
    assert lib.eqd_node_mlp_tc(G, L, nat.ptr(h), nat.ptr(aggr), nat.ptr(mu), nat.ptr(h0), nat.ptr(hout), None) == 0
    torch.cuda.synchronize()
    inp = torch.cat([hd, aggr.double(), mu.double(), h0[:, :69].double()], 1)
    hid = lr(inp @ sd[pre + 'node_mlp.0.weight'].t() + sd[pre + 'node_mlp.0.bias'])
    hid = torch.nn.functional.layer_norm(hid, (64,), sd[pre + 'node_mlp.3.weight'], sd[pre + 'node_mlp.3.bias'])
    out = hid @ sd[pre + 'node_mlp.4.weight'].t() + sd[pre + 'node_mlp.4.bias']
    out = 0.75 * out + 0.25 * hd
    print('   node mlp: err %.2e (max|h| %.2f)' % ((hout.double() - out).abs().max().item(), out.abs().max().item()), flush=True)
