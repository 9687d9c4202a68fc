"""GPU: the layer-0 (69-wide) tensor-core kernels vs a torch fp64 evaluation of the same formulas."""
import os, sys, ctypes as C
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import numpy as np, torch
import golden_io as gio
from equidock_public_b200 import _native as nat, hetero_graph as hg, synthetic
from equidock_public_b200.engine import GraphPlan
dev = torch.device('cuda:0')
lib = nat.load()
for which in ('dips', 'db5'):
    model = gio.build_model(which, dev)
    sd = {k: v.double() for k, v in model.state_dict().items()}
    pre = 'iegmn_original.iegmn_layers.0.'
    lay = model.iegmn_original.iegmn_layers[0].packed(dev)
    slope = 0.01
    names, pairs, _, _ = gio.load_pairs(which)
    for desc, prs in ((f'golden {which}', [pairs[n] for n in names]), ('synthetic 6x(200+200)+ragged', synthetic.synthetic_batch(6) + synthetic.synthetic_batch(2, 37, 301, 10, seed=5))):
        g = gio.make_batch(prs, dev)
        plan = GraphPlan.from_graph(g, dev, 10)
        N = plan.N
        torch.manual_seed(1)
        h0 = torch.zeros(N, 72, device=dev); h0[:, :69] = torch.randn(N, 69, device=dev) * 0.7
        G = C.byref(plan.struct); L = C.byref(lay.struct)
        proj = torch.zeros(N, 344, device=dev)
        kv = torch.zeros(lib.eqd_kv_blocks_bytes(N), dtype=torch.uint8, device=dev)
        nrow = ((N + 7) // 8 + 8) * 8
        x5 = torch.zeros(nrow, 16, device=dev)
        rc = lib.eqd_project_tc0(G, L, nat.ptr(h0), nat.ptr(proj), nat.ptr(kv), nat.ptr(x5), None)
        torch.cuda.synchronize(); assert rc == 0, rc
        w1 = sd[pre + 'edge_mlp.0.weight']; b1 = sd[pre + 'edge_mlp.0.bias']
        lr = lambda t: torch.nn.functional.leaky_relu(t, slope)
        hd = h0[:, :69].double()
        Q = lr(hd @ sd[pre + 'att_mlp_Q.0.weight'].t()); K = lr(hd @ sd[pre + 'att_mlp_K.0.weight'].t()); V = hd @ sd[pre + 'att_mlp_V.0.weight'].t()
        ps, pd = hd @ w1[:, :69].t(), hd @ w1[:, 69:138].t() + b1
        e = lambda a, b: (a.double() - b).abs().max().item()
        ng = (N + 7) // 8 + 8
        blocks = kv.view(torch.bfloat16).view(2, 3, ng, 8, 8, 8).float().sum(1)
        rec = blocks.permute(0, 1, 3, 2, 4).reshape(2, ng * 8, 64)[:, :N]
        x = x5[:N]
        print(f'{desc}: proj0 Psrc {e(proj[:, :64], ps):.2e} Pdst {e(proj[:, 64:128], pd):.2e} Q64 {e(proj[:, 128:192], Q[:, :64]):.2e} '
              f'K64 {e(rec[0], K[:, :64]):.2e} V64 {e(rec[1], V[:, :64]):.2e} K5 {max(e(x[:, 0:4], K[:, 64:68]), e(x[:, 8], K[:, 68])):.2e} '
              f'V5 {max(e(x[:, 4:8], V[:, 64:68]), e(x[:, 9], V[:, 68])):.2e} Q5 {e(x[:, 10:15], Q[:, 64:69]):.2e} pad {x5[N:].abs().max().item():.1e}', flush=True)
        # attention
        mu = torch.full((N, 72), 7.0, device=dev)
        rc = lib.eqd_attention_tc0(G, nat.ptr(proj), nat.ptr(kv), nat.ptr(x5), nat.ptr(mu), None)
        torch.cuda.synchronize(); assert rc == 0, rc
        seg = plan.seg_ptr_host; B = plan.n_pairs
        mu_ref = torch.zeros(N, 69, dtype=torch.float64, device=dev)
        for s in range(2 * B):
            p_ = s + B if s < B else s - B
            mu_ref[seg[s]:seg[s + 1]] = torch.softmax(Q[seg[s]:seg[s + 1]] @ K[seg[p_]:seg[p_ + 1]].t(), 1) @ V[seg[p_]:seg[p_ + 1]]
        print(f'   attention0: err {e(mu[:, :69], mu_ref):.2e} (max|mu| {mu_ref.abs().max().item():.2f}) pad {mu[:, 69:].abs().max().item():.1e}', flush=True)
        # node MLP
        aggr = torch.randn(N, 64, device=dev) * 0.3
        hout = torch.zeros(N, 64, device=dev)
        rc = lib.eqd_node_mlp_tc0(G, L, nat.ptr(h0), nat.ptr(aggr), nat.ptr(mu), nat.ptr(hout), None)
        torch.cuda.synchronize(); assert rc == 0, rc
        inp = torch.cat([hd, aggr.double(), mu[:, :69].double(), hd], 1)
        hid = lr(inp @ sd[pre + 'node_mlp.0.weight'].t() + sd[pre + 'node_mlp.0.bias'])
        hid = torch.nn.functional.layer_norm(hid, (69,), sd[pre + 'node_mlp.3.weight'], sd[pre + 'node_mlp.3.bias'])
        out = hid @ sd[pre + 'node_mlp.4.weight'].t() + sd[pre + 'node_mlp.4.bias']
        print(f'   node mlp0: err {e(hout, out):.2e} (max|h| {out.abs().max().item():.2f})', flush=True)
print('done')
