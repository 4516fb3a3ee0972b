"""Per-block error growth: ours vs fp64 oracle vs fp32 oracle (whole SGCN, chebconv / gcnconv)."""
import sys, copy, torch
sys.path.insert(0, '.')
from oracle import pyg_ref as O
from semigcn_b200 import meshgen
from semigcn_b200.networks import SingleScaleGCN
dev = 'cuda:0'
def rel(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return ((a - b).abs().max() / b.abs().max()).item()
prob = meshgen.synth_inpainting_problem(10, smooth_iters=10, n_dummy=4)
mesh = prob['mesh']
for conv in ('gcnconv', 'chebconv'):
    torch.manual_seed(314)
    ref = O.SingleScaleGCN(conv); ref64 = copy.deepcopy(ref).double()
    ours = SingleScaleGCN(dev, conv=conv); ours.load_state_dict(ref.state_dict()); ours = ours.to(dev)
    x = torch.cat([prob['z1'], torch.ones(mesh.num_vertices, 1)], 1)
    acts = {}
    for name, net, dt, d in (('ours', ours, torch.float32, dev), ('r32', ref, torch.float32, 'cpu'), ('r64', ref64, torch.float64, 'cpu')):
        h = x.to(dt).to(d).detach().clone().requires_grad_(True)
        ei = mesh.edge_index.to(d)
        outs = [h]
        for b in net.blocks:
            h = b(h, ei); h.retain_grad(); outs.append(h)
        torch.manual_seed(1)
        w = torch.randn(h.shape, dtype=torch.float64).to(dt).to(d)
        (h * w).sum().backward()
        acts[name] = outs
    print(conv)
    for i in range(len(acts['ours'])):
        o, r32, r64 = acts['ours'][i], acts['r32'][i], acts['r64'][i]
        print(f"  block {i:2d} out: ours {rel(o, r64):.2e} r32 {rel(r32, r64):.2e} | grad: ours {rel(o.grad, r64.grad):.2e} r32 {rel(r32.grad, r64.grad):.2e}")
