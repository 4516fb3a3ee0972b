"""Whole-network gradient noise: ours vs fp32 oracle, both against the fp64 oracle, several seeds."""
import sys, copy, torch
sys.path.insert(0, '.')
from oracle import pyg_ref as O
from semigcn_b200 import meshgen
from semigcn_b200.data import Data
from semigcn_b200.networks import SingleScaleGCN
dev = 'cuda:0'
def rel(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return ((a - b).abs().max() / b.abs().max()).item()
prob = meshgen.synth_inpainting_problem(10, smooth_iters=10, n_dummy=4)
mesh = prob['mesh']
for conv in ('chebconv', 'gcnconv'):
  for seed in (314, 1, 2):
    torch.manual_seed(seed)
    ref = O.SingleScaleGCN(conv); ref64 = copy.deepcopy(ref).double()
    ours = SingleScaleGCN(dev, conv=conv); ours.load_state_dict(ref.state_dict()); ours = ours.to(dev)
    dm = prob["vmask_dummy"][:, :1] * prob["v_mask"].float().reshape(-1, 1)
    def run_ref(net, dt):
        out = net(prob["z1"].to(dt), prob["x_pos"].to(dt), mesh.edge_index, dm.to(dt))
        O.mask_pos_rec_loss(out, prob["ini_vs"], prob["v_mask"]).backward()
    run_ref(ref, torch.float32); run_ref(ref64, torch.float64)
    out = ours(Data(z1=prob["z1"].to(dev), x_pos=prob["x_pos"].to(dev), edge_index=mesh.edge_index.to(dev)), dm)
    O.mask_pos_rec_loss(out, prob["ini_vs"].to(dev), prob["v_mask"].to(dev)).backward()
    r32, r64 = dict(ref.named_parameters()), dict(ref64.named_parameters())
    print(conv, 'seed', seed)
    for name, p in ours.named_parameters():
        if r32[name].grad is None or name.endswith('module_0.bias') or 'lins.1' in name or 'lins.2' in name or 'module_1.bias' in name: continue
        print(f"   {name:36s} ours {rel(p.grad, r64[name].grad):.2e}   r32 {rel(r32[name].grad, r64[name].grad):.2e}")
