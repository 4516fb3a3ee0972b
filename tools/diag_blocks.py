"""Gradient checks of the multi-resolution building blocks against the CPU oracle (fp32 and fp64)."""
import sys
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
import numpy as np, torch, torch.nn as nn
from helpers import load_golden, rel_err
from test_oracle_meshnet import _hier
from oracle import meshnet_ref as M, pyg_ref as O
from semigcn_b200 import ops
from semigcn_b200.meshnet import DownConv, UpConv
from semigcn_b200.nn import ChebConv, Sequential, MeshUnpool
DEV = torch.device('cuda:0')
gold = load_golden("ref_meshnet_n4.npz")
sizes, e, ph, uh, sm = _hier(gold)
g = torch.Generator().manual_seed(5)

def cmp(tag, net, ref, x, fwd, fwd_ref):
    xg = x.clone().to(DEV).requires_grad_(True)
    xr = x.clone().requires_grad_(True)
    y = fwd(net, xg); yr = fwd_ref(ref, xr)
    r = torch.randn(yr.shape, generator=torch.Generator().manual_seed(9))
    (y * r.to(DEV)).sum().backward(); (yr * r).sum().backward()
    print(f'== {tag}: fwd {rel_err(y, yr):.2e}  dx {rel_err(xg.grad, xr.grad):.2e}')
    for (k, p), (_, q) in zip(net.named_parameters(), ref.named_parameters()):
        if q.grad is not None:
            print(f'     {k:40s} {rel_err(p.grad, q.grad):.2e}   |g| {q.grad.abs().max().item():.2e}')

# (a) standalone BN + LeakyReLU through Sequential
for c, m in ((16, 642), (32, 385), (128, 231)):
    torch.manual_seed(1)
    ref = nn.Sequential(nn.BatchNorm1d(c), nn.LeakyReLU())
    net = Sequential("x, edge_index", [(nn.BatchNorm1d(c), "x -> x"), (nn.LeakyReLU(), "x -> x")]).to(DEV)
    with torch.no_grad():
        ref[0].weight.uniform_(0.5, 1.5); ref[0].bias.uniform_(-0.5, 0.5)
        net.module_0.weight.copy_(ref[0].weight); net.module_0.bias.copy_(ref[0].bias)
    x = torch.randn(m, c, generator=g) * 2 + 0.3
    cmp(f'BN+act c={c} m={m}', net, ref, x, lambda n, t: n(t, None), lambda n, t: n(t))

# (b) conv -> unpool -> BN -> act  (UpConv.model1)
for cin, cout in ((32, 16), (16, 16), (128, 32)):
    torch.manual_seed(2)
    ref = O.Sequential("x, edge_index", [(O.ChebConv(cin, cout, K=3), "x, edge_index -> x"), (M.MeshUnpool(uh[0]), "x -> x"),
                                           (nn.BatchNorm1d(cout), "x -> x"), (nn.LeakyReLU(), "x -> x")])
    net = Sequential("x, edge_index", [(ChebConv(cin, cout, K=3), "x, edge_index -> x"), (MeshUnpool(uh[0]), "x -> x"),
                                         (nn.BatchNorm1d(cout), "x -> x"), (nn.LeakyReLU(), "x -> x")])
    net.load_state_dict(ref.state_dict()); net = net.to(DEV)
    x = torch.randn(sizes[1], cin, generator=g)
    cmp(f'conv-unpool-BN-act {cin}->{cout}', net, ref, x, lambda n, t: n(t, e[1].to(DEV)), lambda n, t: n(t, e[1]))

# (c) conv (no BN) alone, and conv -> BN -> act fused, at 16 -> 16 on 642 vertices
for cin, cout in ((16, 16), (32, 16), (16, 3)):
    torch.manual_seed(3)
    ref = O.Sequential("x, edge_index", [(O.ChebConv(cin, cout, K=3), "x, edge_index -> x")])
    net = Sequential("x, edge_index", [(ChebConv(cin, cout, K=3), "x, edge_index -> x")])
    net.load_state_dict(ref.state_dict()); net = net.to(DEV)
    x = torch.randn(sizes[0], cin, generator=g)
    cmp(f'cheb alone {cin}->{cout}', net, ref, x, lambda n, t: n(t, e[0].to(DEV)), lambda n, t: n(t, e[0]))
    torch.manual_seed(3)
    ref = O.Sequential("x, edge_index", [(O.ChebConv(cin, cout, K=3), "x, edge_index -> x"), (nn.BatchNorm1d(cout), "x -> x"), (nn.LeakyReLU(), "x -> x")])
    net = Sequential("x, edge_index", [(ChebConv(cin, cout, K=3), "x, edge_index -> x"), (nn.BatchNorm1d(cout), "x -> x"), (nn.LeakyReLU(), "x -> x")])
    net.load_state_dict(ref.state_dict()); net = net.to(DEV)
    cmp(f'cheb-BN-act fused {cin}->{cout}', net, ref, x, lambda n, t: n(t, e[0].to(DEV)), lambda n, t: n(t, e[0]))

# (d) whole UpConv / DownConv
torch.manual_seed(4)
ref = M.UpConv(16, 8, e[1], e[0], uh[0], K=3, drop_rate=0.0)
net = UpConv(16, 8, e[1].to(DEV), e[0].to(DEV), uh[0], K=3, drop_rate=0.0)
net.load_state_dict(ref.state_dict()); net = net.to(DEV)
cmp('UpConv 16->8', net, ref, torch.randn(sizes[1], 16, generator=g), lambda n, t: n(t), lambda n, t: n(t))
torch.manual_seed(4)
ref = M.DownConv(4, 16, e[0], e[1], ph[0], K=3, drop_rate=0.0)
net = DownConv(4, 16, e[0].to(DEV), e[1].to(DEV), ph[0], K=3, drop_rate=0.0)
net.load_state_dict(ref.state_dict()); net = net.to(DEV)
cmp('DownConv 4->16', net, ref, torch.randn(sizes[0], 4, generator=g), lambda n, t: n(t), lambda n, t: n(t))
