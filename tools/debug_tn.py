"""Diagnostic: decode how k_gemm_tn_tc (MN-major operands) maps inputs to outputs."""
import sys, torch
sys.path.insert(0, '.')
from semigcn_b200 import ops
dev = 'cuda:0'
torch.set_printoptions(linewidth=200, precision=4, sci_mode=False)
m, n, k = 32, 128, 16
# probe 1: all ones -> D should be m everywhere
g = torch.ones(m, n, device=dev); a = torch.ones(m, k, device=dev)
d = ops.gemm_tn(g, a, engine=2)
print('all-ones: min/max', d.min().item(), d.max().item(), 'expected', m)
# probe 2: one-hot in G
for (ms, ns) in [(0, 0), (0, 1), (0, 4), (1, 0), (8, 0), (3, 37), (31, 127)]:
    g = torch.zeros(m, n, device=dev); g[ms, ns] = 1.0
    a = (torch.arange(k, device=dev).float() + 1).repeat(m, 1) + 100 * torch.arange(m, device=dev).float().reshape(-1, 1)
    d = ops.gemm_tn(g, a, engine=2)
    nz = torch.nonzero(d.abs() > 1e-6)
    rows = sorted(set(nz[:, 0].tolist()))
    print(f'G one-hot at (m={ms}, n={ns}): nonzero rows {rows[:8]} ; expected row {ns} = {a[ms].tolist()[:6]}')
    for r in rows[:3]:
        print('   row', r, d[r].tolist())
# probe 3: random vs reference
g = torch.randn(m, n, device=dev); a = torch.randn(m, k, device=dev)
d = ops.gemm_tn(g, a, engine=2); want = g.double().t() @ a.double()
print('random: max err', (d.double() - want).abs().max().item(), 'max want', want.abs().max().item())
