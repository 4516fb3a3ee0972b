import sys, torch
sys.path.insert(0, '.')
from semigcn_b200 import ops
dev = 'cuda:0'
torch.manual_seed(0)
m, n, k = 128, 256, 256
a = torch.randn(m, k, device=dev); w = torch.randn(n, k, device=dev)
amx = a.abs().max().reshape(1)
got = ops.gemm(a, w, engine=3, a_amax=amx)
torch.cuda.synchronize()
want = (a.double() @ w.double().t())
print('rel', ((got.double() - want).abs().max() / want.abs().max()).item())
