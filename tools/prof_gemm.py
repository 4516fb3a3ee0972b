"""One launch per GEMM engine on a 1M-row operand (for ncu):  python tools/prof_gemm.py K N [engine] [tn]"""
import sys, torch
sys.path.insert(0, '.')
from semigcn_b200 import ops
dev = 'cuda:0'
k, n = int(sys.argv[1]), int(sys.argv[2])
eng = int(sys.argv[3]) if len(sys.argv) > 3 else 3
tn = len(sys.argv) > 4
m = 998562
a = torch.randn(m, k, device=dev)
if tn:
    g = torch.randn(m, n, device=dev)
    ga, aa = g.abs().max().reshape(1), a.abs().max().reshape(1)
    for _ in range(3):
        ops.gemm_tn(g, a, engine=eng, g_amax=ga if eng == 3 else None, a_amax=aa if eng == 3 else None)
else:
    w = torch.randn(n, k, device=dev)
    amx = a.abs().max().reshape(1)
    for _ in range(3):
        ops.gemm(a, w, engine=eng, a_amax=amx if eng == 3 else None)
torch.cuda.synchronize()
