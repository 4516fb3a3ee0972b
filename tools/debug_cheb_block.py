"""Debug: every transform of a ChebConv block's forward / backward against fp64, with the max|operand| scalars it was given."""
import sys
import torch
import torch.nn as nn
sys.path.insert(0, ".")
from semigcn_b200 import meshgen, ops
import semigcn_b200.nn as N

dev = "cuda:0"
mesh = meshgen.icosphere(32)
n = mesh.vs.shape[0]
cin, cout = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (256, 128)
torch.manual_seed(314)
blk = N.Sequential("x, edge_index", [(N.ChebConv(cin, cout, K=3), "x, edge_index -> x"), nn.BatchNorm1d(cout), nn.LeakyReLU()]).to(dev)
g = torch.Generator().manual_seed(7)
x, dz = torch.randn(n, cin, generator=g).to(dev), torch.randn(n, cout, generator=g).to(dev)
_gemm, _tn = ops.gemm, ops.gemm_tn


def rel(a, b):
    return ((a.double() - b).abs().max() / b.abs().max()).item()


def gemm(a, b, transb=True, **kw):
    c0 = kw["out"].double().clone() if kw.get("accumulate") else None
    r = _gemm(a, b, transb=transb, **kw)
    c = r[0] if isinstance(r, tuple) else r
    want = a.double() @ (b.double().t() if transb else b.double())
    if kw.get("bias") is not None:
        want = want + kw["bias"].double()
    if c0 is not None:
        want = want + c0
    am = kw.get("a_amax")
    print(f"gemm m={a.shape[0]} n={c.shape[1]} k={a.shape[1]} transb={transb} lda={a.stride(0)} a_amax={None if am is None else float(am)} true max|A|={float(a.abs().max()):.4g} err={rel(c, want):.2e}")
    return r


def gemm_tn(gm, a, **kw):
    d = _tn(gm, a, **kw)
    want = gm.double().t() @ a.double()
    ga, aa = kw.get("g_amax"), kw.get("a_amax")
    print(f"gemm_tn n={gm.shape[1]} k={a.shape[1]} ldg={gm.stride(0)} lda={a.stride(0)} g_amax={None if ga is None else float(ga)} max|G|={float(gm.abs().max()):.4g} "
          f"a_amax={None if aa is None else float(aa)} max|A|={float(a.abs().max()):.4g} err={rel(d, want):.2e}")
    return d


ops.gemm, ops.gemm_tn = gemm, gemm_tn
xg = x.clone().requires_grad_(True)
out = blk(xg, mesh.edge_index.to(dev))
out.backward(dz)
torch.cuda.synchronize()
