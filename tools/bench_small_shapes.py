"""Engine choice for the narrow layers of the SGCN stack (998 562 rows): CUDA-core tiles (engine 1) against the fp16-split tensor-core
tiles (engine 3; forward also on CTA pairs), max|operand| supplied as in the train step.  CUDA events, median of 8."""
import os
import sys
import torch
sys.path.insert(0, ".")
from semigcn_b200 import ops

dev = torch.device("cuda:0")
m = 998562
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def med(fn):
    fn()
    ts = []
    for _ in range(8):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return sorted(ts)[len(ts) // 2]


print("forward / dX  C[m,n] = A[m,k] W^T")
for n, k in [(32, 64), (64, 32), (32, 16), (16, 32), (16, 4), (64, 128), (128, 64), (16, 3), (3, 16)]:
    a = torch.randn(m, k, device=dev); w = torch.randn(n, k, device=dev); amx = a.abs().max().reshape(1)
    c = torch.empty(m, n, device=dev)
    row = [f"n{n}_k{k}"]
    for label, engine, pair in (("simt", 1, None), ("f16", 3, "0"), ("f16pair", 3, "1")):
        if pair is not None:
            os.environ["SGB_F16_PAIR"] = pair
        try:
            ms = med(lambda: ops.gemm(a, w, engine=engine, a_amax=amx if engine == 3 else None, out=c))
            row.append(f"{label} {ms:.3f}")
        except Exception as e:  # noqa: BLE001
            row.append(f"{label} n/a")
    os.environ.pop("SGB_F16_PAIR", None)
    print("  " + " | ".join(row), flush=True)
print("weight gradient  D[n,k] = G[m,n]^T A[m,k]")
for n, k in [(32, 64), (64, 32), (32, 16), (16, 32), (16, 4), (3, 16), (64, 128), (128, 64), (128, 256)]:
    g = torch.randn(m, n, device=dev) * 1e-3; a = torch.randn(m, k, device=dev)
    ga, aa = g.abs().max().reshape(1), a.abs().max().reshape(1)
    row = [f"n{n}_k{k}"]
    for label, engine in (("simt", 1), ("f16", 3)):
        try:
            ms = med(lambda: ops.gemm_tn(g, a, engine=engine, g_amax=ga if engine == 3 else None, a_amax=aa if engine == 3 else None))
            row.append(f"{label} {ms:.3f}")
        except Exception as e:  # noqa: BLE001
            row.append(f"{label} n/a")
    print("  " + " | ".join(row), flush=True)
