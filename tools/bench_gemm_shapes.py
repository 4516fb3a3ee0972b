"""Micro-benchmark of the fp16-split transform at the bench's shapes (M = 998 562 rows, max|A| supplied), CUDA events.
   python tools/bench_gemm_shapes.py [label]      (SGB_LIB_PATH selects an A/B build of the library)"""
import sys, torch
sys.path.insert(0, '.')
from semigcn_b200 import ops
dev, m = 'cuda:0', 998562
label = sys.argv[1] if len(sys.argv) > 1 else 'default'
def timeit(fn, reps=6, warm=2):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
import os
out = []
for (k, n) in ((256, 256), (256, 512), (512, 256), (128, 256), (256, 128), (64, 128), (128, 64), (32, 64), (64, 32), (16, 32)):
    a = torch.randn(m, k, device=dev); w = torch.randn(n, k, device=dev); amx = a.abs().max().reshape(1)
    c = torch.empty(m, n, device=dev)
    ms = timeit(lambda: ops.gemm(a, w, engine=3, a_amax=amx, out=c))
    out.append(f"k{k}_n{n} {ms:.3f} ms {4.0 * m * (k + n) / ms / 1e6:.0f} GB/s")
print(f"{label:>14s}: " + " | ".join(out), flush=True)
