"""Weight-gradient kernel, engine 3 (single CTA) against engine 4 (CTA pairs), CUDA-event timed on the bench's shapes."""
import sys
import torch
sys.path.insert(0, ".")
from semigcn_b200 import ops

dev = torch.device("cuda:0")
m = 998562
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
for n, k in [(256, 256), (512, 256), (256, 512), (256, 128), (256, 64)]:
    g = torch.randn(m, n, device=dev) * 1e-3
    a = torch.randn(m, k, device=dev)
    want = g.double().t() @ a.double()
    ga, aa = ops.amax(g), ops.amax(a)
    for engine in (3, 4):
        out = ops.gemm_tn(g, a, engine=engine, g_amax=ga, a_amax=aa)
        err = ((out.double() - want).norm() / want.norm()).item()
        ts = []
        for _ in range(8):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            ops.gemm_tn(g, a, engine=engine, g_amax=ga, a_amax=aa)
            e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        ms = sorted(ts)[len(ts) // 2]
        gb = 4.0 * (m * (n + k) + n * k) / 1e9
        print(f"gemm_tn n={n} k={k} engine={engine}: {ms:.3f} ms  {gb / ms * 1e3:.0f} GB/s  err {err:.2e}", flush=True)
