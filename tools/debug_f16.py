import sys, torch
sys.path.insert(0, '.')
from semigcn_b200 import ops
dev = 'cuda:0'
torch.manual_seed(0)
for (m, n, k) in [(128, 256, 160), (128, 256, 192), (128, 256, 224), (128, 256, 256), (128, 64, 256), (128, 256, 512)]:
    a = torch.randn(m, k, device=dev); w = torch.randn(n, k, device=dev)
    got = ops.gemm(a, w, engine=3)
    want = (a.double() @ w.double().t())
    err = (got.double() - want).abs()
    rel = err.max().item() / want.abs().max().item()
    print(f'm={m} n={n} k={k}: rel {rel:.2e}')
    if rel > 1e-4:
        bad = []
        for q in range((k + 31) // 32):
            a2 = torch.zeros_like(a); a2[:, q * 32:(q + 1) * 32] = a[:, q * 32:(q + 1) * 32]
            g2 = ops.gemm(a2, w, engine=3); w2 = a2.double() @ w.double().t()
            e = ((g2.double() - w2).abs().amax(dim=1) / w2.abs().max()).cpu()
            if e.max() > 1e-4:
                rows = (e > 1e-4).nonzero().flatten().tolist()
                bad.append((q, len(rows), rows[:4], rows[-4:]))
        print('   bad chunks:', bad)
