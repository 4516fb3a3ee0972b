"""SpMM throughput vs vertex ordering on the 1M-vertex icosphere: generator order, Morton order, random order."""
import sys, torch
sys.path.insert(0, '.')
from semigcn_b200 import ops, meshgen
dev = 'cuda:0'
mesh = meshgen.icosphere(316, device=dev)
n, nnz = mesh.num_vertices, mesh.nnz
def timeit(fn, reps=5, warm=2):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
def morton(vs, bits=10):
    q = ((vs - vs.min(0)[0]) / (vs.max(0)[0] - vs.min(0)[0]) * (2 ** bits - 1)).long()
    code = torch.zeros(vs.shape[0], dtype=torch.int64, device=vs.device)
    for b in range(bits):
        for d in range(3):
            code |= ((q[:, d] >> b) & 1) << (3 * b + d)
    return code
orders = {'generator': None, 'morton': torch.argsort(morton(mesh.vs.float())), 'random': torch.randperm(n, device=dev)}
for name, perm in orders.items():
    ei = mesh.edge_index
    if perm is not None:
        inv = torch.empty_like(perm); inv[perm] = torch.arange(n, device=dev)
        ei = inv[ei]
    g = ops.MeshGraph(ei.contiguous(), n, 0)
    for c in (64, 128, 256):
        x = torch.randn(n, c, device=dev)
        ms = timeit(lambda: ops.spmm(g, x))
        by = 4.0 * (2 * n * c + nnz + n + 2 * n + 1)
        print(f'{name:10s} c={c:4d}: {ms:7.3f} ms  {by / ms / 1e6:8.1f} GB/s')
