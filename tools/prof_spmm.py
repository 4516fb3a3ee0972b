"""One SpMM launch per width on the 1M-vertex icosphere (for ncu):  python tools/prof_spmm.py 128,256 [mode]"""
import sys, torch
sys.path.insert(0, '.')
from semigcn_b200 import ops, meshgen
dev = 'cuda:0'
widths = [int(c) for c in (sys.argv[1] if len(sys.argv) > 1 else '128,256').split(',')]
mode = int(sys.argv[2]) if len(sys.argv) > 2 else 0
stats = len(sys.argv) > 3
mesh = meshgen.icosphere(316, device=dev)
g = ops.MeshGraph(mesh.edge_index, mesh.num_vertices, mode)
for c in widths:
    x = torch.randn(mesh.num_vertices, c, device=dev)
    for _ in range(3):
        ops.spmm(g, x, want_stats=stats)
torch.cuda.synchronize()
