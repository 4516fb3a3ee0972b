"""Per-kernel-shape CUDA-event breakdown of one SGCN train step on the 1M-vertex mesh."""
import sys, torch
sys.path.insert(0, '.')
import bench
from semigcn_b200 import profile
from semigcn_b200.data import Data
from semigcn_b200.networks import SingleScaleGCN
dev = torch.device('cuda:0')
prob = bench.make_problem(316, dev)
mesh = prob['mesh']
torch.manual_seed(314)
net = SingleScaleGCN(dev, conv=sys.argv[1] if len(sys.argv) > 1 else 'gcnconv').to(dev)
opt = torch.optim.Adam(net.parameters(), lr=0.01)
data = Data(z1=prob['z1'], x_pos=prob['x_pos'], edge_index=mesh.edge_index)
def step(i):
    opt.zero_grad(set_to_none=True)
    out = net(data, prob['dms'][:, i % 8:i % 8 + 1])
    loss = bench.step_losses(out, prob)
    loss.backward(); opt.step()
for i in range(3): step(i)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for i in range(3): step(i)
e1.record(); torch.cuda.synchronize()
print('step ms', e0.elapsed_time(e1) / 3)
with profile.KernelProfile() as kp:
    for i in range(3): step(i)
    fam = kp.summary()
tot = 0
for k, v in sorted(fam.items(), key=lambda kv: -kv[1]['ms']):
    print(f"{k:24s} {v['ms']/3:7.3f} ms  x{v['launches']//3:2d}  {v['bytes']/max(v['ms'],1e-9)/1e6:8.1f} GB/s")
    tot += v['ms'] / 3
print('sum', tot)
