"""Kernel census of one SGCN train step (torch profiler, CUDA activities):  python tools/kernel_census.py [freq]"""
import sys, torch, collections
sys.path.insert(0, '.')
import bench
from semigcn_b200.data import Data
from semigcn_b200.networks import SingleScaleGCN
from torch.profiler import profile, ProfilerActivity
dev = torch.device('cuda:0')
freq = int(sys.argv[1]) if len(sys.argv) > 1 else 32
prob = bench.make_problem(freq, dev)
mesh = prob['mesh']
torch.manual_seed(314)
net = SingleScaleGCN(dev, conv='gcnconv').to(dev)
opt = torch.optim.Adam(net.parameters(), lr=0.01)
data = Data(z1=prob['z1'], x_pos=prob['x_pos'], edge_index=mesh.edge_index)
dm = prob['dms'][:, 0:1].contiguous()
def step():
    opt.zero_grad(set_to_none=True)
    out = net(data, dm)
    loss = bench.step_losses(out, prob)
    loss.backward()
    opt.step()
for _ in range(3): step()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    step()
    torch.cuda.synchronize()
cnt, tim = collections.Counter(), collections.Counter()
for ev in prof.events():
    if ev.device_type == torch.autograd.DeviceType.CUDA:
        cnt[ev.name[:90]] += 1
        tim[ev.name[:90]] += ev.device_time
print('total device events', sum(cnt.values()), 'device us', sum(tim.values()))
for k, v in cnt.most_common(60):
    print(f'{v:4d} {tim[k]:9.1f} us  {k}')
