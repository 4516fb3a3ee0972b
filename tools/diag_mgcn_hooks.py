"""Where does the skip-mode gradient discrepancy enter?  Captures gradients of intermediate tensors of decoder1."""
import sys
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
import numpy as np, torch
from helpers import load_golden, rel_err
from test_oracle_meshnet import _hier
from oracle import meshnet_ref as M
from semigcn_b200.data import Data
from semigcn_b200.meshnet import MGCN
DEV = torch.device('cuda:0')
gold = load_golden("ref_meshnet_n4.npz")
_, e, ph, uh, sm = _hier(gold)
torch.manual_seed(int(gold["seed"]))
ref = M.MGCN(e, ph, uh, sm, skip=True, drop_rate=0.0)
net = MGCN(DEV, e, ph, uh, sm, skip=True, drop_rate=0.0)
net.load_state_dict(ref.state_dict()); net = net.to(DEV)
z1 = torch.from_numpy(gold["z1"])
tgt = [torch.from_numpy(gold[f"smpos_{l}"]) * 1.01 for l in range(4)]

def instrument(model, store):
    def mk(name):
        def hook(mod, inp, out):
            if isinstance(out, torch.Tensor) and out.requires_grad:
                store[name + '.out'] = out
                out.register_hook(lambda g, n=name: store.__setitem__(n + '.grad', g.detach().clone()))
        return hook
    for name, mod in model.named_modules():
        if name.startswith('decoder1') or name in ('skip1', 'mcnn1', 'decoder2'):
            mod.register_forward_hook(mk(name))

s_o, s_r = {}, {}
instrument(net, s_o); instrument(ref, s_r)
sum(((y - t) ** 2).mean() for y, t in zip(ref(z1, gold["dm"]), tgt)).backward()
ys = net(Data(z1=z1.to(DEV), x_pos=z1.to(DEV)), gold["dm"])
sum(((y - t.to(DEV)) ** 2).mean() for y, t in zip(ys, tgt)).backward()
torch.cuda.synchronize()
for k in sorted(s_r):
    if k in s_o:
        print(f'{k:50s} {rel_err(s_o[k], s_r[k]):.2e}   shape {tuple(s_r[k].shape)}')
    else:
        print(f'{k:50s} missing in drop-in')
for k in ('decoder1.0.model1.module_2.weight', 'decoder1.0.model1.module_2.bias'):
    p, q = dict(net.named_parameters())[k], dict(ref.named_parameters())[k]
    print(k, rel_err(p.grad, q.grad), p.grad[:4].cpu().numpy(), q.grad[:4].numpy())

# ---- isolate: recompute the BN+LeakyReLU backward of decoder1.0.model1.module_2 from the captured tensors
import torch.nn.functional as F
from semigcn_b200 import ops
bn = net.decoder1[0].model1.module_2
y = s_o['decoder1.0.model1.module_1.out'].detach()
dz = s_o['decoder1.0.model1.grad']
dy_ours = s_o['decoder1.0.model1.module_1.grad']
yt = y.clone().requires_grad_(True)
zt = F.leaky_relu(F.batch_norm(yt, None, None, bn.weight.detach(), bn.bias.detach(), True, 0.1, bn.eps), 0.01)
zt.backward(dz)
print('captured drop-in dy vs torch BN backward on the same (y, dz):', rel_err(dy_ours, yt.grad))
# the kernel path again, standalone, on the same tensors
y2 = y.clone().requires_grad_(True)
bn2 = torch.nn.BatchNorm1d(16).to(DEV)
bn2.load_state_dict(bn.state_dict())
z2 = ops.bn_act(y2, bn2, 0.01)
z2.backward(dz.clone())
print('standalone ops.bn_act backward vs torch:', rel_err(y2.grad, yt.grad), ' forward:', rel_err(z2, zt))
torch.save({'y': y.cpu(), 'dz': dz.cpu(), 'dy_ours': dy_ours.cpu(), 'dy_torch': yt.grad.cpu(), 'gamma': bn.weight.detach().cpu(), 'beta': bn.bias.detach().cpu(),
            'dbeta_ours': bn.bias.grad.cpu(), 'dgamma_ours': bn.weight.grad.cpu()}, 'gpurun_out/bn_case.pt')
print('dz strides', dz.stride(), 'y strides', y.stride(), 'dz ptr % 16', dz.data_ptr() % 16, 'y ptr % 16', y.data_ptr() % 16)
