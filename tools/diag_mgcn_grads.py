"""Per-parameter gradient error of the MGCN drop-in and of the fp32 oracle against an fp64 oracle evaluation."""
import sys
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
import numpy as np, torch
from helpers import load_golden
from test_oracle_meshnet import _hier
from oracle import meshnet_ref as M
from semigcn_b200.data import Data
from semigcn_b200.meshnet import MGCN
DEV = torch.device('cuda:0')
gold = load_golden("ref_meshnet_n4.npz")
_, e, ph, uh, sm = _hier(gold)
skip = len(sys.argv) > 1 and sys.argv[1] == 'skip'
torch.manual_seed(int(gold["seed"]))
ref = M.MGCN(e, ph, uh, sm, skip=skip, drop_rate=0.0)
ref64 = M.MGCN(e, [p.double() for p in ph], [u.double() for u in uh], sm, skip=skip, drop_rate=0.0)
ref64.load_state_dict({k: v for k, v in ref.state_dict().items() if not v.is_sparse}, strict=False)
ref64 = ref64.double(); ref64.smposs_list = [s.double() for s in sm]
net = MGCN(DEV, e, ph, uh, sm, skip=skip, drop_rate=0.0)
net.load_state_dict(ref.state_dict()); net = net.to(DEV)
z1 = torch.from_numpy(gold["z1"])
tgt = [torch.from_numpy(gold[f"smpos_{l}"]) * 1.01 for l in range(4)]
sum(((y - t) ** 2).mean() for y, t in zip(ref(z1, gold["dm"]), tgt)).backward()
sum(((y - t.double()) ** 2).mean() for y, t in zip(ref64(z1.double(), gold["dm"].astype(np.float64)), tgt)).backward()
ys = net(Data(z1=z1.to(DEV), x_pos=z1.to(DEV)), gold["dm"])
sum(((y - t.to(DEV)) ** 2).mean() for y, t in zip(ys, tgt)).backward()
gmax = max(p.grad.abs().max().item() for p in ref64.parameters() if p.grad is not None)
rows = []
for (k, p), (_, q), (_, r) in zip(net.named_parameters(), ref.named_parameters(), ref64.named_parameters()):
    if r.grad is None or p.grad is None:
        rows.append((0, 0, k, 'NO GRAD ours=%s ref=%s' % (p.grad is None, r.grad is None))); continue
    scale = max(r.grad.abs().max().item(), 1e-3 * gmax)
    rows.append(((p.grad.cpu().double() - r.grad).abs().max().item() / scale, (q.grad.double() - r.grad).abs().max().item() / scale, k,
                 '|g|=%.2e' % r.grad.abs().max().item()))
print('gmax', gmax)
for a, b, k, extra in rows:
    print(f'ours {a:.2e}  oracle32 {b:.2e}  {k}  {extra}')
