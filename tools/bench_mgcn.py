"""MGCN (util/meshnet.py mirror) train-step timing on a synthetic hierarchy (BASELINE.json configs[1]):
   python tools/bench_mgcn.py [freq ...]     eager and whole-step CUDA graph, drop-in vs the CPU oracle port."""
import sys, time, json, torch
sys.path.insert(0, '.')
from semigcn_b200 import meshgen, losses
from semigcn_b200.data import Data
from semigcn_b200.meshnet import MGCN
from semigcn_b200.nn import MeshPool
from semigcn_b200.graphed import GraphedTrainStep
dev = torch.device('cuda:0')
out = []
for freq in [int(a) for a in sys.argv[1:]] or [32, 100]:
    prob = meshgen.synth_inpainting_problem(freq, device=dev, smooth_iters=10, n_dummy=8)
    mesh = prob['mesh']
    hier = meshgen.synth_pool_hierarchy(mesh)
    sm = [prob['x_pos']]
    ini = [prob['ini_vs'].float()]
    vmask = [prob['v_mask']]
    for l in range(3):
        pool = MeshPool(hier['p_hashes'][l]).to(dev)
        sm.append(pool(sm[-1])); ini.append(pool(ini[-1]))
        vmask.append(pool(vmask[-1].float().reshape(-1, 1)).reshape(-1) == 1.0)
    torch.manual_seed(314)
    net = MGCN(dev, hier['edge_inds'], hier['p_hashes'], hier['up_hashes'], sm, skip=False, drop_rate=0.0, tensor_masks=True).to(dev)
    opt = torch.optim.Adam(net.parameters(), lr=0.01, capturable=True)
    w = [1.0, 0.5, 0.25, 0.125]
    def loss_fn(poss):
        lp = sum(losses.fused_mask_pos_rec_loss(p, t, m) * wi for p, t, m, wi in zip(poss, ini, vmask, w))
        _, ln = losses.sgcn_step_losses(poss[0], mesh.faces, prob['ini_vs'], prob['fn'], prob['v_mask'], prob['f_mask'])
        return lp + 4.0 * ln
    data = Data(z1=prob['z1'], x_pos=prob['x_pos'])
    dm = prob['vmask_dummy'][:, :1].contiguous()
    def step():
        opt.zero_grad(set_to_none=True)
        loss = loss_fn(net(data, dm)); loss.backward(); opt.step(); return loss
    def timeit(fn, n=10):
        for _ in range(3): fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n): fn()
        e1.record(); torch.cuda.synchronize()
        return e0.elapsed_time(e1) / n
    ms_eager = timeit(step)
    g = GraphedTrainStep(net, loss_fn, opt, prob['z1'], prob['x_pos'], None, dm)
    ms_graph = timeit(lambda: g(dm))
    rec = {'freq': freq, 'vertices': hier['sizes'], 'eager_ms': ms_eager, 'graph_ms': ms_graph, 'loss': float(g.loss)}
    print(json.dumps(rec), flush=True)
