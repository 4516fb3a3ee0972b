"""Generates tests/golden/ref_meshnet_n4.npz by executing the REFERENCE's own util/meshnet.py (read-only, in the
authoring container):   python tests/golden/make_golden_meshnet.py

torch_geometric is not installable here, so GCNConv / ChebConv / Sequential resolve to the ORACLE classes
(oracle/pyg_ref.py); everything else is the reference's code running: MeshPool / MeshUnpool (util/meshnet.py:9-27),
DownConv / UpConv (:31-160) and MGCN.forward (:278-317).  MGCN.__init__ cannot run (it QEM-simplifies Mesh objects,
:169-193, host-side and out of scope), so the instance is assembled here from the reference's own block classes in the
reference's construction order (:212-248) on a synthetic hierarchy (semigcn_b200/meshgen.py::synth_pool_hierarchy) and the
reference's unbound forward is called on it.  Dropout: the fixture is generated in train mode with the reference's
drop rates set to 0.0 through the constructor argument the reference itself exposes, and in eval mode with 0.2.
The .npz stores inputs, the hierarchy, outputs and state_dict checksums.  /root/reference does not exist on the GPU box.
"""
import os
import sys
import types
import warnings

import numpy as np
import torch
import torch.nn as nn

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.abspath(os.path.join(HERE, "..", ".."))
sys.path.insert(0, ROOT)
warnings.filterwarnings("ignore")

from oracle import pyg_ref as O                      # noqa: E402
from semigcn_b200 import meshgen                     # noqa: E402

REF_FILE = "/root/reference/util/meshnet.py"
SEED = 314


def load_reference_meshnet():
    pyg = types.ModuleType("torch_geometric")
    pyg_nn = types.ModuleType("torch_geometric.nn")
    pyg_nn.GCNConv, pyg_nn.ChebConv, pyg_nn.Sequential = O.GCNConv, O.ChebConv, O.Sequential
    pyg.nn = pyg_nn
    sys.modules["torch_geometric"], sys.modules["torch_geometric.nn"] = pyg, pyg_nn
    mod = types.ModuleType("ref_meshnet")
    exec(compile(open(REF_FILE).read(), REF_FILE, "exec"), mod.__dict__)
    return mod


def checksums(state_dict):
    names = sorted(k for k, v in state_dict.items() if not v.is_sparse)
    sums = np.array([[float(state_dict[k].double().sum()), float(state_dict[k].double().abs().sum())] for k in names])
    return np.array(names), sums


def assemble_mgcn(R, hier, smposs, skip, drop):
    """The reference's MGCN module tree (util/meshnet.py:212-248) without its Mesh-dependent __init__."""
    net = R.MGCN.__new__(R.MGCN)
    nn.Module.__init__(net)
    e, ph, uh = hier["edge_inds"], hier["p_hashes"], hier["up_hashes"]
    net.device, net.skip, net.edge_inds, net.smposs_list = "cpu", skip, e, smposs
    net.encoder1 = R.DownConv(4, 32, e[0], e[1], ph[0], K=3, drop_rate=0.0)
    net.encoder2 = R.DownConv(32, 128, e[1], e[2], ph[1], K=3, drop_rate=drop)
    net.encoder3 = R.DownConv(128, 256, e[2], e[3], ph[2], K=3, drop_rate=drop)
    net.decoder3 = R.UpConv(256, 128, e[3], e[2], uh[2], K=3, drop_rate=drop)
    net.decoder2 = R.UpConv(128, 32, e[2], e[1], uh[1], K=3, drop_rate=drop)
    net.decoder1 = nn.Sequential(R.UpConv(32, 16, e[1], e[0], uh[0], K=3, drop_rate=0.0), nn.Linear(16, 3))
    for name, cin in (("mcnn3", 256), ("mcnn2", 128), ("mcnn1", 32)):
        setattr(net, name, R.Sequential("x, edge_index", [(R.ChebConv(cin, 32, K=3), "x, edge_index -> x"), (nn.BatchNorm1d(32), "x -> x"),
                                                           (nn.LeakyReLU(), "x -> x"), (nn.Linear(32, 3), "x -> x")]))
    net.skip2 = nn.Linear(256, 128)
    net.skip1 = nn.Linear(64, 32)
    return net


def main():
    R = load_reference_meshnet()
    prob = meshgen.synth_inpainting_problem(4, smooth_iters=5, n_dummy=2)
    mesh = prob["mesh"]
    hier = meshgen.synth_pool_hierarchy(mesh, levels=3, ratio=0.6, seed=SEED)
    out = {"seed": np.array(SEED), "z1": prob["z1"].numpy(), "x_pos": prob["x_pos"].numpy(), "dm": prob["vmask_dummy"][:, :1].numpy(),
           "sizes": np.array(hier["sizes"])}
    for l, e in enumerate(hier["edge_inds"]):
        out[f"edge_index_{l}"] = e.numpy()
    for l, p in enumerate(hier["p_hashes"]):
        out[f"pool_idx_{l}"] = p.indices().numpy()          # [2, n_fine]: (coarse, fine); unpool = the transpose
    # smoothed positions per level: level 0 given, coarser ones pooled (as util/meshnet.py:253-275 does for org_pos)
    smposs = [prob["x_pos"]]
    for l in range(3):
        smposs.append(R.MeshPool(hier["p_hashes"][l])(smposs[-1]))
    for l, p in enumerate(smposs):
        out[f"smpos_{l}"] = p.numpy()

    # ---- MeshPool / MeshUnpool alone
    g = torch.Generator().manual_seed(SEED)
    x_f = torch.randn(hier["sizes"][0], 8, generator=g)
    x_c = torch.randn(hier["sizes"][1], 8, generator=g)
    out["pool_in"], out["pool_out"] = x_f.numpy(), R.MeshPool(hier["p_hashes"][0])(x_f).numpy()
    out["unpool_in"], out["unpool_out"] = x_c.numpy(), R.MeshUnpool(hier["up_hashes"][0])(x_c).numpy()

    # ---- DownConv / UpConv blocks (train mode, drop 0)
    e = hier["edge_inds"]
    torch.manual_seed(SEED)
    down = R.DownConv(4, 16, e[0], e[1], hier["p_hashes"][0], K=3, drop_rate=0.0)
    up = R.UpConv(16, 8, e[1], e[0], hier["up_hashes"][0], K=3, drop_rate=0.0)
    x4 = torch.randn(hier["sizes"][0], 4, generator=g)
    out["block_in"] = x4.numpy()
    out["down_names"], out["down_sums"] = checksums(down.state_dict())
    out["up_names"], out["up_sums"] = checksums(up.state_dict())
    y_d = down(x4)
    y_u = up(y_d)
    out["down_out"], out["up_out"] = y_d.detach().numpy(), y_u.detach().numpy()

    # ---- whole MGCN through the reference's forward
    data = types.SimpleNamespace(z1=prob["z1"], x_pos=prob["x_pos"])
    dm = prob["vmask_dummy"][:, :1].numpy()
    for skip in (False, True):
        tag = "skip" if skip else "noskip"
        torch.manual_seed(SEED)
        net = assemble_mgcn(R, hier, smposs, skip, drop=0.0)
        out[f"mgcn_{tag}_names"], out[f"mgcn_{tag}_sums"] = checksums(net.state_dict())
        net.train()
        ys = R.MGCN.forward(net, data, dm)                         # np.ndarray mask: honoured
        for l, y in enumerate(ys):
            out[f"mgcn_{tag}_train_{l}"] = y.detach().numpy()
        ys_t = R.MGCN.forward(net, data, torch.from_numpy(dm))      # torch mask: the reference substitutes ones (:290-293)
        out[f"mgcn_{tag}_train_tensormask_0"] = ys_t[0].detach().numpy()
        torch.manual_seed(SEED)
        net2 = assemble_mgcn(R, hier, smposs, skip, drop=0.2)
        net2.eval()
        ys = R.MGCN.forward(net2, data, dm)
        for l, y in enumerate(ys):
            out[f"mgcn_{tag}_eval_{l}"] = y.detach().numpy()
        print(tag, "params", sum(p.numel() for p in net.parameters()), [float(y.abs().max()) for y in ys])
    np.savez_compressed(os.path.join(HERE, "ref_meshnet_n4.npz"), **out)
    print("wrote ref_meshnet_n4.npz", os.path.getsize(os.path.join(HERE, "ref_meshnet_n4.npz")), "bytes")


if __name__ == "__main__":
    main()
