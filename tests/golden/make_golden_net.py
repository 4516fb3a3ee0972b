"""Generates tests/golden/ref_sgcn_n4.npz by running the REFERENCE's own network file
(/root/reference/util/networks.py::SingleScaleGCN, read-only) in the authoring container:

    python tests/golden/make_golden_net.py

torch_geometric is not installable here (oracle/pyg_ref.py header), so the three symbols that file
imports (GCNConv, ChebConv, Sequential) resolve to the ORACLE classes; everything else -- layer
widths, block order, input normalisation, mask concat, skip wiring, residual, construction order
(= RNG consumption under torch.manual_seed) -- is the reference's code executing.  This pins the
oracle's restated ``SingleScaleGCN`` (oracle/pyg_ref.py) and, through it, the drop-in's network mirror
against util/networks.py:9-103.  The reference hard-codes ``conv = "chebconv"`` (util/networks.py:13);
the GCNConv branch (util/networks.py:22-37) is reached by compiling the same source with that one
string literal replaced in memory (nothing is copied into the repo).
The fixture stores inputs, outputs and per-tensor checksums of the seeded initial state_dict (not the
1.4 M parameters themselves).  /root/reference does not exist on the GPU box: tests read the .npz.
"""
import os
import sys
import types
import warnings

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.abspath(os.path.join(HERE, "..", ".."))
sys.path.insert(0, ROOT)
warnings.filterwarnings("ignore")

from oracle import pyg_ref as O                      # noqa: E402
from semigcn_b200 import meshgen                     # noqa: E402

REF_FILE = "/root/reference/util/networks.py"
SEED = 314                                           # sgcn.py:76 torch_fix_seed(314)


def load_reference_networks(conv: str):
    pyg = types.ModuleType("torch_geometric")
    pyg_nn = types.ModuleType("torch_geometric.nn")
    pyg_nn.GCNConv, pyg_nn.ChebConv, pyg_nn.Sequential = O.GCNConv, O.ChebConv, O.Sequential
    pyg.nn = pyg_nn
    sys.modules["torch_geometric"], sys.modules["torch_geometric.nn"] = pyg, pyg_nn
    src = open(REF_FILE).read()
    assert src.count('conv = "chebconv"') == 1
    if conv == "gcnconv":
        src = src.replace('conv = "chebconv"', 'conv = "gcnconv"')
    mod = types.ModuleType(f"ref_networks_{conv}")
    exec(compile(src, REF_FILE, "exec"), mod.__dict__)
    return mod


def checksums(state_dict):
    names = sorted(state_dict.keys())
    sums = np.array([[float(state_dict[k].double().sum()), float(state_dict[k].double().abs().sum())] for k in names])
    return np.array(names), sums


def main():
    prob = meshgen.synth_inpainting_problem(4, smooth_iters=5, n_dummy=2)
    mesh = prob["mesh"]
    data = types.SimpleNamespace(z1=prob["z1"], x_pos=prob["x_pos"], edge_index=mesh.edge_index)
    dm = prob["vmask_dummy"][:, :1].contiguous()
    out = {"z1": prob["z1"].numpy(), "x_pos": prob["x_pos"].numpy(), "edge_index": mesh.edge_index.numpy(), "dm": dm.numpy(),
           "seed": np.array(SEED)}
    for conv in ("chebconv", "gcnconv"):
        ref_mod = load_reference_networks(conv)
        for skip in (False, True):
            torch.manual_seed(SEED)
            net = ref_mod.SingleScaleGCN("cpu", skip=skip)
            names, sums = checksums(net.state_dict())
            net.train()
            y_train = net(data, dm.numpy())                       # np.ndarray mask, as sgcn.py passes it
            net.eval()
            y_eval = net(data, dm)
            tag = f"{conv}_{'skip' if skip else 'noskip'}"
            out[f"{tag}_names"], out[f"{tag}_sums"] = names, sums
            out[f"{tag}_train"], out[f"{tag}_eval"] = y_train.detach().numpy(), y_eval.detach().numpy()
            print(tag, "params", sum(p.numel() for p in net.parameters()), "out", y_train.abs().max().item())
    np.savez_compressed(os.path.join(HERE, "ref_sgcn_n4.npz"), **out)
    print("wrote", os.path.join(HERE, "ref_sgcn_n4.npz"))


if __name__ == "__main__":
    main()
