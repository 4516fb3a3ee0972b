"""Generates tests/golden/ref_masks_n4.npz by running the REFERENCE's own util/datamaker.py::make_dummy_mask and
vmask_to_fmask (read-only import, authoring container):   python tests/golden/make_golden_masks.py

util/datamaker.py imports torch_geometric.data.Data at module level (library absent): the import is satisfied with
semigcn_b200's attribute-bag Data; the two functions exercised here do not touch it.  The reference Mesh is built with
build_mat=True (dense N x N AdjI, util/mesh.py:271-272 -- fine at 162 vertices).  The .ply dumps of make_dummy_mask go to
a scratch directory.  np.random.seed(314) is the seed the test replays.
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

from semigcn_b200 import data as sdata                 # noqa: E402
from semigcn_b200.meshgen import icosphere, write_obj  # noqa: E402

sys.modules.setdefault("turtle", types.SimpleNamespace(pd=None))
pyg = types.ModuleType("torch_geometric")
pyg_data = types.ModuleType("torch_geometric.data")
pyg_data.Data = sdata.Data
pyg.data = pyg_data
sys.modules.setdefault("torch_geometric", pyg)
sys.modules.setdefault("torch_geometric.data", pyg_data)
sys.path.insert(0, "/root/reference")
from util.mesh import Mesh            # noqa: E402
import util.datamaker as RD           # noqa: E402


def main():
    n = 4
    m = icosphere(n, dtype=torch.float64)
    os.makedirs("/tmp/golden_masks/dummy_mask", exist_ok=True)
    path = "/tmp/golden_masks/ico4.obj"
    write_obj(path, m.vs, m.faces)
    rm = Mesh(path, build_mat=True)
    np.random.seed(314)
    exist_face = np.ones(len(rm.faces))
    vmask, fmask = RD.make_dummy_mask(rm, dm_size=6, kn=[1, 2, 3], exist_face=exist_face)
    g = torch.Generator().manual_seed(7)
    v_real = (torch.rand(len(rm.vs), generator=g) > 0.15).numpy()
    f_real = RD.vmask_to_fmask(rm, v_real.astype(np.float32))
    out = dict(faces=rm.faces, edge_index=rm.edge_index.numpy(), vmask_dummy=vmask.numpy(), fmask_dummy=fmask.numpy(),
               v_real=v_real, f_real=f_real.numpy(), seed=np.array(314), dm_size=np.array(6), kn=np.array([1, 2, 3]))
    np.savez_compressed(os.path.join(HERE, "ref_masks_n4.npz"), **out)
    print("wrote ref_masks_n4.npz", vmask.shape, fmask.shape, "holes per mask", (1 - vmask).sum(0)[:6])


if __name__ == "__main__":
    main()
