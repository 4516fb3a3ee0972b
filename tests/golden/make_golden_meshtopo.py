"""Generates tests/golden/ref_meshtopo_n{4,8}.npz by running the REFERENCE's own util/mesh.py::Mesh (build_mat=True) and
Mesh.mesh_merge (read-only import, authoring container):   python tests/golden/make_golden_meshtopo.py

Pins the topology tables / sparse matrices of the ``semigcn_b200.mesh.Mesh`` drop-in and the refinement solve:
  edges, edge_index, f2f, fn, fa, fc, vn, v_dims, vf (as sorted CSR), Lap / AdjI / Adj / f2v_mat (indices + values of the
  reference's coalesced sparse tensors), and ``mesh_merge`` -- the reference's fp32 dense solve plus an fp64 dense solve of the
  SAME normal equations (numpy), which is what an iterative solver can be held to tightly.
/root/reference does not exist on the GPU box: tests only read the committed .npz files.
"""
import os
import sys
import types
import warnings

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))
warnings.filterwarnings("ignore")

from semigcn_b200.meshgen import icosphere, write_obj  # noqa: E402

sys.modules.setdefault("turtle", types.SimpleNamespace(pd=None))   # util/mesh.py:1 needs tkinter
sys.path.insert(0, "/root/reference")
from util.mesh import Mesh            # noqa: E402


def sp(t):
    t = t.coalesce()
    return t.indices().numpy(), t.values().numpy()


def main():
    for n in (4, 8):
        m = icosphere(n, dtype=torch.float64)
        g = torch.Generator().manual_seed(77 + n)
        vs = m.vs * (1.0 + 0.05 * torch.randn(m.vs.shape[0], 1, generator=g, dtype=torch.float64))
        path = f"/tmp/golden_topo{n}.obj"
        write_obj(path, vs, m.faces)
        rm = Mesh(path, build_mat=True)
        nv = len(rm.vs)
        vf_rowptr = np.cumsum([0] + [len(s) for s in rm.vf])
        vf_faces = np.concatenate([np.sort(np.array(list(s), dtype=np.int64)) for s in rm.vf])
        out = dict(obj_vs=vs.numpy(), vs=rm.vs, faces=rm.faces, edges=rm.edges, edge_index=rm.edge_index.numpy(), f2f=rm.f2f,
                   fn=rm.fn, fa=rm.fa, fc=rm.fc, vn=rm.vn, v_dims=rm.v_dims.numpy(), vf_rowptr=vf_rowptr, vf_faces=vf_faces)
        for name in ("Lap", "AdjI", "Adj", "f2v_mat"):
            i, v = sp(getattr(rm, name))
            out[name + "_idx"], out[name + "_val"] = i, v
        # refinement: a hole of ~8 % of the vertices (a cap), network output = the original positions + noise
        preserve = torch.from_numpy(rm.vs[:, 2] < 0.75 * rm.vs[:, 2].max())
        new_pos = torch.from_numpy(rm.vs).float() + 0.02 * torch.randn(nv, 3, generator=g)
        ref = Mesh.mesh_merge(rm.Lap, rm, new_pos, preserve, w=1.0)
        # fp64 dense solve of the same normal equations
        L = rm.Lap.to_dense().double().numpy()
        org = torch.from_numpy(rm.vs).float().double().numpy()
        keep = preserve.numpy()
        adji = rm.AdjI.to_dense().double().numpy()
        s_set = (adji @ (1.0 - keep.astype(np.float64))) == 0
        b_mix = L @ new_pos.double().numpy()
        b_mix[s_set] = (L @ org)[s_set]
        D = np.diag(s_set.astype(np.float64))
        x64 = np.linalg.solve(L.T @ L + D, L.T @ b_mix + D @ org)
        out.update(merge_preserve=keep, merge_new_pos=new_pos.numpy(), merge_ref_f32=ref.numpy(), merge_f64=x64)
        np.savez_compressed(os.path.join(HERE, f"ref_meshtopo_n{n}.npz"), **out)
        print(n, nv, "hole vertices", int((~keep).sum()), "ref vs f64", float(np.abs(ref.numpy() - x64).max()))


if __name__ == "__main__":
    main()
