"""Generates tests/golden/ref_mesh_n{4,8}.npz by IMPORTING THE REFERENCE (read-only) in the
authoring container:  python tests/golden/make_golden.py

Pins, bit for bit on CPU, the reference-owned pieces of the step path:
  * util/mesh.py::Mesh   -> edges / edge_index ordering (:60-100, :229-230), v_dims (:267),
                            face normals fn, f2f
  * util/models.py::compute_fn (:121-126)
  * util/loss.py::mask_pos_rec_loss (:14-34), mask_norm_rec_loss (:78-107),
                  mesh_laplacian_loss (:60-76), fn_bnf_detach_loss (:197-253)
The PyG conv arithmetic itself cannot be pinned this way (library absent) -- see
oracle/pyg_ref.py header.  /root/reference does not exist on the GPU box: tests only read
the committed .npz files.
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
import util.loss as RLoss             # noqa: E402
import util.models as RModels         # noqa: E402


def main():
    for n in (4, 8):
        m = icosphere(n, dtype=torch.float64)
        g = torch.Generator().manual_seed(314 + n)
        vs = m.vs * (1.0 + 0.05 * torch.randn(m.vs.shape[0], 1, generator=g, dtype=torch.float64))
        path = f"/tmp/golden_ico{n}.obj"
        write_obj(path, vs, m.faces)
        rm = Mesh(path, build_mat=False)
        nv, nf = len(rm.vs), len(rm.faces)
        pred = torch.from_numpy(rm.vs).float() + 0.01 * torch.randn(nv, 3, generator=g)
        pred64 = pred.double()
        v_mask = torch.rand(nv, generator=g) > 0.2
        f_mask = torch.rand(nf, generator=g) > 0.2
        fn_pred = RModels.compute_fn(pred, rm.faces)
        fn_pred64 = RModels.compute_fn(pred64, rm.faces)
        out = dict(
            vs=rm.vs, faces=rm.faces, edges=rm.edges, edge_index=rm.edge_index.numpy(),
            v_dims=rm.v_dims.numpy(), fn=rm.fn, f2f=rm.f2f,
            pred=pred.numpy(), v_mask=v_mask.numpy(), f_mask=f_mask.numpy(),
            compute_fn_f32=fn_pred.numpy(), compute_fn_f64=fn_pred64.numpy(),
            # sgcn.py passes float64 numpy targets (sgcn.py:127,131-132) -> float64 losses
            loss_pos_f64=RLoss.mask_pos_rec_loss(pred, rm.vs, v_mask.numpy()).numpy(),
            loss_norm_f64=RLoss.mask_norm_rec_loss(fn_pred, rm.fn, f_mask.numpy()).numpy(),
            # mgcn.py passes float32 tensors (mgcn.py:140)
            loss_pos_f32=RLoss.mask_pos_rec_loss(pred, torch.from_numpy(rm.vs).float(), v_mask.numpy()).numpy(),
            loss_norm_f32=RLoss.mask_norm_rec_loss(fn_pred, torch.from_numpy(rm.fn).float(), f_mask.numpy()).numpy(),
            loss_lap_f32=RLoss.mesh_laplacian_loss(pred, rm).numpy(),
        )
        bnf_loss, bnf_fn = RLoss.fn_bnf_detach_loss(pred, fn_pred, rm, loop=5)
        out["loss_bnf_f32"] = bnf_loss.numpy()
        out["bnf_fn_f32"] = bnf_fn.numpy()
        np.savez_compressed(os.path.join(HERE, f"ref_mesh_n{n}.npz"), **out)
        print(n, nv, nf, {k: (v.shape, v.dtype) for k, v in out.items() if hasattr(v, "shape")})


if __name__ == "__main__":
    main()
