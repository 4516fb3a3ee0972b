"""``Mesh``-shaped drop-in for the topology tables and the refinement solve of the reference.

The reference's ``util/mesh.py::Mesh`` builds its tables with Python loops over faces (``build_gemm`` :60-100, ``build_vf``
:191-227) and DENSE N x N matrices (``build_v2v`` :254-274: ``Adj.to_dense()``, ``torch.eye(N)``), and ``mesh_merge``
(:679-698) solves the refinement with a dense ``torch.linalg.solve`` of an N x N system: 20 s at 10 k vertices, impossible
at 1 M (4 TB).  This class exposes the same attributes, built with vectorised tensor ops on any device (sorting / scatter
over half-edges, O(F log F) time, O(F) memory) and SPARSE matrices, and solves the same least-squares problem with a
Jacobi-preconditioned conjugate gradient on the normal equations in fp64:

    vs, vc, faces, fn, fa, fc                 numpy, as the reference holds them (:17-19, :100-124)
    edges [E, 2]                              unique undirected edges in the reference's first-seen order (:69-87)
    edge_index [2, 2E]                        ``[edges.T || flipped]`` (:229-230)
    f2f [F, 3], f_edges                       faces across the three sides, -1 on a boundary (:215-227)
    vf                                        vertex -> incident faces, as CSR (``vf_rowptr``, ``vf_faces``); ``vf`` itself (a list of
                                              sets like the reference's) is materialised lazily, small meshes only
    v_dims, Adj, AdjI, Diag, Lap              degree and the sparse adjacency / uniform Laplacian I - D^-1 A (:254-274), built sparse
    f2v_mat, v2f_mat                          face <-> vertex incidence (:207-213)
    save(filename)                            OBJ writer (:704-732)
    Mesh.mesh_merge(lap, org_mesh, new_pos, preserve, w, w_b)   the refinement of sgcn.py:186-193 / refinement.py

Host-side preprocessing (one-off per mesh, not on the per-step path): plain tensor ops on the chosen device, no kernel of
ours involved.  Out of scope, as in DESIGN.md §9: QEM simplification, cotangent Laplacians, eigen-decompositions, PLY dumps.
"""
from __future__ import annotations

from typing import List, Optional

import numpy as np
import torch
from torch import Tensor

from . import meshgen


def read_obj(path: str):
    """``fill_from_file`` (util/mesh.py:35-58): ``v x y z [r g b]`` and triangular ``f a[/..] b c`` records, 1-based or negative ids."""
    vs, vc, faces = [], [], []
    with open(path) as fh:
        for line in fh:
            sp = line.split()
            if not sp:
                continue
            if sp[0] == "v":
                vs.append([float(v) for v in sp[1:4]])
                if len(sp) == 7:
                    vc.append([float(v) for v in sp[4:7]])
            elif sp[0] == "f":
                ids = [int(c.split("/")[0]) for c in sp[1:]]
                if len(ids) != 3:
                    raise ValueError(f"{path}: only triangular faces are supported")
                faces.append([(i - 1) if i >= 0 else (len(vs) + i) for i in ids])
    vs, vc, faces = np.asarray(vs, dtype=np.float64), np.asarray(vc, dtype=np.float64), np.asarray(faces, dtype=np.int64)
    if faces.size and not np.logical_and(faces >= 0, faces < len(vs)).all():
        raise ValueError(f"{path}: face index out of range")
    return vs, vc, faces


def _coo(idx: Tensor, val: Tensor, shape) -> Tensor:
    return torch.sparse_coo_tensor(idx, val, size=shape).coalesce()


class Mesh:
    def __init__(self, path: Optional[str] = None, manifold: bool = True, build_mat: bool = True, device="cpu",
                 vs: Optional[np.ndarray] = None, faces: Optional[np.ndarray] = None):
        """``Mesh(path)`` like the reference, or ``Mesh(vs=..., faces=...)`` from arrays.  ``device``: where the tables are
        computed (and where the sparse matrices live); the per-vertex / per-face arrays are numpy on the host like the
        reference's.  ``build_mat=False`` skips the matrices (the reference skips its mesh Laplacians)."""
        self.path = path
        if path is not None:
            self.vs, self.vc, self.faces = read_obj(path)
        else:
            self.vs, self.vc, self.faces = np.asarray(vs, dtype=np.float64), np.zeros((0,)), np.asarray(faces, dtype=np.int64)
        self.device = torch.device(device)
        self.simp = False
        self._vf: Optional[List[set]] = None
        dev = self.device
        v = torch.from_numpy(self.vs).to(dev)
        f = torch.from_numpy(self.faces).to(dev)
        nv, nf = v.shape[0], f.shape[0]
        # compute_face_normals / compute_face_center (util/mesh.py:100-124): float64, |n| + 1e-24
        n = torch.linalg.cross(v[f[:, 1]] - v[f[:, 0]], v[f[:, 2]] - v[f[:, 0]])
        self.fa = (0.5 * torch.sqrt((n ** 2).sum(dim=1))).cpu().numpy()
        self.fn = (n / (torch.linalg.norm(n, dim=1, keepdim=True) + 1e-24)).cpu().numpy()
        self.fc = (v[f].sum(dim=1) / 3.0).cpu().numpy()
        if not manifold:
            return
        edges = meshgen.edges_from_faces(f, nv)                       # first-seen order of build_gemm (:69-87)
        self.edges = edges.cpu().numpy()
        self.edges_count = int(edges.shape[0])
        self.edge_index = meshgen.edge_index_from_edges(edges).cpu()  # :229-230 (a CPU LongTensor like the reference's)
        ei = self.edge_index.to(dev)
        # vertex normals (:107-119): sum of incident face normals, l2-normalised rows (sklearn normalize: zero rows stay zero)
        vn = torch.zeros((nv, 3), dtype=torch.float64, device=dev).index_add_(0, f.reshape(-1), torch.from_numpy(self.fn).to(dev).repeat_interleave(3, dim=0))
        ln = torch.linalg.norm(vn, dim=1, keepdim=True)
        self.vn = (vn / torch.where(ln > 0, ln, torch.ones_like(ln))).cpu().numpy()
        # vertex -> faces incidence as CSR (build_vf :191-197), faces ascending inside a row
        corner_v = f.reshape(-1)
        order = torch.sort(corner_v, stable=True)[1]
        self.vf_faces = (order // 3).cpu().numpy()
        rp = torch.zeros(nv + 1, dtype=torch.int64, device=dev)
        rp[1:] = torch.cumsum(torch.bincount(corner_v, minlength=nv), 0)
        self.vf_rowptr = rp.cpu().numpy()
        # faces across the three sides (:215-227)
        f2f = meshgen.face_adjacency(f)
        self.f2f = f2f.cpu().numpy()
        has = f2f >= 0
        src = torch.arange(nf, device=dev).unsqueeze(1).expand(nf, 3)[has]
        self.f_edges = torch.stack([src, f2f[has]]).cpu().numpy()
        self.face_index = torch.from_numpy(self.f_edges)
        # degree and sparse matrices (build_v2v :254-274), no dense N x N anywhere
        ones = torch.ones(ei.shape[1], dtype=torch.float32, device=dev)
        self.v_dims = torch.zeros(nv, dtype=torch.float32, device=dev).index_add_(0, ei[0], ones).cpu()
        if build_mat:
            self._build_matrices(ei, f, nv, nf)

    # ------------------------------------------------------------------------------------------------------------
    def _build_matrices(self, ei: Tensor, f: Tensor, nv: int, nf: int) -> None:
        dev = ei.device
        ones = torch.ones(ei.shape[1], dtype=torch.float32, device=dev)
        diag = torch.arange(nv, device=dev).repeat(2, 1)
        deg = self.v_dims.to(dev)
        self.Adj = _coo(ei, ones, (nv, nv))
        self.Diag = _coo(diag, 1.0 / deg, (nv, nv))
        self.AdjI = _coo(torch.cat([ei, diag], dim=1), torch.cat([ones, torch.ones(nv, dtype=torch.float32, device=dev)]), (nv, nv))
        # Lap = I - D^-1 A: entry (i, j) = -1 / deg_i for an edge, 1 on the diagonal
        self.Lap = _coo(torch.cat([ei, diag], dim=1), torch.cat([-1.0 / deg[ei[0]], torch.ones(nv, dtype=torch.float32, device=dev)]), (nv, nv))
        corner_v = f.reshape(-1)
        corner_f = torch.arange(nf, device=dev).repeat_interleave(3)
        inc = torch.ones(3 * nf, dtype=torch.float32, device=dev)
        self.v2f_mat = _coo(torch.stack([corner_v, corner_f]), inc, (nv, nf))
        self.f2v_mat = _coo(torch.stack([corner_f, corner_v]), inc, (nf, nv))

    @property
    def vf(self) -> List[set]:
        """The reference's list of face-id sets (util/mesh.py:191-197); materialised on first use (Python objects: small meshes)."""
        if self._vf is None:
            rp, ff = self.vf_rowptr, self.vf_faces
            self._vf = [set(ff[rp[i]:rp[i + 1]].tolist()) for i in range(len(self.vs))]
        return self._vf

    # ------------------------------------------------------------------------------------------------------------
    def save(self, filename: str, color: bool = False) -> None:
        """OBJ writer with the reference's formatting (util/mesh.py:704-732)."""
        vs = np.asarray(self.vs, dtype=np.float32)
        vc = np.asarray(self.vc, dtype=np.float32).reshape(-1, 3) if np.size(self.vc) else np.zeros((0, 3), dtype=np.float32)
        with open(filename, "w") as fp:
            if len(vc) == 0 or not color:
                for x, y, z in vs:
                    fp.write("v {0:.8f} {1:.8f} {2:.8f}\n".format(x, y, z))
            else:
                for (x, y, z), (c1, c2, c3) in zip(vs, vc):
                    fp.write("v {0:.8f} {1:.8f} {2:.8f} {3:.8f} {4:.8f} {5:.8f}\n".format(x, y, z, c1, c2, c3))
            for a, b, c in np.asarray(self.faces, dtype=np.int64) + 1:
                fp.write("f {0} {1} {2}\n".format(a, b, c))

    # ------------------------------------------------------------------------------------------------------------
    @staticmethod
    def mesh_merge(lap, org_mesh, new_pos: Tensor, preserve: Tensor, w: float = 1, w_b: float = 0, tol: float = 1e-10,
                   max_iter: int = 20000, return_info: bool = False):
        """The refinement of sgcn.py:186-193 / refinement.py (util/mesh.py:679-698): the least-squares problem

            min_x  |L x - b|^2  +  w^2 |x_S - org_S|^2  +  w_b^2 |x_B - org_B|^2,
            b = L new_pos, rows S replaced by (L org)_S;   S = vertices whose closed 1-ring is entirely preserved,
            B = the other preserved vertices (the rim of the holes)

        solved on its normal equations  (L^T L + w^2 P_S + w_b^2 P_B) x = L^T b + w^2 P_S org + w_b^2 P_B org  by a
        Jacobi-preconditioned conjugate gradient with SPARSE matrix-vector products in fp64 -- the reference densifies A
        and calls ``torch.linalg.solve`` on the N x N system (fp32).  ``lap``: a sparse [N, N] matrix (``mesh.Lap``);
        ``org_mesh``: anything with ``vs`` and ``AdjI``; runs on ``lap``'s device; returns fp32 like the reference."""
        dev = lap.device
        L = lap.coalesce().to(torch.float64)
        Lc = torch.sparse_coo_tensor(L.indices(), L.values(), L.shape).to_sparse_csr()
        Lt = torch.sparse_coo_tensor(L.indices()[[1, 0]], L.values(), (L.shape[1], L.shape[0])).coalesce().to_sparse_csr()
        org = torch.as_tensor(np.asarray(org_mesh.vs)).to(dev).to(torch.float32).to(torch.float64)       # the reference rounds org to fp32 first
        new = torch.as_tensor(new_pos).to(dev).to(torch.float64)
        keep = torch.as_tensor(preserve).to(dev).reshape(-1).bool()
        adji = org_mesh.AdjI.to(dev).coalesce().to(torch.float64)
        hole_nb = torch.sparse.mm(adji, (1.0 - keep.to(torch.float64)).reshape(-1, 1)).reshape(-1)
        s_set = hole_nb == 0
        b_set = torch.logical_xor(keep, s_set)
        dvec = (float(w) ** 2) * s_set.to(torch.float64) + (float(w_b) ** 2) * b_set.to(torch.float64)
        b_mix = torch.sparse.mm(Lc, new)
        b_mix[s_set] = torch.sparse.mm(Lc, org)[s_set]
        rhs = torch.sparse.mm(Lt, b_mix) + dvec.unsqueeze(1) * org

        def apply(x):
            return torch.sparse.mm(Lt, torch.sparse.mm(Lc, x)) + dvec.unsqueeze(1) * x

        vals2 = L.values() ** 2
        jac = torch.zeros(L.shape[1], dtype=torch.float64, device=dev).index_add_(0, L.indices()[1], vals2) + dvec     # diag(L^T L) + D
        minv = (1.0 / jac).unsqueeze(1)
        x = torch.where(keep.unsqueeze(1), org, new).clone()           # start from the merged positions
        r = rhs - apply(x)
        z = minv * r
        p = z.clone()
        rz = (r * z).sum(dim=0)
        rhs_norm = torch.linalg.norm(rhs, dim=0).clamp_min(1e-300)
        it = 0
        for it in range(1, int(max_iter) + 1):
            ap = apply(p)
            alpha = rz / (p * ap).sum(dim=0).clamp_min(1e-300)
            x = x + alpha * p
            r = r - alpha * ap
            if it % 10 == 0 and bool((torch.linalg.norm(r, dim=0) <= tol * rhs_norm).all()):
                break
            z = minv * r
            rz_new = (r * z).sum(dim=0)
            p = z + (rz_new / rz.clamp_min(1e-300)) * p
            rz = rz_new
        out = x.to(torch.float32)
        if return_info:
            return out, {"iterations": it, "relative_residual": float((torch.linalg.norm(rhs - apply(x), dim=0) / rhs_norm).max())}
        return out
