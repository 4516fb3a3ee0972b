"""Step-path mesh losses (reference util/models.py:121-126, util/loss.py:14-34,60-107), GPU side.

``sgcn_step_losses`` / ``sgcn_step_loss`` run the fused kernels of csrc/loss.cu (one forward pass
over vertices + faces, one per-vertex gather backward, no boolean-index gathers => no host sync per
step).  They mirror ``compute_fn`` + ``mask_pos_rec_loss`` ("rmse") + ``mask_norm_rec_loss``
("l1mae") with the same argument meaning; float64 targets give float64 losses exactly as
sgcn.py:127,131-132 does, float32 targets (mgcn.py:140) float32 ones.

The stand-alone ``compute_fn`` / ``mask_*`` functions below are kept for callers that need the
intermediate normals; they are plain torch ops on the caller's device.
"""
from __future__ import annotations

from collections import OrderedDict
from typing import Optional, Tuple

import numpy as np
import torch
from torch import Tensor

from . import _lib as L
from . import profile as _prof
from ._lib import SgbError, check, ptr, require_cuda, stream_ptr


# ----------------------------------------------------------------------------------------
# face topology cache: device faces + vertex -> face-corner incidence (for the backward gather)
# ----------------------------------------------------------------------------------------
class FaceTopology:
    def __init__(self, faces: Tensor, num_vertices: int):
        require_cuda(faces)
        lib = L.load()
        f = faces.to(torch.int64).contiguous()
        if f.dim() != 2 or f.shape[1] != 3:
            raise SgbError("faces must have shape [F, 3]")
        dev = f.device
        nf, n = int(f.shape[0]), int(num_vertices)
        self.faces, self.nf, self.n = f, nf, n
        self.rowptr = torch.empty(n + 1, dtype=torch.int32, device=dev)
        self.inc = torch.empty(max(3 * nf, 1), dtype=torch.int32, device=dev)
        err = torch.zeros(1, dtype=torch.int32, device=dev)
        wsb = lib.sgb_incidence_build_workspace_bytes(nf, n)
        ws = torch.empty(max(wsb, 16), dtype=torch.uint8, device=dev)
        with torch.cuda.device(dev):
            check(lib.sgb_incidence_build(ptr(f), nf, n, ptr(self.rowptr), ptr(self.inc), ptr(err), ptr(ws), wsb, stream_ptr(dev)),
                  "sgb_incidence_build")
        L.count(6)
        if int(err.item()) != 0:
            raise SgbError("faces contain vertex ids outside [0, num_vertices)")


_TOPO_CACHE: "OrderedDict[tuple, tuple]" = OrderedDict()


def topology_for(faces, num_vertices: int, device) -> FaceTopology:
    """Cached per faces object (the reference passes the same ``mesh.faces`` array every step, sgcn.py:130)."""
    if isinstance(faces, np.ndarray):
        key = ("np", id(faces), faces.shape, str(device), int(num_vertices))
    else:
        key = ("t", faces.data_ptr(), tuple(faces.shape), faces._version, str(faces.device), int(num_vertices))
    hit = _TOPO_CACHE.get(key)
    if hit is not None:
        _TOPO_CACHE.move_to_end(key)
        return hit[0]
    ft = torch.from_numpy(np.ascontiguousarray(faces)).to(device) if isinstance(faces, np.ndarray) else faces.to(device)
    topo = FaceTopology(ft, num_vertices)
    _TOPO_CACHE[key] = (topo, faces)
    while len(_TOPO_CACHE) > 8:
        _TOPO_CACHE.popitem(last=False)
    return topo


def _dev_tensor(x, device, dtype=None) -> Optional[Tensor]:
    if x is None:
        return None
    if isinstance(x, np.ndarray):
        x = torch.from_numpy(x)
    x = x.to(device)
    if dtype is not None and x.dtype != dtype:
        x = x.to(dtype)
    return x.contiguous()


def _mask_u8(mask, device, size: int) -> Optional[Tensor]:
    if mask is None:
        return None
    m = _dev_tensor(mask, device)
    if m.dtype == torch.bool:
        m = m.view(torch.uint8)
    elif m.dtype != torch.uint8:
        m = (m != 0).view(torch.uint8)
    if m.numel() != size:
        raise SgbError(f"mask has {m.numel()} entries, expected {size}")
    return m.reshape(-1).contiguous()


class StepLossFn(torch.autograd.Function):
    """(loss_p, loss_n) of one training step as a single autograd node."""

    @staticmethod
    def forward(ctx, pos: Tensor, topo: Optional[FaceTopology], tpos, tfn, vmask, fmask):
        lib = L.load()
        require_cuda(pos)
        if pos.dtype != torch.float32 or pos.dim() != 2 or pos.shape[1] != 3:
            raise SgbError("pos must be a float32 [N, 3] tensor")
        if pos.stride(1) != 1:
            pos = pos.contiguous()
        dev = pos.device
        n = int(pos.shape[0])
        ref = tpos if tpos is not None else tfn
        if ref is None:
            raise SgbError("at least one of the targets must be given")
        t64 = ref.dtype == torch.float64
        want = torch.float64 if t64 else torch.float32
        for t in (tpos, tfn):
            if t is not None and t.dtype != want:
                raise SgbError("target positions and target normals must share one dtype (float64 or float32)")
        nf = topo.nf if (topo is not None and tfn is not None) else 0
        rows = lib.sgb_loss_partial_rows()
        partials = torch.empty((rows, 4), dtype=torch.float64, device=dev)
        out = torch.empty(4, dtype=torch.float64, device=dev)
        with torch.cuda.device(dev), _prof.region("step_loss", 4.0 * 3 * n + (8.0 if t64 else 4.0) * 3 * (n + nf) + 24.0 * nf):
            check(lib.sgb_step_loss_fwd(ptr(pos), pos.stride(0), n, ptr(tpos), 1 if t64 else 0, ptr(vmask),
                                        ptr(topo.faces) if nf else None, nf, ptr(tfn) if nf else None, ptr(fmask) if nf else None,
                                        None, ptr(partials), ptr(out), stream_ptr(dev)), "sgb_step_loss_fwd")
        L.count(2)
        ctx.save_for_backward(pos, out)
        ctx.topo, ctx.tpos, ctx.tfn, ctx.vmask, ctx.fmask, ctx.t64, ctx.nf = topo, tpos, tfn, vmask, fmask, t64, nf
        res = out[:2]
        return res if t64 else res.to(torch.float32)

    @staticmethod
    def backward(ctx, dres: Tensor):
        lib = L.load()
        pos, out = ctx.saved_tensors
        dev = pos.device
        n = int(pos.shape[0])
        grads = dres.to(torch.float64).contiguous()
        dpos = torch.empty((n, 3), dtype=torch.float32, device=dev)
        topo, nf = ctx.topo, ctx.nf
        with torch.cuda.device(dev), _prof.region("step_loss", 4.0 * 6 * n + (8.0 if ctx.t64 else 4.0) * 3 * (n + nf) + 36.0 * nf):
            check(lib.sgb_step_loss_bwd(ptr(pos), pos.stride(0), n, ptr(ctx.tpos), 1 if ctx.t64 else 0, ptr(ctx.vmask),
                                        ptr(topo.faces) if nf else None, nf, ptr(ctx.tfn) if nf else None, ptr(ctx.fmask) if nf else None,
                                        ptr(topo.rowptr) if nf else None, ptr(topo.inc) if nf else None, ptr(out), ptr(grads),
                                        ptr(dpos), 3, stream_ptr(dev)), "sgb_step_loss_bwd")
        L.count(1)
        return dpos, None, None, None, None, None


def sgcn_step_losses(pos: Tensor, faces, ini_vs, fn_real, v_mask, f_mask) -> Tuple[Tensor, Tensor]:
    """(loss_p, loss_n) of sgcn.py:130-132: ``mask_pos_rec_loss(pos, ini_vs, v_mask)`` and
    ``mask_norm_rec_loss(compute_fn(pos, faces), fn_real, f_mask)``, fused."""
    dev = pos.device
    n = int(pos.shape[0])
    topo = topology_for(faces, n, dev)
    tpos, tfn = _dev_tensor(ini_vs, dev), _dev_tensor(fn_real, dev)
    if tpos is not None and tfn is not None and tpos.dtype != tfn.dtype:
        tfn = tfn.to(tpos.dtype)
    res = StepLossFn.apply(pos, topo, tpos, tfn, _mask_u8(v_mask, dev, n), _mask_u8(f_mask, dev, topo.nf))
    return res[0], res[1]


def sgcn_step_loss(pos: Tensor, faces, ini_vs, fn_real, v_mask, f_mask, k1: float = 4.0) -> Tensor:
    """loss = loss_p + k1 * loss_n of sgcn.py:130-137 (non-CAD branch)."""
    lp, ln = sgcn_step_losses(pos, faces, ini_vs, fn_real, v_mask, f_mask)
    return lp + k1 * ln


def fused_mask_pos_rec_loss(pred_pos: Tensor, real_pos, mask=None) -> Tensor:
    """util/loss.py:14-34 (``ltype="rmse"``) / pos_rec_loss (mask=None) on the fused kernel."""
    dev = pred_pos.device
    res = StepLossFn.apply(pred_pos, None, _dev_tensor(real_pos, dev), None, _mask_u8(mask, dev, int(pred_pos.shape[0])), None)
    return res[0]


class LapLossFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, pos: Tensor, graph):
        lib = L.load()
        require_cuda(pos)
        if pos.dtype != torch.float32 or pos.dim() != 2 or pos.shape[1] != 3:
            raise SgbError("pos must be a float32 [N, 3] tensor")
        if pos.stride(1) != 1:
            pos = pos.contiguous()
        dev, n = pos.device, int(pos.shape[0])
        diff = torch.empty((n, 3), dtype=torch.float32, device=dev)
        partials = torch.empty(lib.sgb_loss_partial_rows(), dtype=torch.float64, device=dev)
        out = torch.empty(1, dtype=torch.float64, device=dev)
        with torch.cuda.device(dev):
            check(lib.sgb_lap_loss_fwd(ptr(pos), pos.stride(0), n, ptr(graph.rowptr), ptr(graph.edges), ptr(diff), ptr(partials),
                                       ptr(out), stream_ptr(dev)), "sgb_lap_loss_fwd")
        L.count(2)
        ctx.save_for_backward(diff, out)
        ctx.graph = graph
        return out[0].to(torch.float32)

    @staticmethod
    def backward(ctx, g: Tensor):
        lib = L.load()
        diff, out = ctx.saved_tensors
        graph = ctx.graph
        dev, n = diff.device, int(diff.shape[0])
        grad = g.to(torch.float64).reshape(1).contiguous()
        dpos = torch.empty((n, 3), dtype=torch.float32, device=dev)
        with torch.cuda.device(dev):
            check(lib.sgb_lap_loss_bwd(ptr(diff), n, ptr(graph.rowptr), ptr(graph.rowptr_t), ptr(graph.edges_t), ptr(out), ptr(grad),
                                       ptr(dpos), 3, stream_ptr(dev)), "sgb_lap_loss_bwd")
        L.count(1)
        return dpos, None


def mesh_laplacian_loss(pred_pos: Tensor, edge_index: Tensor) -> Tensor:
    """util/loss.py:60-76 (``ltype="rmse"``): sqrt(mean_v |pos_v - (Adj pos)_v / deg_v|^2 + 1e-12), with the
    adjacency given as the mesh ``edge_index`` (util/mesh.py:229-230) instead of the dense-built ``mesh.Adj``."""
    from . import ops
    g = ops.graph_for(edge_index.to(pred_pos.device), int(pred_pos.shape[0]), L.MODE_ADJ)
    return LapLossFn.apply(pred_pos, g)


# ----------------------------------------------------------------------------------------
# unfused forms (torch ops; for callers that need the normals themselves)
# ----------------------------------------------------------------------------------------
def compute_fn(vs: Tensor, faces: Tensor) -> Tensor:
    """Unit face normals, util/models.py:121-126."""
    a = vs[faces[:, 0]]
    n = torch.linalg.cross(vs[faces[:, 1]] - a, vs[faces[:, 2]] - a)
    return n / torch.sqrt(torch.sum(n * n, dim=1, keepdim=True))


def mask_pos_rec_loss(pred_pos: Tensor, real_pos: Tensor, mask: Tensor) -> Tensor:
    """sqrt(mean_{i in mask} |real_i - pred_i|^2 + 1e-6), util/loss.py:23-28 (mask-multiplicative: no sync)."""
    real_pos = real_pos.to(pred_pos.device)
    m = mask.to(pred_pos.device).to(real_pos.dtype).reshape(-1, 1)
    d = (real_pos - pred_pos) ** 2
    return torch.sqrt(torch.sum(d * m) / torch.sum(m) + 1.0e-6)


def mask_norm_rec_loss(pred_norm: Tensor, real_norm: Tensor, mask: Tensor) -> Tensor:
    """mean_{f in mask} sum_d |pred - real|, util/loss.py:91-93."""
    real_norm = real_norm.to(pred_norm.device)
    m = mask.to(pred_norm.device).to(real_norm.dtype).reshape(-1, 1)
    d = torch.abs(pred_norm - real_norm)
    return torch.sum(d * m) / torch.sum(m)


_BNF_LTYPES = {"mae": 0, "l1mae": 1, "rmse": 2, "l1rmse": 3}
_F2F_CACHE: "OrderedDict[tuple, tuple]" = OrderedDict()


def _f2f_i32(f2f, device) -> Tensor:
    """``mesh.f2f`` (int64 numpy array in the reference, util/mesh.py:227) as the int32 device tensor the kernel reads, cached
    per array object (the reference passes the same mesh every step, sgcn.py:134)."""
    if isinstance(f2f, np.ndarray):
        key = ("np", id(f2f), f2f.shape, str(device))
    else:
        key = ("t", f2f.data_ptr(), tuple(f2f.shape), f2f._version, str(f2f.device), str(device))
    hit = _F2F_CACHE.get(key)
    if hit is not None:
        _F2F_CACHE.move_to_end(key)
        return hit[0]
    t = torch.from_numpy(np.ascontiguousarray(f2f)) if isinstance(f2f, np.ndarray) else f2f
    t = t.to(device).to(torch.int32).contiguous()
    if t.dim() != 2 or t.shape[1] != 3:
        raise SgbError("f2f must have shape [F, 3]")
    _F2F_CACHE[key] = (t, f2f)
    while len(_F2F_CACHE) > 8:
        _F2F_CACHE.popitem(last=False)
    return t


class BnfLossFn(torch.autograd.Function):
    """(loss, new_fn) of util/loss.py:197-253 as one autograd node on the kernels of csrc/bnf_loss.cu."""

    @staticmethod
    def forward(ctx, pos: Tensor, fn: Tensor, faces: Tensor, f2f: Tensor, ltype: int, loop: int):
        lib = L.load()
        require_cuda(pos, fn, faces, f2f)
        if fn.dtype != torch.float32 or pos.dtype != torch.float32:
            raise SgbError("fn_bnf_detach_loss: pos and fn must be float32 (the network output and its face normals)")
        pos = pos.detach()
        if pos.stride(1) != 1:
            pos = pos.contiguous()
        fn_c = fn.detach().contiguous()
        dev = fn.device
        nf, n = int(fn_c.shape[0]), int(pos.shape[0])
        if tuple(faces.shape) != (nf, 3) or tuple(f2f.shape) != (nf, 3):
            raise SgbError("fn_bnf_detach_loss: faces and f2f must be [F, 3] with F = fn.shape[0]")
        work = torch.empty(lib.sgb_bnf_work_floats(nf), dtype=torch.float32, device=dev)
        partials = torch.empty(lib.sgb_bnf_partial_rows(), dtype=torch.float64, device=dev)
        new_fn = torch.empty((nf, 3), dtype=torch.float32, device=dev)
        out = torch.empty(2, dtype=torch.float64, device=dev)
        with torch.cuda.device(dev), _prof.region("bnf_loss", 4.0 * nf * (13 + 24 * max(loop, 0))):
            check(lib.sgb_bnf_loss_fwd(ptr(pos), pos.stride(0), n, ptr(faces), ptr(f2f), nf, ptr(fn_c), int(loop), int(ltype),
                                       ptr(work), ptr(partials), ptr(new_fn), ptr(out), stream_ptr(dev)), "sgb_bnf_loss_fwd")
        L.count(5 + max(int(loop), 0))
        ctx.save_for_backward(fn_c, new_fn, out)
        ctx.ltype = int(ltype)
        ctx.mark_non_differentiable(new_fn)
        return out[0].to(torch.float32), new_fn

    @staticmethod
    def backward(ctx, g_loss, _g_new_fn):
        lib = L.load()
        fn_c, new_fn, out = ctx.saved_tensors
        dev = fn_c.device
        nf = int(fn_c.shape[0])
        grad = g_loss.detach().to(torch.float64).reshape(1).contiguous()
        dfn = torch.empty((nf, 3), dtype=torch.float32, device=dev)
        with torch.cuda.device(dev), _prof.region("bnf_loss", 4.0 * nf * 9):
            check(lib.sgb_bnf_loss_bwd(ptr(fn_c), ptr(new_fn), nf, ctx.ltype, ptr(out), ptr(grad), ptr(dfn), stream_ptr(dev)),
                  "sgb_bnf_loss_bwd")
        L.count(1)
        return None, dfn, None, None, None, None


def fn_bnf_detach_loss(pos, fn: Tensor, faces, f2f, ltype: str = "l1mae", loop: int = 5):
    """``Loss.fn_bnf_detach_loss(pos, fn, mesh, ltype, loop)`` (util/loss.py:197-253; the ``-CAD`` regulariser of the step,
    sgcn.py:133-136) on the fused kernels of csrc/bnf_loss.cu: ``faces`` / ``f2f`` are ``mesh.faces`` / ``mesh.f2f`` (numpy
    arrays as the reference holds them, or tensors; ``meshgen.face_adjacency`` builds ``f2f`` without the reference's
    ``Mesh``), -1 = no neighbour.  Returns ``(loss, new_fn)`` like the reference; the gradient reaches ``fn`` only (the
    filtered normals are detached).  CUDA tensors only -- the CPU restatement used by the tests lives in
    oracle/loss_ref.py."""
    if ltype not in _BNF_LTYPES:
        raise SgbError(f"fn_bnf_detach_loss: unknown ltype {ltype!r}")
    require_cuda(fn)
    dev = fn.device
    if isinstance(pos, np.ndarray):
        pos = torch.from_numpy(pos)
    pos = pos.to(dev)
    topo = topology_for(faces, int(pos.shape[0]), dev)           # int64 device faces, cached per faces object
    return BnfLossFn.apply(pos, fn, topo.faces, _f2f_i32(f2f, dev), _BNF_LTYPES[ltype], int(loop))
