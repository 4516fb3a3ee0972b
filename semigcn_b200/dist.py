"""Multi-GPU runtime of the vertex-partitioned mode: communicators, halo exchange, partitioned
graph operator, SyncBN / gradient reductions.  One process per GPU,
``torch.distributed`` (NCCL over NVLink/NVSwitch) for the plumbing; every compute kernel is ours.

Per propagation (SpMM forward, and SpMM backward on the transpose CSR) there is ONE exchange:
owned boundary rows are packed with ``sgb_gather_rows``, sent with a single all-to-all, and land
in a ghost block that ``sgb_spmm_halo`` addresses directly (no concatenated copy of X).  Per
BatchNorm layer the ranks all-gather their (count, mean, M2) partial rows (forward) and all-reduce
the two backward sums; parameter gradients are summed once per step (``sync_gradients``).
"""
from __future__ import annotations

from typing import List, Optional

import torch
from torch import Tensor

from . import _lib as L
from . import ops
from .partition import PartitionPlan


# ----------------------------------------------------------------------------------------
# communicators
# ----------------------------------------------------------------------------------------
class TorchComm:
    """torch.distributed backend: NCCL on the GPUs of one box (NVLink / NVSwitch).  With a ``gloo`` group
    (CPU tests of the host logic, and the 2-process single-GPU equivalence test) CUDA tensors are staged
    through host memory, because gloo's all-to-all only takes CPU tensors."""

    def __init__(self, group=None):
        import torch.distributed as dist
        self.dist, self.group = dist, group
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        self.stage = dist.get_backend(group) == "gloo"

    def _in(self, t: Tensor) -> Tensor:
        return t.detach().cpu().contiguous() if (self.stage and t.is_cuda) else t.contiguous()

    def all_to_all_rows(self, send: Tensor, send_counts: List[int], recv_counts: List[int]) -> Tensor:
        s = self._in(send)
        recv = torch.empty((sum(recv_counts),) + tuple(s.shape[1:]), dtype=s.dtype, device=s.device)
        self.dist.all_to_all_single(recv, s, output_split_sizes=list(recv_counts), input_split_sizes=list(send_counts), group=self.group)
        return recv.to(send.device)

    def all_to_all_rows_start(self, send: Tensor, send_counts: List[int], recv_counts: List[int]):
        """Asynchronous form: returns ``(recv, work)``; the collective runs on the communicator's own stream while the caller
        keeps launching on the current one, ``work.wait()`` makes the current stream wait for it (no host block).  The gloo
        staging path has no asynchronous form: it completes here and returns ``work = None``."""
        if self.stage and send.is_cuda:
            return self.all_to_all_rows(send, send_counts, recv_counts), None
        s = send.contiguous()
        recv = torch.empty((sum(recv_counts),) + tuple(s.shape[1:]), dtype=s.dtype, device=s.device)
        work = self.dist.all_to_all_single(recv, s, output_split_sizes=list(recv_counts), input_split_sizes=list(send_counts), group=self.group,
                                           async_op=True)
        return recv, work

    def all_gather_cat(self, t: Tensor) -> Tensor:
        s = self._in(t)
        out = torch.empty((self.world * s.shape[0],) + tuple(s.shape[1:]), dtype=s.dtype, device=s.device)
        self.dist.all_gather_into_tensor(out, s, group=self.group)
        return out.to(t.device)

    def _reduce(self, t: Tensor, op) -> Tensor:
        if self.stage and t.is_cuda:
            s = t.detach().cpu()
            self.dist.all_reduce(s, op=op, group=self.group)
            t.copy_(s)
        else:
            self.dist.all_reduce(t, op=op, group=self.group)
        return t

    def all_reduce_sum(self, t: Tensor) -> Tensor:
        return self._reduce(t, self.dist.ReduceOp.SUM)

    def all_reduce_min(self, t: Tensor) -> Tensor:
        return self._reduce(t, self.dist.ReduceOp.MIN)

    def all_reduce_max(self, t: Tensor) -> Tensor:
        return self._reduce(t, self.dist.ReduceOp.MAX)


# ----------------------------------------------------------------------------------------
# halo exchange + partitioned operator
# ----------------------------------------------------------------------------------------
class HaloExchange:
    def __init__(self, plan: PartitionPlan, comm):
        self.plan, self.comm = plan, comm
        self.bytes_last, self.bytes_total, self.calls = 0, 0, 0

    def exchange(self, x_own: Tensor) -> Tensor:
        """Ghost rows of ``x_own`` [n_own, C] -> [n_ghost, C] (rows ordered like ``plan.ghost_gid``)."""
        return self.finish(self.start(x_own))

    def start(self, x_own: Tensor):
        """Pack the boundary rows (one gather kernel for all destinations) and put the all-to-all in flight."""
        p = self.plan
        send = ops.gather_rows(x_own, p.send_idx)
        recv, work = self.comm.all_to_all_rows_start(send, p.send_counts, p.recv_counts)
        self.bytes_last = int(send.numel() + recv.numel()) * 4
        self.bytes_total += self.bytes_last
        self.calls += 1
        return recv, work, send, x_own

    def finish(self, pending) -> Tensor:
        recv, work, _send, x_own = pending          # `_send` stays referenced until here: the collective reads it
        if work is not None:
            work.wait()
        if recv.shape[0] == 0:       # keep a valid (aligned) pointer for the kernel
            recv = torch.zeros((1, x_own.shape[1]), dtype=x_own.dtype, device=x_own.device)
        return recv


class PartitionedGraph(ops.MeshGraph):
    """The rows of the normalised operator this rank owns.  Same interface as ``ops.MeshGraph``; ``ops.spmm``
    sees ``halo`` and fetches the ghost rows first.  Degrees / ``dis`` of owned vertices are exact locally
    (every edge with an owned endpoint is kept); ghost ``dis`` values come from their owners once, then
    the per-edge weights are recomputed (fl(dis_src * dis_dst), identical rounding to the builder)."""

    def __init__(self, plan: PartitionPlan, comm, mode: int, overlap: bool = True):
        super().__init__(plan.edge_index, plan.n_local, mode)
        n_own = plan.n_own
        self.plan = plan
        halo = HaloExchange(plan, comm)
        dis_own = self.dis[:n_own].contiguous()
        dis_ghost = halo.exchange(dis_own.unsqueeze(1)).reshape(-1)[:plan.n_ghost]
        dis_full = torch.cat([dis_own, dis_ghost])
        for rowptr, colidx, edges in ((self.rowptr, self.colidx, self.edges), (self.rowptr_t, self.colidx_t, self.edges_t)):
            counts = (rowptr[1:] - rowptr[:-1]).long()
            rowid = torch.repeat_interleave(torch.arange(plan.n_local, device=rowptr.device), counts)
            m = int(rowid.numel())
            if mode == L.MODE_ADJ or m == 0:
                continue
            w = dis_full[colidx[:m].long()] * dis_full[rowid]
            if mode == L.MODE_CHEB:
                w = -w
            edges[:m, 1] = w.view(torch.int32)
        self.dis = dis_own
        self.n = n_own                      # rows this rank computes
        self.n_local = plan.n_local
        self.halo, self.comm, self.n_global = halo, comm, plan.n_global
        # interior-first numbering (partition.interior_first_order): rows [0, n_interior) touch no ghost -> ops.spmm runs them
        # while the halo exchange is in flight.  0 = no overlap (any other numbering: a boundary row may come first).
        from .partition import count_interior
        self.n_interior = count_interior(plan.edge_index, n_own) if overlap else 0


def register_partition(plan: PartitionPlan, comm, modes=(L.MODE_GCN, L.MODE_CHEB), overlap: bool = True) -> Tensor:
    """Make the drop-in convs use the partitioned operator: returns the LOCAL ``edge_index`` tensor to pass to
    ``forward(x_own, edge_index)``; ``ops.graph_for`` resolves it (by identity) to a ``PartitionedGraph``."""
    ei = plan.edge_index
    for mode in modes:
        g = PartitionedGraph(plan, comm, mode, overlap=overlap)
        key = (ei.data_ptr(), tuple(ei.shape), ei._version, str(ei.device), mode, plan.n_own)
        ops._GRAPH_CACHE[key] = (g, ei, 0)
    ops._GRAPH_CACHE_PINNED.update(k for k in ops._GRAPH_CACHE if k[0] == ei.data_ptr())
    return ei


def sync_gradients(module: torch.nn.Module, comm) -> None:
    """Sum the per-rank parameter-gradient contributions (each rank back-propagated its own vertices)."""
    grads = [p.grad for p in module.parameters() if p.grad is not None]
    if not grads:
        return
    flat = torch.cat([g.reshape(-1) for g in grads])
    comm.all_reduce_sum(flat)
    off = 0
    for g in grads:
        g.copy_(flat[off:off + g.numel()].view_as(g))
        off += g.numel()


class DistPosLossFn(torch.autograd.Function):
    """mask_pos_rec_loss (util/loss.py:14-34, "rmse") over a vertex-partitioned mesh: local masked sums,
    one all-reduce of (sum, count), local gradient."""

    @staticmethod
    def forward(ctx, pos_own: Tensor, target_own: Tensor, mask_own: Tensor, comm):
        m = mask_own.to(target_own.dtype).reshape(-1, 1)
        d = (pos_own.to(target_own.dtype) - target_own) * m
        acc = torch.stack([(d * d).sum(), m.sum()]).to(torch.float64)
        comm.all_reduce_sum(acc)
        loss = torch.sqrt(acc[0] / acc[1] + 1.0e-6)
        ctx.save_for_backward(d, acc, loss)
        return loss.to(target_own.dtype)

    @staticmethod
    def backward(ctx, g):
        d, acc, loss = ctx.saved_tensors
        return (g.to(torch.float64) * d.to(torch.float64) / (acc[1] * loss)).to(torch.float32), None, None, None


def dist_mask_pos_rec_loss(pos_own: Tensor, target_own: Tensor, mask_own: Tensor, comm) -> Tensor:
    return DistPosLossFn.apply(pos_own, target_own, mask_own, comm)


# ----------------------------------------------------------------------------------------
# face-normal loss across partition cuts (sgcn.py:130-137 on a vertex-partitioned mesh)
# ----------------------------------------------------------------------------------------
class HaloRowsFn(torch.autograd.Function):
    """ghost rows of ``x_own`` with a gradient: forward = the halo exchange, backward = the reverse exchange (every ghost
    row's gradient goes back to its owner and is added to the owner's row; one ``index_add_`` per source rank with unique
    indices each, in rank order -> deterministic)."""

    @staticmethod
    def forward(ctx, x_own: Tensor, halo: "HaloExchange"):
        ctx.halo, ctx.n_own = halo, x_own.shape[0]
        p = halo.plan
        send = x_own.detach()[p.send_idx.long()].contiguous()
        recv = halo.comm.all_to_all_rows(send, p.send_counts, p.recv_counts)
        return recv

    @staticmethod
    def backward(ctx, d_recv: Tensor):
        halo, p = ctx.halo, ctx.halo.plan
        back = halo.comm.all_to_all_rows(d_recv.contiguous(), p.recv_counts, p.send_counts)     # grouped like send_idx
        dx = torch.zeros((ctx.n_own,) + tuple(d_recv.shape[1:]), dtype=d_recv.dtype, device=d_recv.device)
        off = 0
        idx = p.send_idx.long()
        for cnt in p.send_counts:
            if cnt:
                dx.index_add_(0, idx[off:off + cnt], back[off:off + cnt])
                off += cnt
        return dx, None


class AllReduceSumFn(torch.autograd.Function):
    """sum over ranks of a local partial sum; every rank's partial enters the global value with weight 1."""

    @staticmethod
    def forward(ctx, t: Tensor, comm):
        return comm.all_reduce_sum(t.detach().clone())

    @staticmethod
    def backward(ctx, g: Tensor):
        return g, None


def dist_mask_norm_rec_loss(pos_own: Tensor, halo: "HaloExchange", faces_local: Tensor, target_fn: Tensor, fmask: Tensor, comm) -> Tensor:
    """mask_norm_rec_loss(compute_fn(pos), fn, f_mask) (util/models.py:121-126, util/loss.py:78-107 "l1mae") over a
    vertex-partitioned mesh.  ``faces_local`` / ``target_fn`` / ``fmask``: the faces this rank accounts for
    (``partition.local_faces``) in local vertex ids.  Ghost positions arrive through the halo exchange and their gradients
    travel back through it; the masked L1 sum and the face count are all-reduced."""
    ghosts = HaloRowsFn.apply(pos_own, halo)
    pos = torch.cat([pos_own, ghosts.to(pos_own.dtype)], dim=0) if ghosts.shape[0] else pos_own
    a = pos[faces_local[:, 0]]
    n = torch.linalg.cross(pos[faces_local[:, 1]] - a, pos[faces_local[:, 2]] - a)
    n = n / torch.sqrt(torch.sum(n * n, dim=1, keepdim=True))
    m = fmask.to(target_fn.dtype).reshape(-1, 1)
    local = torch.sum(torch.abs(n.to(target_fn.dtype) - target_fn) * m).reshape(1)
    count = comm.all_reduce_sum(m.sum().reshape(1).to(torch.float64).clone())
    total = AllReduceSumFn.apply(local.to(torch.float64), comm)
    return (total / count).reshape(()).to(target_fn.dtype)
