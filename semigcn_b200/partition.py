"""Vertex partition of one large mesh graph over P ranks (BASELINE.json configs[3]).

The reference is single-device (sgcn.py:77) and its ``Mesh`` cannot even hold such a mesh
(dense N x N, util/mesh.py:267), so there is no reference counterpart: this is the scaling layer
around the drop-in convs.  Pure index arithmetic on torch tensors (CPU or CUDA), no communication:

  * ranks own contiguous vertex ranges of the given numbering (mesh generators / a prior BFS or
    space-filling-curve renumbering give the locality; cut quality = how few edges cross ranges);
  * a rank keeps every directed edge with an owned endpoint; non-owned endpoints become ghost
    vertices, numbered after the owned ones in ascending global id (= grouped by owner);
  * ``send_idx`` lists, per destination rank, the owned vertices that rank needs as ghosts; by
    construction rank q's ghost block from rank r has the same order as r's send block to q.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import List, Optional, Sequence, Tuple

import torch
from torch import Tensor


def vertex_ranges(n: int, world: int) -> List[Tuple[int, int]]:
    """Balanced contiguous ranges [lo, hi) (first n % world ranks get one extra vertex)."""
    base, extra = divmod(int(n), int(world))
    out, lo = [], 0
    for r in range(world):
        hi = lo + base + (1 if r < extra else 0)
        out.append((lo, hi))
        lo = hi
    return out


def morton_order(vs: Tensor, bits: int = 16) -> Tensor:
    """Permutation ``perm`` (new id -> old id) that sorts the vertices along a Z-order (Morton) curve of their positions
    quantised to ``bits`` bits per axis.  Contiguous ranges of the renumbered vertices are compact patches of the surface,
    so the cut of ``build_plan`` is short AND balanced -- a generator's own numbering can put every "seam" vertex into
    the first range (the icosphere's edge vertices: rank 0 held half of all ghost rows at N = 8)."""
    v = vs.detach().to(torch.float64)
    lo, hi = v.min(dim=0)[0], v.max(dim=0)[0]
    scale = torch.where(hi > lo, (2 ** bits - 1) / (hi - lo), torch.zeros_like(hi))
    q = ((v - lo) * scale).round().to(torch.int64).clamp_(0, 2 ** bits - 1)
    code = torch.zeros(v.shape[0], dtype=torch.int64, device=vs.device)
    for b in range(bits):                    # interleave x, y, z bits: 3 * bits <= 63
        for a in range(3):
            code |= ((q[:, a] >> b) & 1) << (3 * b + a)
    return torch.argsort(code, stable=True)


def renumber(perm: Tensor, edge_index: Tensor, *vertex_tensors: Tensor):
    """Apply ``perm`` (new id -> old id): returns the renumbered ``edge_index`` (same edge order) and the vertex tensors
    gathered into the new order.  ``inverse = argsort(perm)`` maps results back (``out_old = out_new[inverse]``)."""
    inv = torch.empty_like(perm)
    inv[perm] = torch.arange(perm.numel(), device=perm.device)
    return (inv[edge_index],) + tuple(t[perm] for t in vertex_tensors)


@dataclass
class PartitionPlan:
    rank: int
    world: int
    n_global: int
    lo: int
    hi: int
    ghost_gid: Tensor            # int64 [n_ghost], ascending global ids (grouped by owner rank)
    send_idx: Tensor             # int32 [n_send], LOCAL owned ids grouped by destination rank, ascending inside a group
    send_counts: List[int]       # rows sent to each rank
    recv_counts: List[int]       # ghost rows received from each rank
    edge_index: Tensor           # int64 [2, m] local ids: owned in [0, n_own), ghosts in [n_own, n_own + n_ghost)

    @property
    def n_own(self) -> int:
        return self.hi - self.lo

    @property
    def n_ghost(self) -> int:
        return int(self.ghost_gid.numel())

    @property
    def n_local(self) -> int:
        return self.n_own + self.n_ghost

    def to(self, device) -> "PartitionPlan":
        return PartitionPlan(self.rank, self.world, self.n_global, self.lo, self.hi, self.ghost_gid.to(device),
                             self.send_idx.to(device), list(self.send_counts), list(self.recv_counts), self.edge_index.to(device))


def build_plan(edge_index: Tensor, n: int, rank: int, world: int,
               ranges: Optional[Sequence[Tuple[int, int]]] = None) -> PartitionPlan:
    """Plan of ``rank`` from the GLOBAL ``edge_index`` [2, nnz] (every rank holds it in the synthetic benchmarks;
    a loader that only sees its own edges would exchange the ghost lists instead)."""
    ranges = list(ranges) if ranges is not None else vertex_ranges(n, world)
    lo, hi = ranges[rank]
    row, col = edge_index[0], edge_index[1]
    own_r = (row >= lo) & (row < hi)
    own_c = (col >= lo) & (col < hi)
    keep = own_r | own_c
    r_k, c_k = row[keep], col[keep]
    ghosts = torch.unique(torch.cat([r_k[~own_r[keep]], c_k[~own_c[keep]]]))          # sorted ascending
    n_own = hi - lo

    def to_local(v: Tensor) -> Tensor:
        owned = (v >= lo) & (v < hi)
        g_pos = torch.searchsorted(ghosts, v.clamp(min=0))
        g_pos = g_pos.clamp(max=max(int(ghosts.numel()) - 1, 0))
        return torch.where(owned, v - lo, n_own + g_pos)

    ei_local = torch.stack([to_local(r_k), to_local(c_k)]).contiguous()
    bounds = torch.tensor([b for (b, _) in ranges] + [ranges[-1][1]], dtype=torch.int64, device=edge_index.device)
    owner_of_ghost = torch.searchsorted(bounds, ghosts, right=True) - 1
    recv_counts = torch.bincount(owner_of_ghost, minlength=world).tolist() if ghosts.numel() else [0] * world
    send_parts, send_counts = [], []
    for q in range(world):
        if q == rank:
            send_counts.append(0)
            continue
        qlo, qhi = ranges[q]
        q_r = (row >= qlo) & (row < qhi)
        q_c = (col >= qlo) & (col < qhi)
        mine = torch.unique(torch.cat([row[own_r & q_c], col[own_c & q_r]]))           # owned vertices adjacent to q's
        send_parts.append(mine - lo)
        send_counts.append(int(mine.numel()))
    send_idx = (torch.cat(send_parts) if send_parts else torch.zeros(0, dtype=torch.int64, device=edge_index.device)).to(torch.int32)
    return PartitionPlan(rank, world, int(n), lo, hi, ghosts, send_idx, send_counts, [int(c) for c in recv_counts], ei_local)


def interior_first_order(edge_index: Tensor, n: int, ranges: Sequence[Tuple[int, int]]) -> Tensor:
    """Permutation ``perm`` (new id -> old id) that keeps every rank's vertex RANGE but lists, inside it, the interior
    vertices (all neighbours owned by the same rank) before the boundary ones (at least one neighbour on another rank),
    each group in its previous order.  Rows [0, n_interior) of a rank's operator then have no ghost columns: their
    aggregation can run while the halo exchange of the boundary rows is still in flight (dist.PartitionedGraph)."""
    dev = edge_index.device
    bounds = torch.tensor([lo for (lo, _) in ranges] + [ranges[-1][1]], dtype=torch.int64, device=dev)
    owner = torch.searchsorted(bounds, torch.arange(n, device=dev), right=True) - 1
    row, col = edge_index[0], edge_index[1]
    cross = owner[row] != owner[col]
    boundary = torch.zeros(n, dtype=torch.bool, device=dev)
    boundary[row[cross]] = True
    boundary[col[cross]] = True
    key = owner * 2 + boundary.to(torch.int64)              # (rank, interior < boundary), stable inside
    return torch.argsort(key, stable=True)


def count_interior(edge_index_local: Tensor, n_own: int) -> int:
    """Number of leading owned rows of a plan's local ``edge_index`` that touch no ghost (valid as a split point when the
    numbering came from ``interior_first_order``: the first boundary vertex ends the interior block)."""
    src, dst = edge_index_local[0], edge_index_local[1]
    touched = torch.zeros(n_own + 1, dtype=torch.bool, device=src.device)
    ghost_src = src >= n_own
    touched[dst[ghost_src].clamp(max=n_own)] = True          # owned target with a ghost source
    ghost_dst = dst >= n_own
    touched[src[ghost_dst].clamp(max=n_own)] = True          # owned source with a ghost target (transpose operator)
    first = torch.nonzero(touched[:n_own])
    return int(first[0]) if first.numel() else n_own


def local_faces(plan: "PartitionPlan", faces: Tensor) -> Tuple[Tensor, Tensor]:
    """Faces this rank accounts for in a face-wise loss: those whose FIRST vertex it owns (every face is counted by exactly
    one rank).  The other two vertices are neighbours of the first, hence owned or in the rank's ghost set.  Returns
    (global face ids [F_r], faces in LOCAL vertex ids [F_r, 3])."""
    lo, hi, n_own = plan.lo, plan.hi, plan.n_own
    mine = (faces[:, 0] >= lo) & (faces[:, 0] < hi)
    fid = torch.nonzero(mine).reshape(-1)
    f = faces[fid]
    owned = (f >= lo) & (f < hi)
    gpos = torch.searchsorted(plan.ghost_gid, f.reshape(-1).clamp(min=0)).reshape(f.shape)
    gpos = gpos.clamp(max=max(plan.n_ghost - 1, 0))
    if plan.n_ghost:
        ok = owned | (plan.ghost_gid[gpos] == f)
        if not bool(ok.all()):
            raise ValueError("local_faces: a face vertex is neither owned nor a ghost of this rank")
    return fid, torch.where(owned, f - lo, n_own + gpos).contiguous()
