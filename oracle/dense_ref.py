"""ORACLE B -- TEST INFRASTRUCTURE ONLY (independent cross-check of oracle/pyg_ref.py).

Dense-matrix formulation of the same operators, written from the textbook definitions
rather than from the gather/scatter op sequence (SURVEY.md §4 "Oracle B"):

  GCNConv :  out = D~^-1/2 (A + I) D~^-1/2 (X W^T) + b ,  D~ = in-degree of (A+I)
  ChebConv:  out = sum_k T_k(L^) X W_k^T + b ,  L^ = -D^-1/2 A D^-1/2  (lambda_max = 2),
             T_0 = X, T_1 = L^ X, T_k = 2 L^ T_{k-1} - T_{k-2}

``A[i, j]`` = number of non-loop edges j -> i in ``edge_index`` (duplicates add).  Small N only.
"""
from __future__ import annotations

from typing import Sequence

import torch
from torch import Tensor


def dense_adjacency(edge_index: Tensor, n: int, dtype=torch.float64) -> Tensor:
    a = torch.zeros(n, n, dtype=dtype)
    row, col = edge_index[0], edge_index[1]
    keep = row != col
    a.index_put_((col[keep], row[keep]), torch.ones(int(keep.sum()), dtype=dtype), accumulate=True)
    return a                                     # a[target, source]


def gcn_operator(edge_index: Tensor, n: int, dtype=torch.float64) -> Tensor:
    a = dense_adjacency(edge_index, n, dtype) + torch.eye(n, dtype=dtype)
    deg = a.sum(dim=1)                           # in-degree incl. the self loop
    dis = torch.where(deg > 0, deg.rsqrt(), torch.zeros_like(deg))
    return dis[:, None] * a * dis[None, :]


def cheb_operator(edge_index: Tensor, n: int, dtype=torch.float64) -> Tensor:
    a = dense_adjacency(edge_index, n, dtype)
    deg = a.sum(dim=0)                           # PyG's get_laplacian sums over `row` = source
    dis = torch.where(deg > 0, deg.rsqrt(), torch.zeros_like(deg))
    return -(dis[:, None] * a * dis[None, :])


def gcn_conv(x: Tensor, edge_index: Tensor, weight: Tensor, bias: Tensor | None) -> Tensor:
    op = gcn_operator(edge_index, x.shape[0], x.dtype)
    out = op @ (x @ weight.t())
    return out if bias is None else out + bias


def cheb_conv(x: Tensor, edge_index: Tensor, weights: Sequence[Tensor], bias: Tensor | None) -> Tensor:
    op = cheb_operator(edge_index, x.shape[0], x.dtype)
    t0 = x
    out = t0 @ weights[0].t()
    if len(weights) > 1:
        t1 = op @ x
        out = out + t1 @ weights[1].t()
        for w in weights[2:]:
            t2 = 2.0 * (op @ t1) - t0
            out = out + t2 @ w.t()
            t0, t1 = t1, t2
    return out if bias is None else out + bias
