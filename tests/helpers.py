"""Shared test utilities.  Tolerances follow SURVEY.md §8(d): layer outputs and gradients
max|a-b| / max|b| <= 1e-5 (norm-relative); builder outputs and aggregation bit-exact."""
import os

import numpy as np
import torch

REL_TOL = 1e-5
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def rel_err(a: torch.Tensor, b: torch.Tensor) -> float:
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    denom = b.abs().max().item()
    if denom == 0.0:
        return a.abs().max().item()
    return (a - b).abs().max().item() / denom


def assert_close(a, b, tol=REL_TOL, what=""):
    e = rel_err(a, b)
    assert e <= tol, f"{what}: norm-relative error {e:.3e} > {tol:.1e}"


def bits(t: torch.Tensor) -> torch.Tensor:
    return t.detach().cpu().contiguous().view(torch.int32)


def assert_bit_equal(a, b, what=""):
    a, b = a.detach().cpu(), b.detach().cpu()
    assert a.shape == b.shape, f"{what}: shape {a.shape} vs {b.shape}"
    if a.dtype.is_floating_point:
        same = bits(a) == bits(b)
        # +0 / -0 are the same value
        same |= (a == 0) & (b == 0)
    else:
        same = a == b
    nbad = int((~same).sum())
    assert nbad == 0, f"{what}: {nbad} of {a.numel()} elements differ bitwise (max abs diff {(a.double()-b.double()).abs().max().item():.3e})"


def load_golden(n):
    """``n``: icosphere frequency of a reference-mesh fixture (ref_mesh_n{n}.npz) or a fixture file name."""
    name = n if isinstance(n, str) else f"ref_mesh_n{n}.npz"
    return np.load(os.path.join(GOLDEN, name))


def random_graph(n: int, nnz: int, seed: int, self_loops: int = 0, duplicates: int = 0, isolated: int = 0,
                 symmetric: bool = False) -> torch.Tensor:
    """Random directed edge list with optional self loops, duplicate edges and isolated vertices."""
    g = torch.Generator().manual_seed(seed)
    live = n - isolated
    row = torch.randint(0, live, (nnz,), generator=g)
    col = torch.randint(0, live, (nnz,), generator=g)
    keep = row != col
    row, col = row[keep], col[keep]
    if symmetric:
        row, col = torch.cat([row, col]), torch.cat([col, row])
    if duplicates:
        idx = torch.randint(0, row.numel(), (duplicates,), generator=g)
        row, col = torch.cat([row, row[idx]]), torch.cat([col, col[idx]])
    if self_loops:
        v = torch.randint(0, live, (self_loops,), generator=g)
        row, col = torch.cat([row, v]), torch.cat([col, v])
    perm = torch.randperm(row.numel(), generator=g)
    return torch.stack([row[perm], col[perm]]).contiguous()


def stable_csr_from_edges(edge_index: torch.Tensor, n: int, by_source: bool = False):
    """Reference CSR: group non-loop edges by target (or source), stable in edge order."""
    row, col = edge_index[0], edge_index[1]
    keep = row != col
    eid = torch.nonzero(keep).flatten()
    key = (row if by_source else col)[eid]
    other = (col if by_source else row)[eid]
    order = torch.sort(key, stable=True)[1]
    counts = torch.bincount(key, minlength=n)
    rowptr = torch.zeros(n + 1, dtype=torch.int64)
    rowptr[1:] = torch.cumsum(counts, 0)
    perm = torch.full((edge_index.shape[1],), -1, dtype=torch.int64)
    perm[eid[order]] = torch.arange(order.numel())
    return rowptr.to(torch.int32), other[order].to(torch.int32), perm.to(torch.int32)
