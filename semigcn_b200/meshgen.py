"""Synthetic closed triangle meshes for tests and benchmarks.

The reference ships no data (``datasets.zip`` is absent) and its ``Mesh``
class densifies N x N matrices (``util/mesh.py:267``), so every mesh above a few thousand
vertices is produced here instead: a geodesic icosphere of frequency ``n`` with
``V = 10 n^2 + 2``, ``E = 30 n^2``, ``F = 20 n^2``.

``edge_index`` follows the reference's construction contract exactly
(``util/mesh.py:69-87`` + ``:229-230``, ``util/datamaker.py:76-77``): unique undirected
edges ``(min, max)`` in first-seen order while walking faces in file order and each
face's sides ``(f0,f1), (f1,f2), (f2,f0)``; then ``[edges.T || edges.T flipped]``.

All functions are torch ops so the 1 M / 16 M vertex meshes can be generated on the GPU
(input synthesis only -- not part of the measured path).
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import Optional

import torch

_PHI = (1.0 + math.sqrt(5.0)) / 2.0

# 12 icosahedron corners and 20 outward-oriented (counter-clockwise) faces.
_ICO_VERTS = [
    (-1, _PHI, 0), (1, _PHI, 0), (-1, -_PHI, 0), (1, -_PHI, 0),
    (0, -1, _PHI), (0, 1, _PHI), (0, -1, -_PHI), (0, 1, -_PHI),
    (_PHI, 0, -1), (_PHI, 0, 1), (-_PHI, 0, -1), (-_PHI, 0, 1),
]
_ICO_FACES = [
    (0, 11, 5), (0, 5, 1), (0, 1, 7), (0, 7, 10), (0, 10, 11),
    (1, 5, 9), (5, 11, 4), (11, 10, 2), (10, 7, 6), (7, 1, 8),
    (3, 9, 4), (3, 4, 2), (3, 2, 6), (3, 6, 8), (3, 8, 9),
    (4, 9, 5), (2, 4, 11), (6, 2, 10), (8, 6, 7), (9, 8, 1),
]


@dataclass
class SynthMesh:
    """A closed manifold triangle mesh in the tensors the hot path consumes."""
    vs: torch.Tensor          # [V, 3] float32 (or float64) vertex positions
    faces: torch.Tensor       # [F, 3] int64
    edges: torch.Tensor       # [E, 2] int64, row < col, first-seen order
    edge_index: torch.Tensor  # [2, 2E] int64 = [edges.T || flipped]

    @property
    def num_vertices(self) -> int:
        return int(self.vs.shape[0])

    @property
    def nnz(self) -> int:
        return int(self.edge_index.shape[1])


def icosphere_counts(n: int):
    return 10 * n * n + 2, 30 * n * n, 20 * n * n


def edges_from_faces(faces: torch.Tensor, num_vertices: Optional[int] = None) -> torch.Tensor:
    """Unique undirected edges in the reference's first-seen order (util/mesh.py:69-87).

    Vectorised restatement of the ``edge2key`` dictionary walk: half-edge ``3*f + s`` for
    side ``s`` of face ``f``; the first occurrence of each sorted pair keeps its place.
    """
    f = faces.to(torch.int64)
    nv = int(num_vertices) if num_vertices is not None else int(f.max().item()) + 1
    a = f.reshape(-1)                                   # side s starts at f[s]
    b = f[:, [1, 2, 0]].reshape(-1)                     # ... and ends at f[(s+1)%3]
    lo, hi = torch.minimum(a, b), torch.maximum(a, b)
    key = lo * nv + hi
    # stable sort by key -> the first element of each run is the first-seen half-edge
    skey, order = torch.sort(key, stable=True)
    first = torch.ones_like(skey, dtype=torch.bool)
    first[1:] = skey[1:] != skey[:-1]
    first_pos = order[first]                            # half-edge index of the first sighting
    first_pos, _ = torch.sort(first_pos)                # back to first-seen order
    return torch.stack([lo[first_pos], hi[first_pos]], dim=1)


def edge_index_from_edges(edges: torch.Tensor) -> torch.Tensor:
    """``[edges.T || edges.T[[1,0]]]`` -- util/mesh.py:229-230, util/datamaker.py:76-77."""
    e = edges.t().contiguous()
    return torch.cat([e, e[[1, 0], :]], dim=1).contiguous()


def icosphere(n: int, device="cpu", dtype=torch.float32, radius: float = 1.0) -> SynthMesh:
    """Geodesic icosphere of frequency ``n`` (each icosahedron side split into n segments).

    Vertex numbering: 12 corners, then the (n-1) interior points of each of the 30
    icosahedron edges, then the interior points of each of the 20 faces row by row
    (a 2-D-grid-like order, so neighbouring vertices are close in memory).
    """
    assert n >= 1
    dev = torch.device(device)
    corners = torch.tensor(_ICO_VERTS, dtype=torch.float64, device=dev)
    # the 30 icosahedron edges, numbered in first-seen order over the face list
    ekey = {}
    for (a, b, c) in _ICO_FACES:
        for (p, q) in ((a, b), (b, c), (c, a)):
            k = (min(p, q), max(p, q))
            if k not in ekey:
                ekey[k] = len(ekey)
    assert len(ekey) == 30
    nv, _, nf = icosphere_counts(n)
    n_edge_in = n - 1
    n_face_in = (n - 1) * (n - 2) // 2
    base_edge = 12
    base_face = 12 + 30 * n_edge_in

    vs = torch.empty((nv, 3), dtype=torch.float64, device=dev)
    faces = torch.empty((nf, 3), dtype=torch.int64, device=dev)

    ii, jj = torch.meshgrid(torch.arange(n + 1, device=dev), torch.arange(n + 1, device=dev), indexing="ij")
    valid = (ii + jj) <= n
    kk = n - ii - jj

    def edge_ids(p, q, t_from_p):
        # id of the point at distance t (1..n-1) from corner p along icosahedron edge (p,q)
        e = ekey[(min(p, q), max(p, q))]
        t = t_from_p if p < q else n - t_from_p
        return base_edge + e * n_edge_in + (t - 1)

    for fidx, (a, b, c) in enumerate(_ICO_FACES):
        # local grid id L[i, j]: i steps towards b, j steps towards c, k = n-i-j towards a
        L = torch.full((n + 1, n + 1), -1, dtype=torch.int64, device=dev)
        inter = (ii >= 1) & (jj >= 1) & (kk >= 1)
        off_j = (jj - 1) * (n - 1) - ((jj - 1) * jj) // 2
        L[inter] = (base_face + fidx * n_face_in + off_j + (ii - 1))[inter]
        if n > 1:
            m = (jj == 0) & (ii >= 1) & (ii <= n - 1)            # side a-b
            L[m] = edge_ids(a, b, ii)[m]
            m = (ii == 0) & (jj >= 1) & (jj <= n - 1)            # side a-c
            L[m] = edge_ids(a, c, jj)[m]
            m = (kk == 0) & (jj >= 1) & (jj <= n - 1)            # side b-c
            L[m] = edge_ids(b, c, jj)[m]
        L[0, 0] = a
        L[n, 0] = b
        L[0, n] = c
        # positions
        w = torch.stack([kk, ii, jj], dim=-1).to(torch.float64) / n     # barycentric (a, b, c)
        p = w[..., 0:1] * corners[a] + w[..., 1:2] * corners[b] + w[..., 2:3] * corners[c]
        p = p / torch.linalg.norm(p, dim=-1, keepdim=True) * radius
        vs[L[valid]] = p[valid]
        # triangles: "up" (i,j),(i+1,j),(i,j+1) and "down" (i+1,j),(i+1,j+1),(i,j+1)
        up = (ii + jj) <= n - 1
        iu, ju = ii[up], jj[up]
        f_up = torch.stack([L[iu, ju], L[iu + 1, ju], L[iu, ju + 1]], dim=1)
        dn = (ii + jj) <= n - 2
        idn, jdn = ii[dn], jj[dn]
        f_dn = torch.stack([L[idn + 1, jdn], L[idn + 1, jdn + 1], L[idn, jdn + 1]], dim=1)
        fl = torch.cat([f_up, f_dn], dim=0)
        faces[fidx * n * n:(fidx + 1) * n * n] = fl

    edges = edges_from_faces(faces, nv)
    return SynthMesh(vs=vs.to(dtype), faces=faces, edges=edges, edge_index=edge_index_from_edges(edges))


def uniform_laplacian_smooth(vs: torch.Tensor, edge_index: torch.Tensor, iters: int = 30) -> torch.Tensor:
    """``iters`` rounds of v <- mean(neighbours): the stand-in for the 30x Laplacian smoothing
    of ``preprocess/prepare.py:12,112-113`` (pymeshlab is absent)."""
    row, col = edge_index[0], edge_index[1]
    n = vs.shape[0]
    deg = torch.zeros(n, dtype=vs.dtype, device=vs.device).index_add_(0, col, torch.ones_like(col, dtype=vs.dtype))
    out = vs
    for _ in range(iters):
        acc = torch.zeros_like(out).index_add_(0, col, out[row])
        out = acc / deg.clamp_min(1).unsqueeze(1)
    return out


def synth_inpainting_problem(n: int, device="cpu", seed: int = 314, noise: float = 0.02,
                             smooth_iters: int = 30, hole_frac: float = 0.03, n_dummy: int = 40,
                             dummy_p: float = 0.014, dummy_k: int = 4):
    """The synthetic stand-in for one SeMIGCN data directory:

    initial = sphere + radial noise, smooth = Laplacian-smoothed initial, z1 = ini - smo,
    a real-hole vertex mask (geodesic caps, ~hole_frac of the vertices, 0 = hole) and
    ``n_dummy`` fake-hole masks grown by k-ring dilation (util/datamaker.py:110-136).
    """
    g = torch.Generator(device="cpu").manual_seed(seed)
    mesh = icosphere(n, device=device, dtype=torch.float64)
    nv = mesh.num_vertices
    bump = torch.randn(nv, 1, generator=g, dtype=torch.float64).to(mesh.vs.device)
    ini = mesh.vs * (1.0 + noise * bump)
    smo = uniform_laplacian_smooth(ini, mesh.edge_index, smooth_iters)
    # real hole: a few spherical caps
    n_caps = 3
    centers = torch.randn(n_caps, 3, generator=g, dtype=torch.float64).to(ini.device)
    centers = centers / torch.linalg.norm(centers, dim=1, keepdim=True)
    cos_thr = 1.0 - 2.0 * hole_frac / n_caps             # cap area fraction = (1-cos)/2
    v_mask = torch.ones(nv, dtype=torch.bool, device=ini.device)
    unit = mesh.vs / torch.linalg.norm(mesh.vs, dim=1, keepdim=True)
    for cidx in range(n_caps):
        v_mask &= (unit @ centers[cidx]) < cos_thr
    # dummy masks: Binomial seeds dilated k rings (AdjI @ M > 0)
    row, col = mesh.edge_index[0], mesh.edge_index[1]
    seeds = (torch.rand(nv, n_dummy, generator=g) < dummy_p).to(ini.device).to(torch.float32)
    m = seeds
    for _ in range(dummy_k):
        m = ((torch.zeros_like(m).index_add_(0, col, m[row]) + m) > 0).to(torch.float32)
    vmask_dummy = 1.0 - m
    f_mask = v_mask[mesh.faces].all(dim=1)
    fn = face_normals(ini, mesh.faces)
    return {
        "mesh": mesh, "ini_vs": ini, "smo_vs": smo, "z1": (ini - smo).to(torch.float32),
        "x_pos": smo.to(torch.float32), "v_mask": v_mask, "f_mask": f_mask,
        "vmask_dummy": vmask_dummy, "fn": fn,
    }


def face_normals(vs: torch.Tensor, faces: torch.Tensor) -> torch.Tensor:
    a, b, c = vs[faces[:, 0]], vs[faces[:, 1]], vs[faces[:, 2]]
    nrm = torch.linalg.cross(b - a, c - a)
    return nrm / torch.linalg.norm(nrm, dim=1, keepdim=True)


def write_obj(path: str, vs: torch.Tensor, faces: torch.Tensor) -> None:
    """Plain ``v``/``f`` OBJ the reference's parser accepts (util/mesh.py:35-58)."""
    v = vs.detach().cpu().double().numpy()
    f = faces.detach().cpu().numpy() + 1
    with open(path, "w") as fh:
        for p in v:
            fh.write("v %.17g %.17g %.17g\n" % (p[0], p[1], p[2]))
        for t in f:
            fh.write("f %d %d %d\n" % (t[0], t[1], t[2]))


# ----------------------------------------------------------------------------------------
# synthetic pooling hierarchy (stand-in for Mesh.simplification + pool_hash, util/meshnet.py:169-193)
# ----------------------------------------------------------------------------------------
def synth_pool_level(edge_index: torch.Tensor, n: int, ratio: float = 0.6, seed: int = 314, rounds: int = 12):
    """One coarsening step of ``n`` vertices to ``int(n * ratio)`` clusters by contracting a matching of edges
    (the reference contracts edges by QEM cost, util/mesh.py simplification -- out of scope; any matching exercises
    the same MeshPool / MeshUnpool / coarse-graph code).  Returns ``(cluster_of[n] int64, n_coarse, coarse_edge_index)``;
    ``coarse_edge_index`` follows the Mesh contract ``[unique (lo, hi) pairs || flipped]``.

    Handshake matching, vectorised: every unmatched vertex proposes to its unmatched neighbour of highest random
    priority; mutual proposals match.  Surplus pairs (beyond n - n_coarse) are dropped deterministically."""
    dev = edge_index.device
    n_coarse = int(n * ratio)
    need = n - n_coarse
    g = torch.Generator(device="cpu").manual_seed(seed)
    prio = torch.rand(n, generator=g).to(dev)
    row, col = edge_index[0], edge_index[1]
    mate = torch.full((n,), -1, dtype=torch.int64, device=dev)
    for _ in range(rounds):
        free = mate < 0
        ok = free[row] & free[col]
        if not bool(ok.any()):
            break
        r, c = row[ok], col[ok]
        # best neighbour per vertex: scatter-max on (priority of neighbour), ties broken by id through the key
        key = (prio[c] * (2 ** 20)).to(torch.int64) * n + c
        best = torch.full((n,), -1, dtype=torch.int64, device=dev)
        best.scatter_reduce_(0, r, key, reduce="amax", include_self=True)
        prop = torch.where(best >= 0, best % n, torch.full_like(best, -1))
        v = torch.arange(n, device=dev)
        mutual = (prop >= 0) & (prop[prop.clamp(min=0)] == v) & (v < prop)
        a, b = v[mutual], prop[mutual]
        mate[a], mate[b] = b, a
        if int((mate >= 0).sum()) // 2 >= need:
            break
    heads = torch.nonzero((mate >= 0) & (torch.arange(n, device=dev) < mate)).reshape(-1)
    if heads.numel() > need:                       # keep exactly `need` contractions
        drop = heads[need:]
        mate[mate[drop]] = -1
        mate[drop] = -1
    rep = torch.where((mate >= 0) & (mate < torch.arange(n, device=dev)), mate, torch.arange(n, device=dev))
    uniq, cluster = torch.unique(rep, return_inverse=True)
    n_c = int(uniq.numel())
    cr, cc = cluster[row], cluster[col]
    keep = cr < cc
    keyc = torch.unique(cr[keep] * n_c + cc[keep])
    e = torch.stack([keyc // n_c, keyc % n_c], dim=1)
    return cluster, n_c, edge_index_from_edges(e)


def pool_matrices(cluster: torch.Tensor, n_coarse: int):
    """The reference's ``pool_hash_to_mask`` / ``unpool_hash_to_mask`` (util/meshnet.py:331-341) for the (fine, coarse) pairs
    of ``cluster``: pool [n_coarse, n_fine] and unpool [n_fine, n_coarse] sparse COO matrices of ones."""
    n = int(cluster.numel())
    fine = torch.arange(n, device=cluster.device)
    ones = torch.ones(n, dtype=torch.float32, device=cluster.device)
    pool = torch.sparse_coo_tensor(torch.stack([cluster, fine]), ones, (n_coarse, n), check_invariants=False).coalesce()
    unpool = torch.sparse_coo_tensor(torch.stack([fine, cluster]), ones, (n, n_coarse), check_invariants=False).coalesce()
    return pool, unpool


def synth_pool_hierarchy(mesh: SynthMesh, levels: int = 3, ratio: float = 0.6, seed: int = 314):
    """edge_inds[0..levels], p_hashes[0..levels-1], up_hashes[0..levels-1] as MGCN.__init__ holds them."""
    edge_inds, p_hashes, up_hashes, sizes = [mesh.edge_index], [], [], [mesh.num_vertices]
    for l in range(levels):
        cluster, n_c, ei_c = synth_pool_level(edge_inds[-1], sizes[-1], ratio, seed + l)
        pool, unpool = pool_matrices(cluster, n_c)
        p_hashes.append(pool)
        up_hashes.append(unpool)
        edge_inds.append(ei_c)
        sizes.append(n_c)
    return {"edge_inds": edge_inds, "p_hashes": p_hashes, "up_hashes": up_hashes, "sizes": sizes}


def face_adjacency(faces: torch.Tensor) -> torch.Tensor:
    """``f2f`` [F, 3] of the reference's ``Mesh`` (util/mesh.py:215-227): the faces sharing a side with face i, padded with
    -1 where a side is on a boundary.  Vectorised half-edge matching; slot s holds the face across side
    (f[s], f[(s+1)%3]) -- the reference lists the same set in its ``Counter`` insertion order, which only changes the order
    of a 3-term sum in ``fn_bnf_detach_loss``.  Non-manifold sides (more than two faces) keep the first match."""
    f = faces.to(torch.int64)
    nf = int(f.shape[0])
    nv = int(f.max().item()) + 1 if nf else 0
    a, b = f.reshape(-1), f[:, [1, 2, 0]].reshape(-1)
    key = torch.minimum(a, b) * nv + torch.maximum(a, b)
    skey, order = torch.sort(key, stable=True)
    he_face = order // 3
    same_next = torch.zeros_like(skey, dtype=torch.bool)
    same_next[:-1] = skey[:-1] == skey[1:]
    same_prev = torch.zeros_like(skey, dtype=torch.bool)
    same_prev[1:] = skey[1:] == skey[:-1]
    mate = torch.full((3 * nf,), -1, dtype=torch.int64, device=f.device)
    idx = torch.arange(3 * nf, device=f.device)
    nxt = torch.where(same_next, he_face[(idx + 1).clamp(max=3 * nf - 1)], torch.full_like(idx, -1))
    prv = torch.where(same_prev, he_face[(idx - 1).clamp(min=0)], torch.full_like(idx, -1))
    mate[order] = torch.where(prv >= 0, prv, nxt)
    return mate.reshape(nf, 3)
