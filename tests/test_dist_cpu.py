"""Host logic of the vertex-partitioned mode on CPU: partition plans and the all-to-all halo exchange over a
world_size-2/3 ``gloo`` group (no GPU, no compute kernels)."""
import os
import socket

import pytest
import torch
import torch.multiprocessing as mp

from semigcn_b200 import meshgen, partition


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


@pytest.mark.parametrize("world", [2, 3, 8])
def test_plans_are_mutually_consistent(world):
    m = meshgen.icosphere(7)
    n = m.num_vertices
    plans = [partition.build_plan(m.edge_index, n, r, world) for r in range(world)]
    assert sum(p.n_own for p in plans) == n
    covered = torch.zeros(m.nnz, dtype=torch.int64)
    for r, p in enumerate(plans):
        assert p.recv_counts[r] == 0 and p.send_counts[r] == 0
        off = 0
        for q in range(world):
            # what r expects from q is exactly what q sends to r, in the same order
            so = sum(plans[q].send_counts[:r])
            sent = plans[q].send_idx[so:so + plans[q].send_counts[r]].long() + plans[q].lo
            assert torch.equal(p.ghost_gid[off:off + p.recv_counts[q]], sent)
            off += p.recv_counts[q]
        # local edges map back to global ones; every edge with an owned target appears exactly once
        gid = torch.cat([torch.arange(p.lo, p.hi), p.ghost_gid])
        ge = gid[p.edge_index]
        own_t = (ge[1] >= p.lo) & (ge[1] < p.hi)
        key = ge[0][own_t] * n + ge[1][own_t]
        allkey = m.edge_index[0] * n + m.edge_index[1]
        pos = torch.searchsorted(torch.sort(allkey)[0], key)
        assert torch.equal(torch.sort(allkey)[0][pos], key)
        covered += torch.isin(allkey, key).long()
    assert torch.all(covered == 1)


def _worker(rank, world, port, freq):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    import torch.distributed as dist
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from semigcn_b200.dist import TorchComm
        comm = TorchComm()
        m = meshgen.icosphere(freq)
        n = m.num_vertices
        plan = partition.build_plan(m.edge_index, n, rank, world)
        g = torch.Generator().manual_seed(0)
        x = torch.randn(n, 5, generator=g)                     # same global features on every rank
        x_own = x[plan.lo:plan.hi]
        send = x_own[plan.send_idx.long()]                      # the pack step (sgb_gather_rows on the GPU)
        ghosts = comm.all_to_all_rows(send, plan.send_counts, plan.recv_counts)
        assert torch.equal(ghosts, x[plan.ghost_gid])
        # a partitioned plain-adjacency propagation equals the global one on the owned rows
        x_full = torch.cat([x_own, ghosts])
        src, dst = plan.edge_index
        keep = dst < plan.n_own
        y = torch.zeros(plan.n_own, 5).index_add_(0, dst[keep], x_full[src[keep]])
        y_ref = torch.zeros(n, 5).index_add_(0, m.edge_index[1], x[m.edge_index[0]])[plan.lo:plan.hi]
        assert torch.allclose(y, y_ref, atol=1e-5)
        # reductions used by SyncBN / the loss
        t = torch.tensor([float(rank + 1)])
        assert comm.all_reduce_sum(t.clone()).item() == world * (world + 1) / 2
        assert comm.all_reduce_max(t.clone()).item() == world and comm.all_reduce_min(t.clone()).item() == 1
        cat = comm.all_gather_cat(torch.full((2, 3), float(rank)))
        assert cat.shape == (2 * world, 3) and cat[2 * rank, 0].item() == rank
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_halo_exchange_over_gloo(world):
    mp.spawn(_worker, args=(world, _free_port(), 6), nprocs=world, join=True)


def test_merge_moment_rows_matches_direct_statistics():
    """SyncBN pre-reduction (semigcn_b200.ops.merge_moment_rows): folding per-tile (count, mean, M2) rows -- ragged tiles,
    empty rows, a near-constant channel -- gives the statistics of the concatenated data."""
    from semigcn_b200.ops import merge_moment_rows
    g = torch.Generator().manual_seed(3)
    c = 7
    sizes = [128, 128, 0, 37, 1, 128, 0, 5]
    chunks = [torch.randn(s, c, generator=g, dtype=torch.float64) * 3.0 + torch.arange(c) for s in sizes]
    for ch in chunks:
        if ch.shape[0]:
            ch[:, 2] = 1000.0 + 1e-3 * ch[:, 2]          # |mean| >> std: sum / sum-of-squares forms lose this one
    rows = []
    for ch in chunks:
        if ch.shape[0] == 0:
            rows.append(torch.zeros(3, c, dtype=torch.float64))
        else:
            mu = ch.mean(0)
            rows.append(torch.stack([torch.full((c,), float(ch.shape[0]), dtype=torch.float64), mu, ((ch - mu) ** 2).sum(0)]))
    part = torch.stack(rows).to(torch.float32)
    out = merge_moment_rows(part)
    assert out.shape == (1, 3, c) and out.dtype == torch.float32
    allx = torch.cat(chunks)
    assert torch.equal(out[0, 0], torch.full((c,), float(allx.shape[0])))
    assert torch.allclose(out[0, 1].double(), allx.mean(0), rtol=1e-6, atol=0)
    m2 = ((allx - allx.mean(0)) ** 2).sum(0)
    keep = torch.tensor([q != 2 for q in range(c)])
    assert torch.allclose(out[0, 2].double()[keep], m2[keep], rtol=1e-5, atol=0)
    # the near-constant channel: the per-row means were rounded to float32 (ulp 6e-5 at 1000, spread 3e-3) before merging
    assert torch.allclose(out[0, 2].double()[2], m2[2], rtol=2e-3, atol=0)
    # all-empty input stays empty (no NaN)
    z = merge_moment_rows(torch.zeros(4, 3, c))
    assert torch.equal(z, torch.zeros(1, 3, c))
    # merging merged rows of two halves == merging everything (associativity, what the all-gather relies on)
    two = torch.cat([merge_moment_rows(part[:4]), merge_moment_rows(part[4:])])
    again = merge_moment_rows(two)
    assert torch.allclose(again.double()[..., keep], out.double()[..., keep], rtol=1e-6, atol=0)
    assert torch.allclose(again.double()[..., 2], out.double()[..., 2], rtol=1e-3, atol=0)     # float32 rows in between


def test_morton_order_balances_the_cut():
    """partition.morton_order: a permutation; renumbering keeps the graph; contiguous ranges of the new numbering give
    every rank a similar halo (the generator's numbering gives rank 0 -- the owner of the icosahedron edge vertices --
    several times the average), and the plans stay mutually consistent."""
    m = meshgen.icosphere(30)
    n, world = m.num_vertices, 8
    perm = partition.morton_order(m.vs)
    assert torch.equal(torch.sort(perm)[0], torch.arange(n))
    ei2, vs2 = partition.renumber(perm, m.edge_index, m.vs)
    assert torch.equal(vs2, m.vs[perm])
    # same undirected graph: degrees are permuted, edge count unchanged, endpoints map back
    assert torch.equal(torch.bincount(ei2[1], minlength=n), torch.bincount(m.edge_index[1], minlength=n)[perm])
    assert torch.equal(perm[ei2], m.edge_index)
    ghosts_given = [partition.build_plan(m.edge_index, n, r, world).n_ghost for r in range(world)]
    plans = [partition.build_plan(ei2, n, r, world) for r in range(world)]
    ghosts = [p.n_ghost for p in plans]
    mean_g, mean_given = sum(ghosts) / world, sum(ghosts_given) / world
    assert max(ghosts) <= 2.0 * mean_g, (ghosts, ghosts_given)
    assert max(ghosts) < max(ghosts_given), (ghosts, ghosts_given)
    for r, p in enumerate(plans):
        off = 0
        for q in range(world):
            so = sum(plans[q].send_counts[:r])
            sent = plans[q].send_idx[so:so + plans[q].send_counts[r]].long() + plans[q].lo
            assert torch.equal(p.ghost_gid[off:off + p.recv_counts[q]], sent)
            off += p.recv_counts[q]
