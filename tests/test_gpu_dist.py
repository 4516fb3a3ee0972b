"""Vertex-partitioned SGCN == unpartitioned SGCN (forward, loss, parameter gradients).
Two / three processes share cuda:0 and talk through a gloo group (CUDA tensors staged through the host), so the
whole partitioned path -- plans, sgb_gather_rows packs, all-to-all halo exchange, sgb_spmm_halo, SyncBN
all-gather / all-reduce, gradient sum -- runs without a multi-GPU box.  The NCCL run is bench.py --mode partition."""
import os
import socket

import pytest
import torch
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _rel(a, b):
    return (a.double() - b.double()).abs().max().item() / max(b.double().abs().max().item(), 1e-30)


def _worker(rank, world, port, conv, slope=0.01, ranges=None, order="given", with_normals=False):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    import torch.distributed as dist
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from semigcn_b200 import meshgen, partition, losses
        from semigcn_b200.data import Data
        from semigcn_b200.dist import TorchComm, register_partition, sync_gradients, dist_mask_pos_rec_loss
        from semigcn_b200.networks import SingleScaleGCN
        dev = torch.device("cuda:0")
        torch.cuda.set_device(dev)
        comm = TorchComm()
        prob = meshgen.synth_inpainting_problem(10, smooth_iters=5, n_dummy=2)
        mesh = prob["mesh"]
        n = mesh.num_vertices
        ei = mesh.edge_index.to(dev)
        z1, x_pos, dm = prob["z1"].to(dev), prob["x_pos"].to(dev), prob["vmask_dummy"][:, :1].to(dev)
        target, vmask = prob["ini_vs"].to(dev), prob["v_mask"].to(dev)
        faces, fn_t, fmask = mesh.faces.to(dev), prob["fn"].to(dev), prob["f_mask"].to(dev)
        if order == "interior":
            # Morton patches, then interior vertices before boundary ones inside every rank's range: the interior rows are
            # aggregated while the halo all-to-all is in flight (two launches per propagation, ops.spmm)
            perm = partition.morton_order(mesh.vs.to(dev))
            ei, z1, x_pos, dm, target, vmask = partition.renumber(perm, ei, z1, x_pos, dm, target, vmask)
            inv = torch.empty_like(perm)
            inv[perm] = torch.arange(n, device=dev)
            perm2 = partition.interior_first_order(ei, n, ranges or partition.vertex_ranges(n, world))
            ei, z1, x_pos, dm, target, vmask = partition.renumber(perm2, ei, z1, x_pos, dm, target, vmask)
            inv2 = torch.empty_like(perm2)
            inv2[perm2] = torch.arange(n, device=dev)
            faces = inv2[inv[faces]]
            ei = ei.contiguous()
        torch.manual_seed(314)
        net = SingleScaleGCN(dev, conv=conv).to(dev)
        for mod in net.modules():
            if isinstance(mod, torch.nn.LeakyReLU):
                mod.negative_slope = slope
        # ---- unpartitioned reference on this process
        out_ref = net(Data(z1=z1, x_pos=x_pos, edge_index=ei), dm)
        loss_ref = losses.mask_pos_rec_loss(out_ref, target, vmask)
        if with_normals:      # the full step loss of sgcn.py:130-137
            loss_ref = loss_ref + 4.0 * losses.mask_norm_rec_loss(losses.compute_fn(out_ref, faces), fn_t, fmask)
        loss_ref.backward()
        grads_ref = {k: p.grad.clone() for k, p in net.named_parameters() if p.grad is not None}
        net.zero_grad(set_to_none=True)
        # ---- partitioned
        plan = partition.build_plan(ei, n, rank, world, ranges=ranges)
        ei_local = register_partition(plan, comm)
        net.comm = comm
        lo, hi = plan.lo, plan.hi
        out = net(Data(z1=z1[lo:hi].contiguous(), x_pos=x_pos[lo:hi].contiguous(), edge_index=ei_local), dm[lo:hi].contiguous())
        loss = dist_mask_pos_rec_loss(out, target[lo:hi], vmask[lo:hi], comm)
        if with_normals:
            from semigcn_b200.dist import dist_mask_norm_rec_loss
            from semigcn_b200 import ops
            from semigcn_b200._lib import MODE_GCN, MODE_CHEB
            g_part = ops.graph_for(ei_local, hi - lo, MODE_GCN if conv == "gcnconv" else MODE_CHEB)
            fid, f_loc = partition.local_faces(plan, faces)
            loss = loss + 4.0 * dist_mask_norm_rec_loss(out, g_part.halo, f_loc, fn_t[fid], fmask[fid], comm)
            if order == "interior":
                assert 0 < g_part.n_interior < hi - lo
        loss.backward()
        sync_gradients(net, comm)
        torch.cuda.synchronize()
        assert out.shape == (hi - lo, 3)
        e_out = _rel(out, out_ref[lo:hi])
        assert e_out <= 1e-5, f"rank {rank}: output differs {e_out:.2e}"
        # the position loss agrees to 1e-6; the face-normal term is a function of the OUTPUT (within 1e-5 of the unpartitioned one)
        # through normals of faces whose sides are ~0.1 long: an output deviation of a few 1e-6 moves it by up to ~1e-4 relative
        assert abs(loss.item() - loss_ref.item()) <= (2e-4 if with_normals else 1e-6) * abs(loss_ref.item()), (loss.item(), loss_ref.item())
        worst, who = 0.0, ""
        for k, p in net.named_parameters():
            if p.grad is None:
                continue
            gr = grads_ref[k]
            if k.endswith("module_0.bias"):
                # conv bias in front of BatchNorm: the gradient is analytically zero, both sides hold rounding noise
                assert p.grad.abs().max().item() <= 1e-4 * max(g.abs().max().item() for g in grads_ref.values())
                continue
            e = _rel(p.grad, gr)
            # slope = 1 (every LeakyReLU is the identity: a smooth network): what is left between the partitioned and the
            # unpartitioned run is the reassociation of the BatchNorm sums across ranks.  Weight gradients within 1e-3,
            # BatchNorm gamma / beta gradients (column sums with heavy cancellation) within 5e-3 -- measured on a B200:
            # <= 7e-4 / 2.7e-3 (gcnconv), 6e-5 / 2.3e-4 (chebconv), 3 ranks.
            # slope = 0.01 (the reference's LeakyReLU): a pre-activation within rounding of zero can take the other branch in
            # the partitioned run (SyncBN merges one moment row per rank: a different, equally valid, summation order), and one
            # flipped unit moves a weight gradient of this 1 002-vertex mesh by up to ~1e-2 (measured 9.4e-3, [3-chebconv];
            # the same network with slope 1 agrees to 6e-5).  The kinked run is therefore a sanity bound (a wrong factor or a
            # missing halo row would be O(1)); the smooth run above is the strict one.
            if slope == 1.0:
                tol = 5e-3 if ".module_1." in k else 1e-3
            else:
                tol = 5e-2 if ".module_1." in k else 2e-2
            if e / tol > worst:
                worst, who = e / tol, f"{k} ({e:.2e})"
        if rank == 0:
            print(f"world {world} {conv} slope {slope}: output {e_out:.2e}; worst parameter-gradient error / tolerance {worst:.3f} at {who}")
        assert worst <= 1.0, f"rank {rank}: parameter gradients differ at {who}"
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("conv", ["gcnconv", "chebconv"])
@pytest.mark.parametrize("world", [2, 3])
def test_partitioned_sgcn_equals_single_gpu(world, conv):
    mp.spawn(_worker, args=(world, _free_port(), conv), nprocs=world, join=True)


@pytest.mark.parametrize("conv", ["gcnconv", "chebconv"])
def test_partitioned_sgcn_smooth_network_is_strict(conv):
    """Triage of round 1's open question (SyncBN with one pre-merged moment row per rank moved [3-chebconv] outside its
    gradient tolerance: kink flip or defect?): with LeakyReLU slope 1 the network has no kinks, so the partitioned
    gradients must match much tighter than with the kinked activation.  They do (chebconv: 6e-5 on the weights where the
    kinked run shows 9.4e-3) => the round-1 deviation was a kink flip, not a defect of the pre-merged SyncBN rows."""
    mp.spawn(_worker, args=(3, _free_port(), conv, 1.0), nprocs=3, join=True)


@pytest.mark.parametrize("world,ranges", [(2, [(0, 385), (385, 1002)]), (3, [(0, 130), (130, 700), (700, 1002)])])
def test_partitioned_sgcn_uneven_ranges(world, ranges):
    """Uneven vertex ranges straddling the 128-row tile boundary: the ranks' kernels produce different numbers of
    BatchNorm partial rows; SyncBN gathers ONE merged row per rank, so the payload is rank-invariant."""
    mp.spawn(_worker, args=(world, _free_port(), "gcnconv", 0.01, ranges), nprocs=world, join=True)


@pytest.mark.parametrize("conv", ["gcnconv", "chebconv"])
def test_partitioned_sgcn_overlapped_halo_and_normal_loss(conv):
    """Interior-first numbering (interior rows aggregated while the halo all-to-all is in flight: two launches per
    propagation over disjoint row ranges) + the face-normal loss across the cuts (ghost positions through the exchange,
    their gradients back through the reverse exchange): the partitioned step equals the single-GPU step of sgcn.py:129-143."""
    mp.spawn(_worker, args=(3, _free_port(), conv, 1.0, None, "interior", True), nprocs=3, join=True)
