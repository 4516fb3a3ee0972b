"""Whole-step CUDA graph (semigcn_b200/graphed.py) == the eager step: same parameters after the same mask schedule."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _setup(conv, dev):
    from semigcn_b200 import meshgen, losses
    from semigcn_b200.networks import SingleScaleGCN
    prob = meshgen.synth_inpainting_problem(8, device=dev, smooth_iters=5, n_dummy=6)
    mesh = prob["mesh"]
    torch.manual_seed(314)
    net = SingleScaleGCN(dev, conv=conv).to(dev)
    fn = meshgen.face_normals(prob["ini_vs"], mesh.faces)
    f_mask = prob["v_mask"][mesh.faces].all(dim=1)

    def loss_fn(out):
        return losses.sgcn_step_loss(out, mesh.faces, prob["ini_vs"], fn, prob["v_mask"], f_mask, 4.0)

    return prob, mesh, net, loss_fn


@pytest.mark.parametrize("conv", ["gcnconv", "chebconv"])
def test_graphed_step_matches_eager(conv):
    from semigcn_b200.data import Data
    from semigcn_b200.graphed import GraphedTrainStep
    dev = torch.device("cuda:0")
    prob, mesh, net_e, loss_fn = _setup(conv, dev)
    _, _, net_g, _ = _setup(conv, dev)
    net_g.load_state_dict(net_e.state_dict())
    masks = [prob["vmask_dummy"][:, i:i + 1].contiguous().float() for i in range(6)]
    opt_e = torch.optim.Adam(net_e.parameters(), lr=0.01, capturable=True)
    opt_g = torch.optim.Adam(net_g.parameters(), lr=0.01, capturable=True)
    data = Data(z1=prob["z1"], x_pos=prob["x_pos"], edge_index=mesh.edge_index)
    losses_e = []
    for dm in masks:
        opt_e.zero_grad(set_to_none=True)
        loss = loss_fn(net_e(data, dm))
        loss.backward()
        opt_e.step()
        losses_e.append(float(loss.detach()))
    step = GraphedTrainStep(net_g, loss_fn, opt_g, prob["z1"], prob["x_pos"], mesh.edge_index, masks[0])
    losses_g = [float(step(dm)) for dm in masks]
    torch.cuda.synchronize()
    for a, b in zip(losses_e, losses_g):
        assert abs(a - b) <= 1e-6 * max(abs(a), 1e-12), (losses_e, losses_g)
    for (k, pe), (_, pg) in zip(net_e.state_dict().items(), net_g.state_dict().items()):
        if pe.dtype.is_floating_point:
            err = (pe - pg).abs().max().item() / max(pe.abs().max().item(), 1e-30)
            assert err <= 1e-5, f"{k}: {err:.2e}"
        else:
            assert torch.equal(pe, pg), k


def test_graphed_step_accumulate_matches_reference_loop():
    """sgcn.py:118-146: zero_grad, 3 x (forward, loss, backward), one optimizer step."""
    from semigcn_b200.data import Data
    from semigcn_b200.graphed import GraphedTrainStep
    dev = torch.device("cuda:0")
    prob, mesh, net_e, loss_fn = _setup("gcnconv", dev)
    _, _, net_g, _ = _setup("gcnconv", dev)
    net_g.load_state_dict(net_e.state_dict())
    masks = [prob["vmask_dummy"][:, i:i + 1].contiguous().float() for i in range(6)]
    opt_e = torch.optim.Adam(net_e.parameters(), lr=0.01)
    opt_g = torch.optim.Adam(net_g.parameters(), lr=0.01)
    data = Data(z1=prob["z1"], x_pos=prob["x_pos"], edge_index=mesh.edge_index)
    for b in range(2):
        opt_e.zero_grad()
        for dm in masks[3 * b:3 * b + 3]:
            loss_fn(net_e(data, dm)).backward()
        opt_e.step()
    step = GraphedTrainStep(net_g, loss_fn, opt_g, prob["z1"], prob["x_pos"], mesh.edge_index, masks[0], accumulate=3)
    for dm in masks:
        step(dm)
    torch.cuda.synchronize()
    for (k, pe), (_, pg) in zip(net_e.state_dict().items(), net_g.state_dict().items()):
        if pe.dtype.is_floating_point:
            err = (pe - pg).abs().max().item() / max(pe.abs().max().item(), 1e-30)
            assert err <= 1e-5, f"{k}: {err:.2e}"
