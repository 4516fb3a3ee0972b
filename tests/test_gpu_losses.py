"""GPU parity of the fused step-loss kernels (csrc/loss.cu) against (a) the golden values produced by the
reference's own util/loss.py / util/models.py (tests/golden/make_golden.py) and (b) oracle autograd."""
import numpy as np
import pytest
import torch

from helpers import assert_bit_equal, load_golden, rel_err
from oracle import pyg_ref as O

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _losses():
    from semigcn_b200 import losses
    return losses


@pytest.mark.parametrize("n", [4, 8])
def test_incidence_bit_exact(n):
    gold = load_golden(n)
    faces = torch.from_numpy(gold["faces"]).long()
    nv = gold["vs"].shape[0]
    topo = _losses().FaceTopology(faces.to(DEV), nv)
    flat = faces.reshape(-1)
    order = torch.sort(flat, stable=True)[1]                    # ascending corner id inside each vertex row
    rowptr = torch.zeros(nv + 1, dtype=torch.int64)
    rowptr[1:] = torch.cumsum(torch.bincount(flat, minlength=nv), 0)
    assert_bit_equal(topo.rowptr, rowptr.to(torch.int32), "incidence rowptr")
    assert_bit_equal(topo.inc[:order.numel()], order.to(torch.int32), "incidence entries")


def test_incidence_rejects_bad_faces():
    from semigcn_b200 import SgbError
    with pytest.raises(SgbError):
        _losses().FaceTopology(torch.tensor([[0, 1, 7]], device=DEV), 3)


@pytest.mark.parametrize("n", [4, 8])
def test_step_losses_match_reference_golden(n):
    """Values the reference's util/loss.py + util/models.py produced (fp64 targets = sgcn.py, fp32 = mgcn.py)."""
    L = _losses()
    gold = load_golden(n)
    pred = torch.from_numpy(gold["pred"]).to(DEV)
    faces = torch.from_numpy(gold["faces"]).long().to(DEV)
    lp, ln = L.sgcn_step_losses(pred, faces, gold["vs"], gold["fn"], gold["v_mask"], gold["f_mask"])
    assert lp.dtype == torch.float64 and ln.dtype == torch.float64
    assert abs(lp.item() - gold["loss_pos_f64"].item()) <= 1e-12 * abs(gold["loss_pos_f64"].item())
    assert abs(ln.item() - gold["loss_norm_f64"].item()) <= 1e-7 * abs(gold["loss_norm_f64"].item())
    lp32, ln32 = L.sgcn_step_losses(pred, faces, torch.from_numpy(gold["vs"]).float(), torch.from_numpy(gold["fn"]).float(),
                                    gold["v_mask"], gold["f_mask"])
    assert lp32.dtype == torch.float32
    assert abs(lp32.item() - gold["loss_pos_f32"].item()) <= 1e-6 * abs(gold["loss_pos_f32"].item())
    assert abs(ln32.item() - gold["loss_norm_f32"].item()) <= 1e-6 * abs(gold["loss_norm_f32"].item())
    # numpy faces (what sgcn.py passes) go through the same cache
    lp_np, _ = L.sgcn_step_losses(pred, gold["faces"], gold["vs"], gold["fn"], gold["v_mask"], gold["f_mask"])
    assert lp_np.item() == lp.item()
    ll = L.mesh_laplacian_loss(pred, torch.from_numpy(gold["edge_index"]).to(DEV))
    assert abs(ll.item() - gold["loss_lap_f32"].item()) <= 2e-6 * abs(gold["loss_lap_f32"].item())


@pytest.mark.parametrize("t64", [True, False])
def test_step_loss_gradient_vs_oracle_autograd(t64):
    L = _losses()
    from semigcn_b200 import meshgen
    prob = meshgen.synth_inpainting_problem(12, smooth_iters=3, n_dummy=1)
    mesh = prob["mesh"]
    g = torch.Generator().manual_seed(5)
    pos = (prob["ini_vs"].float() + 0.01 * torch.randn(mesh.num_vertices, 3, generator=g))
    tpos = prob["ini_vs"] if t64 else prob["ini_vs"].float()
    tfn = prob["fn"] if t64 else prob["fn"].float()
    p_ref = pos.clone().requires_grad_(True)
    lp_r = O.mask_pos_rec_loss(p_ref, tpos, prob["v_mask"])
    ln_r = O.mask_norm_rec_loss(O.compute_fn(p_ref, mesh.faces), tfn, prob["f_mask"])
    (lp_r + 4.0 * ln_r).backward()
    p = pos.clone().to(DEV).requires_grad_(True)
    loss = L.sgcn_step_loss(p, mesh.faces.to(DEV), tpos, tfn, prob["v_mask"], prob["f_mask"], 4.0)
    loss.backward()
    tol = 1e-12 if t64 else 2e-6
    assert abs(loss.item() - (lp_r + 4.0 * ln_r).item()) <= max(tol, 1e-7) * abs(loss.item())
    assert rel_err(p.grad, p_ref.grad) <= 1e-5
    # deterministic
    p2 = pos.clone().to(DEV).requires_grad_(True)
    L.sgcn_step_loss(p2, mesh.faces.to(DEV), tpos, tfn, prob["v_mask"], prob["f_mask"], 4.0).backward()
    assert torch.equal(p.grad, p2.grad)


def test_masks_none_equals_all_ones_and_pos_only():
    L = _losses()
    from semigcn_b200 import meshgen
    mesh = meshgen.icosphere(10)
    torch.manual_seed(0)
    pos = (mesh.vs + 0.02 * torch.randn_like(mesh.vs)).to(DEV)
    tfn = meshgen.face_normals(mesh.vs.double(), mesh.faces)
    ones_v = torch.ones(mesh.num_vertices, dtype=torch.bool)
    ones_f = torch.ones(mesh.faces.shape[0], dtype=torch.bool)
    a = L.sgcn_step_losses(pos, mesh.faces.to(DEV), mesh.vs.double(), tfn, None, None)
    b = L.sgcn_step_losses(pos, mesh.faces.to(DEV), mesh.vs.double(), tfn, ones_v, ones_f)
    assert a[0].item() == b[0].item() and a[1].item() == b[1].item()
    lp = L.fused_mask_pos_rec_loss(pos, mesh.vs.double(), ones_v)
    assert lp.item() == a[0].item()
    want = O.mask_pos_rec_loss(pos.cpu(), mesh.vs.double(), ones_v)
    assert abs(lp.item() - want.item()) <= 1e-12 * want.item()


def test_laplacian_loss_gradient_vs_oracle():
    L = _losses()
    from semigcn_b200 import meshgen
    mesh = meshgen.icosphere(9)
    torch.manual_seed(3)
    pos = mesh.vs + 0.05 * torch.randn_like(mesh.vs)
    p_ref = pos.clone().requires_grad_(True)
    l_ref = O.mesh_laplacian_loss(p_ref, mesh.edge_index)
    l_ref.backward()
    p = pos.clone().to(DEV).requires_grad_(True)
    l = L.mesh_laplacian_loss(p, mesh.edge_index.to(DEV))
    l.backward()
    assert abs(l.item() - l_ref.item()) <= 2e-6 * abs(l_ref.item())
    assert rel_err(p.grad, p_ref.grad) <= 1e-5


# ------------------------------------------------------------------------------------------------------------------
# fused bilateral-normal-filter regulariser (csrc/bnf_loss.cu) -- SURVEY.md §8(f)-1
# ------------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("n", [4, 8])
def test_bnf_loss_matches_reference_golden(n):
    """Loss and filtered normals the reference's own Loss.fn_bnf_detach_loss (util/loss.py:197-253) produced on its own
    Mesh.f2f (tests/golden/make_golden.py), numpy faces / f2f as sgcn.py:134 passes them."""
    L = _losses()
    gold = load_golden(n)
    pred = torch.from_numpy(gold["pred"]).to(DEV)
    fn_pred = torch.from_numpy(gold["compute_fn_f32"]).to(DEV)
    loss, new_fn = L.fn_bnf_detach_loss(pred, fn_pred, gold["faces"], gold["f2f"], loop=5)
    assert loss.dtype == torch.float32 and new_fn.shape == fn_pred.shape
    assert torch.allclose(new_fn.cpu(), torch.from_numpy(gold["bnf_fn_f32"]), rtol=0, atol=1e-6)
    assert abs(loss.item() - float(gold["loss_bnf_f32"])) <= 1e-6 * abs(float(gold["loss_bnf_f32"]))
    # compute_fn through our own fused kernel feeds it too (what the step does: sgcn.py:130,134)
    loss2, _ = L.fn_bnf_detach_loss(pred, L.compute_fn(pred, torch.from_numpy(gold["faces"]).long().to(DEV)), gold["faces"], gold["f2f"])
    assert abs(loss2.item() - loss.item()) <= 1e-6 * abs(loss.item())


@pytest.mark.parametrize("ltype", ["mae", "l1mae", "rmse", "l1rmse"])
@pytest.mark.parametrize("loop", [0, 1, 5])
def test_bnf_loss_all_types_and_gradient_vs_oracle(ltype, loop):
    """Every ltype / loop count against the pinned CPU restatement (oracle/loss_ref.py), forward and the gradient with
    respect to fn, on a mesh WITH a boundary (f2f == -1 slots: the reference's wrap-around indexing of the last face)."""
    from oracle import loss_ref
    from semigcn_b200 import meshgen
    L = _losses()
    gold = load_golden(8)
    faces = torch.from_numpy(gold["faces"]).long()
    keep = torch.ones(faces.shape[0], dtype=torch.bool)
    keep[::7] = False
    faces = faces[keep].contiguous()
    f2f = meshgen.face_adjacency(faces)
    assert int((f2f == -1).sum()) > 0
    pred = torch.from_numpy(gold["pred"])
    fn0 = O.compute_fn(pred, faces)
    g = torch.Generator().manual_seed(5)
    fn_in = (fn0 + 0.05 * torch.randn(fn0.shape, generator=g)).requires_grad_(True)      # generic input, not unit length
    loss_r, new_r = loss_ref.fn_bnf_detach_loss(pred, fn_in, faces, f2f, ltype=ltype, loop=loop)
    loss_r.backward()
    fn_dev = fn_in.detach().to(DEV).requires_grad_(True)
    loss, new_fn = L.fn_bnf_detach_loss(pred.to(DEV), fn_dev, faces.to(DEV), f2f.to(DEV), ltype=ltype, loop=loop)
    (3.0 * loss).backward()
    assert abs(loss.item() - loss_r.item()) <= 2e-6 * abs(loss_r.item()), (loss.item(), loss_r.item())
    assert torch.allclose(new_fn.cpu(), new_r, rtol=0, atol=2e-6)
    assert not new_fn.requires_grad
    assert rel_err(fn_dev.grad, 3.0 * fn_in.grad) <= 1e-5


def test_bnf_loss_refuses_cpu_tensors():
    from semigcn_b200 import SgbError
    gold = load_golden(4)
    with pytest.raises(SgbError):
        _losses().fn_bnf_detach_loss(torch.from_numpy(gold["pred"]), torch.from_numpy(gold["compute_fn_f32"]), gold["faces"], gold["f2f"])
