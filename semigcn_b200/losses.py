"""Step-path mesh losses (reference util/models.py:121-126, util/loss.py:14-34,78-107), GPU side.

Mirrors ``compute_fn`` / ``mask_pos_rec_loss`` ("rmse") / ``mask_norm_rec_loss`` ("l1mae") with the
same argument meaning; float64 targets give float64 losses exactly as sgcn.py:127,131-132 does.
Written mask-multiplicatively (no boolean-index gather => no host sync per step).
"""
from __future__ import annotations

import torch
from torch import Tensor


def compute_fn(vs: Tensor, faces: Tensor) -> Tensor:
    """Unit face normals, util/models.py:121-126."""
    a = vs[faces[:, 0]]
    n = torch.linalg.cross(vs[faces[:, 1]] - a, vs[faces[:, 2]] - a)
    return n / torch.sqrt(torch.sum(n * n, dim=1, keepdim=True))


def mask_pos_rec_loss(pred_pos: Tensor, real_pos: Tensor, mask: Tensor) -> Tensor:
    """sqrt(mean_{i in mask} |real_i - pred_i|^2 + 1e-6), util/loss.py:23-28."""
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


def sgcn_step_loss(pos: Tensor, faces: Tensor, ini_vs: Tensor, fn_real: Tensor, v_mask: Tensor, f_mask: Tensor,
                   k1: float = 4.0) -> Tensor:
    """loss = loss_p + k1 * loss_n of sgcn.py:130-137 (non-CAD branch)."""
    norm = compute_fn(pos, faces)
    return mask_pos_rec_loss(pos, ini_vs, v_mask) + k1 * mask_norm_rec_loss(norm, fn_real, f_mask)
