"""Host-side mirror of the reference's SGCN (util/networks.py:8-103) on our drop-in modules.

Same constructor (``device, activation, skip``), same ``forward(data, dm)`` contract, same
module tree => same ``state_dict`` keys (``blocks.{i}.module_{j}...``, ``skip_blocks.{i}...``).
The one addition is ``conv=``: the reference hard-codes ``"chebconv"`` (util/networks.py:13)
while BASELINE.json's metric is quoted on the ``"gcnconv"`` branch (util/networks.py:22-37);
both are exposed.  The reference file itself runs unchanged through ``semigcn_b200.compat``.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn as nn

from . import ops
from . import profile as _prof
from .nn import ChebConv, GCNConv, Sequential

SGCN_WIDTHS = [4, 16, 32, 64, 128, 256, 256, 512, 256, 256, 128, 64, 32, 16, 3]


class SingleScaleGCN(nn.Module):
    def __init__(self, device, activation: str = "lrelu", skip: bool = False, conv: str = "chebconv", widths=None):
        super().__init__()
        self.device, self.skip = device, skip
        self.comm = None        # set to a semigcn_b200.dist communicator in the vertex-partitioned mode
        h = list(widths) if widths is not None else list(SGCN_WIDTHS)
        self.h = h
        act = {"relu": nn.ReLU(), "lrelu": nn.LeakyReLU()}[activation]

        def mk(i):
            if conv == "gcnconv":
                return GCNConv(h[i], h[i + 1])
            if conv == "chebconv":
                return ChebConv(h[i], h[i + 1], K=3)
            raise ValueError(conv)

        nb = len(h) - 2
        blocks = []
        for i in range(nb - 1):
            blocks.append(Sequential("x, edge_index", [(mk(i), "x, edge_index -> x"), nn.BatchNorm1d(h[i + 1]), act]))
        blocks.append(Sequential("x, edge_index", [(mk(nb - 1), "x, edge_index -> x"), nn.BatchNorm1d(h[nb]), act,
                                                   (nn.Linear(h[nb], h[nb + 1]), "x -> x")]))
        self.blocks = nn.ModuleList(blocks)
        self.skip_blocks = nn.ModuleList([nn.Linear(h[i + 1] * 2, h[i + 1]) for i in range(min(6, nb))])

    def forward(self, data, dm=None):
        z1, x_pos, edge_index = data.z1.to(self.device), data.x_pos.to(self.device), data.edge_index.to(self.device)
        if type(dm) == np.ndarray:
            dm = torch.from_numpy(dm)
        elif type(dm) != torch.Tensor:
            dm = torch.ones([z1.shape[0], 1])
        dm = dm.to(self.device)
        if self.comm is None and z1.is_cuda and not z1.requires_grad and dm.numel() == z1.shape[0]:
            # util/networks.py:67-79 as two kernels (bounding box, then normalise + mask + concatenate), bit-identical
            x = ops.input_prep(z1, dm)
        else:
            # the torch form: a differentiable z1 (tests), or the partitioned mode (bounding box of the WHOLE mesh)
            prep = _prof.span("input_prep", 4.0 * 11 * z1.shape[0]) if _prof.ACTIVE is not None else None
            z_min, z_max = torch.min(z1, dim=0, keepdim=True)[0], torch.max(z1, dim=0, keepdim=True)[0]
            if self.comm is not None:           # util/networks.py:67-68 over all ranks
                z_min, z_max = self.comm.all_reduce_min(z_min.clone()), self.comm.all_reduce_max(z_max.clone())
            z_sc = torch.max(z_max - z_min)
            zc = (z_min + z_max) * 0.5
            z1 = (z1 - zc) / z_sc
            z1 = dm * z1
            x = torch.cat([z1, dm], dim=1)
            if prep is not None:
                prep.close()
        skip_in = []
        nblk = len(self.blocks)
        for i, b in enumerate(self.blocks):
            if i <= 5:
                x = b(x, edge_index)
                skip_in.append(x)
            elif i <= 7:
                x = b(x, edge_index)
            else:
                if self.skip:
                    lin = self.skip_blocks[nblk - i]
                    x = ops.linear(torch.cat([skip_in[nblk - i], x], dim=1), lin.weight, lin.bias)
                x = b(x, edge_index)
        return x_pos + x
