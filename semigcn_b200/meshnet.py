"""Host-side mirror of the reference's multi-resolution network (util/meshnet.py) on our drop-in modules.

``MeshPool`` / ``MeshUnpool`` / ``DownConv`` / ``UpConv`` keep the reference's constructors and module trees
(``model1.module_{i}`` / ``model2.module_{i}`` => the same ``state_dict`` keys, util/meshnet.py:31-160); ``MGCN`` keeps
the reference's sub-module names and ``forward`` (util/meshnet.py:212-317).  The one difference: the reference's
``MGCN.__init__`` builds the pooling hierarchy itself by QEM-simplifying its ``Mesh`` objects (util/meshnet.py:169-193;
host-side, out of scope), so the mirror takes the finished hierarchy -- ``edge_inds``, ``p_hashes``,
``up_hashes``, ``smposs_list`` -- as an argument.  The reference file itself runs unchanged through
``semigcn_b200.compat`` (convs / Sequential) and ``compat.patch_meshnet`` (MeshPool / MeshUnpool).
"""
from __future__ import annotations

from typing import List, Sequence

import numpy as np
import torch
import torch.nn as nn
from torch import Tensor

from . import ops
from .nn import ChebConv, GCNConv, MeshPool, MeshUnpool, Sequential  # noqa: F401  (re-exported)


def _conv(kind: str, cin: int, cout: int, K: int):
    return ChebConv(cin, cout, K=K) if kind == "chebconv" else GCNConv(cin, cout)


def _block(kind: str, cin: int, cout: int, K: int, mid=None) -> list:
    mods = [(_conv(kind, cin, cout, K), "x, edge_index -> x")]
    if mid is not None:
        mods.append((mid, "x -> x"))
    mods += [(nn.BatchNorm1d(cout), "x -> x"), (nn.LeakyReLU(), "x -> x")]
    return mods


class DownConv(nn.Module):
    """util/meshnet.py:31-91: conv-BN-act, conv-MeshPool-BN-act on the fine graph, then 3 x conv-BN-act + Dropout on the
    coarse graph."""

    def __init__(self, in_channels, out_channels, edge_index1, edge_index2, pool_hash, K=3, drop_rate=0.0, conv="chebconv"):
        super().__init__()
        self.edge_index1, self.edge_index2 = edge_index1, edge_index2
        self.model1 = Sequential("x, edge_index", _block(conv, in_channels, out_channels, K)
                                 + _block(conv, out_channels, out_channels, K, MeshPool(pool_hash)))
        self.model2 = Sequential("x, edge_index", _block(conv, out_channels, out_channels, K) + _block(conv, out_channels, out_channels, K)
                                 + _block(conv, out_channels, out_channels, K) + [(nn.Dropout(drop_rate), "x -> x")])

    def forward(self, input: Tensor) -> Tensor:
        return self.model2(self.model1(input, self.edge_index1), self.edge_index2)


class UpConv(nn.Module):
    """util/meshnet.py:94-160: conv-MeshUnpool-BN-act on the coarse graph, then 4 x conv-BN-act + Dropout on the fine one."""

    def __init__(self, in_channels, out_channels, edge_index1, edge_index2, unpool_hash, K=3, drop_rate=0.0, conv="chebconv"):
        super().__init__()
        self.edge_index1, self.edge_index2 = edge_index1, edge_index2
        self.model1 = Sequential("x, edge_index", _block(conv, in_channels, out_channels, K, MeshUnpool(unpool_hash)))
        self.model2 = Sequential("x, edge_index", _block(conv, out_channels, out_channels, K) + _block(conv, out_channels, out_channels, K)
                                 + _block(conv, out_channels, out_channels, K) + _block(conv, out_channels, out_channels, K)
                                 + [(nn.Dropout(drop_rate), "x -> x")])

    def forward(self, input: Tensor) -> Tensor:
        return self.model2(self.model1(input, self.edge_index1), self.edge_index2)


class _LinearOnGemm(nn.Linear):
    """nn.Linear with the same parameters / init, forward on our GEMM tiles (the reference uses plain nn.Linear inside
    nn.Sequential and as skip layers, util/meshnet.py:219-221,247-248)."""

    def forward(self, input: Tensor) -> Tensor:
        return ops.linear(input, self.weight, self.bias)


def _head(cin: int, K: int, conv: str) -> Sequential:
    return Sequential("x, edge_index", [(_conv(conv, cin, 32, K), "x, edge_index -> x"), (nn.BatchNorm1d(32), "x -> x"),
                                         (nn.LeakyReLU(), "x -> x"), (nn.Linear(32, 3), "x -> x")])


class MGCN(nn.Module):
    """util/meshnet.py:163-317 given a finished hierarchy (3 pooling levels).  Construction order = the reference's
    (encoder1-3, decoder3-1, mcnn3-1, skip2, skip1) => the same parameters under the same seed.  Dropout rates as the
    reference (0.2 inside encoder2/3, decoder3/2); parity tests run in eval mode or with ``drop_rate=0`` (torch's Philox
    mask stream is not reproducible by construction order alone)."""

    def __init__(self, device, edge_inds: Sequence[Tensor], p_hashes: Sequence[Tensor], up_hashes: Sequence[Tensor],
                 smposs_list: Sequence[Tensor], K: int = 3, skip: bool = False, conv: str = "chebconv", drop_rate: float = 0.2,
                 tensor_masks: bool = False):
        super().__init__()
        if len(edge_inds) != 4 or len(p_hashes) != 3 or len(up_hashes) != 3 or len(smposs_list) != 4:
            raise ValueError("MGCN: expected 4 graphs, 3 pool / unpool matrices and 4 smoothed position arrays")
        self.device, self.skip = device, skip
        # The reference replaces every mask that is not an np.ndarray -- torch tensors included -- by ones
        # (util/meshnet.py:290-293).  tensor_masks=True is the extension that honours device tensors (CUDA-graph replay).
        self.tensor_masks = tensor_masks
        self.edge_inds = [e.to(device) for e in edge_inds]
        self.smposs_list = [p.to(device).float() for p in smposs_list]
        e, d = self.edge_inds, drop_rate
        self.encoder1 = DownConv(4, 32, e[0], e[1], p_hashes[0], K=K, drop_rate=0.0, conv=conv)
        self.encoder2 = DownConv(32, 128, e[1], e[2], p_hashes[1], K=K, drop_rate=d, conv=conv)
        self.encoder3 = DownConv(128, 256, e[2], e[3], p_hashes[2], K=K, drop_rate=d, conv=conv)
        self.decoder3 = UpConv(256, 128, e[3], e[2], up_hashes[2], K=K, drop_rate=d, conv=conv)
        self.decoder2 = UpConv(128, 32, e[2], e[1], up_hashes[1], K=K, drop_rate=d, conv=conv)
        self.decoder1 = nn.Sequential(UpConv(32, 16, e[1], e[0], up_hashes[0], K=K, drop_rate=0.0, conv=conv), _LinearOnGemm(16, 3))
        self.mcnn3, self.mcnn2, self.mcnn1 = _head(256, K, conv), _head(128, K, conv), _head(32, K, conv)
        self.skip2 = _LinearOnGemm(256, 128)
        self.skip1 = _LinearOnGemm(64, 32)

    def forward(self, data, dm=None):
        z1 = data.z1.to(self.device)
        if type(dm) == np.ndarray:
            dm = torch.from_numpy(dm)
        elif not (self.tensor_masks and isinstance(dm, torch.Tensor)):
            dm = torch.ones([z1.shape[0], 1])
        dm = dm.to(self.device).to(z1.dtype)
        if z1.is_cuda and z1.dtype == torch.float32 and z1.shape[1] == 3 and not z1.requires_grad and dm.numel() == z1.shape[0]:
            z1 = ops.input_prep(z1, dm)             # util/meshnet.py:282-293 as two kernels, bit-identical (csrc/prep.cu)
        else:
            z_min, z_max = torch.min(z1, dim=0, keepdim=True)[0], torch.max(z1, dim=0, keepdim=True)[0]
            z_sc = torch.max(z_max - z_min)
            zc = (z_min + z_max) * 0.5
            z1 = (z1 - zc) / z_sc
            z1 = torch.cat([dm * z1[:, 0:3], dm], dim=1)
        res1_enc = self.encoder1(z1)
        res2_enc = self.encoder2(res1_enc)
        res3_bot = self.encoder3(res2_enc)
        out3 = self.mcnn3(res3_bot, self.edge_inds[3])
        res2_dec = self.decoder3(res3_bot)
        if self.skip:
            res2_dec = self.skip2(torch.cat([res2_dec, res2_enc], dim=1))
        out2 = self.mcnn2(res2_dec, self.edge_inds[2])
        res1_dec = self.decoder2(res2_dec)
        if self.skip:
            res1_dec = self.skip1(torch.cat([res1_dec, res1_enc], dim=1))
        out1 = self.mcnn1(res1_dec, self.edge_inds[1])
        out0 = self.decoder1(res1_dec)
        p = self.smposs_list
        return (p[0] + out0, p[1] + out1, p[2] + out2, p[3] + out3)
