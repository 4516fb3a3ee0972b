"""ORACLE (test infrastructure only -- never imported by the product): CPU restatement of the reference's
multi-resolution blocks, util/meshnet.py, in plain torch ops on top of oracle/pyg_ref.py.

  MeshPool    util/meshnet.py:9-17    out = sparse.mm(pool_hash, x) / rowsum(pool_hash.to_dense())
  MeshUnpool  util/meshnet.py:20-27   out = sparse.mm(unpool_hash, x)
  DownConv    util/meshnet.py:31-91   (the live "chebconv" branch and the dead "gcnconv" one)
  UpConv      util/meshnet.py:94-160
  MGCN        util/meshnet.py:212-317 (module tree + forward; the hierarchy -- which the reference builds by QEM
              simplification of its Mesh objects, :169-193 -- is an argument)

Pinned by tests/golden/ref_meshnet_n4.npz, which tests/golden/make_golden_meshnet.py generates by executing the
reference's own util/meshnet.py (classes MeshPool / MeshUnpool / DownConv / UpConv and the unbound MGCN.forward) with
the torch_geometric symbols resolved to oracle/pyg_ref.py.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn as nn

from . import pyg_ref as O


class MeshPool(nn.Module):
    def __init__(self, pool_hash):
        super().__init__()
        self.register_buffer("pool_hash", pool_hash)

    def forward(self, input):
        v_sum = torch.sum(self.pool_hash.to_dense(), dim=1, keepdim=True)
        return torch.sparse.mm(self.pool_hash, input) / v_sum


class MeshUnpool(nn.Module):
    def __init__(self, unpool_hash):
        super().__init__()
        self.register_buffer("unpool_hash", unpool_hash)

    def forward(self, input):
        return torch.sparse.mm(self.unpool_hash, input)


def _conv(kind, cin, cout, K):
    return O.ChebConv(cin, cout, K=K) if kind == "chebconv" else O.GCNConv(cin, cout)


def _block(kind, cin, cout, K, mid=None):
    mods = [(_conv(kind, cin, cout, K), "x, edge_index -> x")]
    if mid is not None:
        mods.append((mid, "x -> x"))
    return mods + [(nn.BatchNorm1d(cout), "x -> x"), (nn.LeakyReLU(), "x -> x")]


class DownConv(nn.Module):
    def __init__(self, in_channels, out_channels, edge_index1, edge_index2, pool_hash, K=3, drop_rate=0.0, conv="chebconv"):
        super().__init__()
        self.edge_index1, self.edge_index2 = edge_index1, edge_index2
        c = out_channels
        self.model1 = O.Sequential("x, edge_index", _block(conv, in_channels, c, K) + _block(conv, c, c, K, MeshPool(pool_hash)))
        self.model2 = O.Sequential("x, edge_index", _block(conv, c, c, K) + _block(conv, c, c, K) + _block(conv, c, c, K)
                                   + [(nn.Dropout(drop_rate), "x -> x")])

    def forward(self, input):
        return self.model2(self.model1(input, self.edge_index1), self.edge_index2)


class UpConv(nn.Module):
    def __init__(self, in_channels, out_channels, edge_index1, edge_index2, unpool_hash, K=3, drop_rate=0.0, conv="chebconv"):
        super().__init__()
        self.edge_index1, self.edge_index2 = edge_index1, edge_index2
        c = out_channels
        self.model1 = O.Sequential("x, edge_index", _block(conv, in_channels, c, K, MeshUnpool(unpool_hash)))
        self.model2 = O.Sequential("x, edge_index", _block(conv, c, c, K) + _block(conv, c, c, K) + _block(conv, c, c, K)
                                   + _block(conv, c, c, K) + [(nn.Dropout(drop_rate), "x -> x")])

    def forward(self, input):
        return self.model2(self.model1(input, self.edge_index1), self.edge_index2)


def _head(cin, K, conv):
    return O.Sequential("x, edge_index", [(_conv(conv, cin, 32, K), "x, edge_index -> x"), (nn.BatchNorm1d(32), "x -> x"),
                                           (nn.LeakyReLU(), "x -> x"), (nn.Linear(32, 3), "x -> x")])


class MGCN(nn.Module):
    def __init__(self, edge_inds, p_hashes, up_hashes, smposs_list, K=3, skip=False, conv="chebconv", drop_rate=0.2):
        super().__init__()
        self.skip, self.edge_inds, self.smposs_list = skip, list(edge_inds), [p.float() for p in smposs_list]
        e, d = self.edge_inds, drop_rate
        self.encoder1 = DownConv(4, 32, e[0], e[1], p_hashes[0], K=K, drop_rate=0.0, conv=conv)
        self.encoder2 = DownConv(32, 128, e[1], e[2], p_hashes[1], K=K, drop_rate=d, conv=conv)
        self.encoder3 = DownConv(128, 256, e[2], e[3], p_hashes[2], K=K, drop_rate=d, conv=conv)
        self.decoder3 = UpConv(256, 128, e[3], e[2], up_hashes[2], K=K, drop_rate=d, conv=conv)
        self.decoder2 = UpConv(128, 32, e[2], e[1], up_hashes[1], K=K, drop_rate=d, conv=conv)
        self.decoder1 = nn.Sequential(UpConv(32, 16, e[1], e[0], up_hashes[0], K=K, drop_rate=0.0, conv=conv), nn.Linear(16, 3))
        self.mcnn3, self.mcnn2, self.mcnn1 = _head(256, K, conv), _head(128, K, conv), _head(32, K, conv)
        self.skip2 = nn.Linear(256, 128)
        self.skip1 = nn.Linear(64, 32)

    def forward(self, z1, dm=None, tensor_masks=False):
        z_min, z_max = torch.min(z1, dim=0, keepdim=True)[0], torch.max(z1, dim=0, keepdim=True)[0]
        z_sc = torch.max(z_max - z_min)
        zc = (z_min + z_max) * 0.5
        z1 = (z1 - zc) / z_sc
        if type(dm) == np.ndarray:
            dm = torch.from_numpy(dm)
        elif not (tensor_masks and isinstance(dm, torch.Tensor)):      # util/meshnet.py:290-293: anything but an ndarray -> ones
            dm = torch.ones([z1.shape[0], 1])
        dm = dm.to(z1.dtype)
        z1 = torch.cat([dm * z1[:, 0:3], dm], dim=1)
        r1 = self.encoder1(z1)
        r2 = self.encoder2(r1)
        r3 = self.encoder3(r2)
        out3 = self.mcnn3(r3, self.edge_inds[3])
        d2 = self.decoder3(r3)
        if self.skip:
            d2 = self.skip2(torch.cat([d2, r2], dim=1))
        out2 = self.mcnn2(d2, self.edge_inds[2])
        d1 = self.decoder2(d2)
        if self.skip:
            d1 = self.skip1(torch.cat([d1, r1], dim=1))
        out1 = self.mcnn1(d1, self.edge_inds[1])
        out0 = self.decoder1(d1)
        p = self.smposs_list
        return (p[0] + out0, p[1] + out1, p[2] + out2, p[3] + out3)
