"""Drop-in replacements for the three torch_geometric symbols SeMIGCN imports
(``from torch_geometric.nn import GCNConv, ChebConv, Sequential`` -- reference
util/networks.py:4, util/meshnet.py:6): same constructors, ``forward(x, edge_index)``
signature, parameter layout and ``state_dict`` keys (``lin.weight``/``bias``,
``lins.{k}.weight``/``bias``, ``module_{i}``), so sgcn.py / mgcn.py run unchanged
(SURVEY.md §8(b), Appendix A).

``Sequential`` additionally recognises the reference's block pattern
``conv -> BatchNorm1d -> LeakyReLU`` (util/networks.py:24-28,41-45) and runs it as one fused
autograd node: BatchNorm statistics come out of the SpMM/GEMM epilogue, the weight
gradient is a split-m GEMM, the aggregation backward is a gather over the transpose CSR.
"""
from __future__ import annotations

import math
from typing import List, Optional, Tuple

import torch
import torch.nn as nn
from torch import Tensor

from . import ops
from ._lib import MODE_CHEB, MODE_GCN, SgbError, require_cuda


def _glorot_(w: Tensor) -> None:
    a = math.sqrt(6.0 / (w.size(-2) + w.size(-1)))
    w.data.uniform_(-a, a)


class _PygLinear(nn.Module):
    """Weight holder with PyG's ``Linear(in, out, bias=False, weight_initializer='glorot')``
    naming and RNG consumption (one uniform draw at construction)."""

    def __init__(self, in_channels: int, out_channels: int):
        super().__init__()
        self.in_channels, self.out_channels = in_channels, out_channels
        self.weight = nn.Parameter(torch.empty(out_channels, in_channels))
        self.register_parameter("bias", None)
        self.reset_parameters()

    def reset_parameters(self) -> None:
        _glorot_(self.weight)

    def forward(self, x: Tensor) -> Tensor:
        return ops.linear(x, self.weight, None)


def _check_inputs(x: Tensor, edge_index: Tensor, in_channels: int) -> None:
    require_cuda(x, edge_index)
    if x.dim() != 2 or x.shape[1] != in_channels:
        raise SgbError(f"expected x of shape [N, {in_channels}], got {tuple(x.shape)}")
    if x.dtype != torch.float32:
        raise SgbError("x must be float32")


class GCNConv(nn.Module):
    """torch_geometric.nn.GCNConv as the reference constructs it (``GCNConv(in, out)``)."""

    def __init__(self, in_channels: int, out_channels: int, improved: bool = False, cached: bool = False,
                 add_self_loops: bool = True, normalize: bool = True, bias: bool = True, **kwargs):
        super().__init__()
        if improved or not add_self_loops or not normalize:
            raise SgbError("GCNConv: only improved=False, add_self_loops=True, normalize=True are supported")
        if kwargs:
            raise SgbError(f"GCNConv: unsupported arguments {sorted(kwargs)}")
        self.in_channels, self.out_channels = in_channels, out_channels
        self.improved, self.cached, self.add_self_loops, self.normalize = improved, cached, add_self_loops, normalize
        self.lin = _PygLinear(in_channels, out_channels)
        if bias:
            self.bias = nn.Parameter(torch.empty(out_channels))
        else:
            self.register_parameter("bias", None)
        self.reset_parameters()

    def reset_parameters(self) -> None:
        self.lin.reset_parameters()
        if self.bias is not None:
            self.bias.data.zero_()

    def forward(self, x: Tensor, edge_index: Tensor, edge_weight: Optional[Tensor] = None) -> Tensor:
        if edge_weight is not None:
            raise SgbError("GCNConv: edge_weight is not supported (the reference never passes one)")
        _check_inputs(x, edge_index, self.in_channels)
        g = ops.graph_for(edge_index, x.shape[0], MODE_GCN)
        return ConvBlockFn.apply(x, g, _BlockCfg("gcn", None, 1.0), self.bias, None, None, self.lin.weight)

    def __repr__(self) -> str:
        return f"GCNConv({self.in_channels}, {self.out_channels})"


class ChebConv(nn.Module):
    """torch_geometric.nn.ChebConv as the reference constructs it (``ChebConv(in, out, K=3)``)."""

    def __init__(self, in_channels: int, out_channels: int, K: int, normalization: Optional[str] = "sym",
                 bias: bool = True, **kwargs):
        super().__init__()
        if K <= 0:
            raise SgbError("ChebConv: K must be positive")
        if normalization != "sym":
            raise SgbError("ChebConv: only normalization='sym' is supported")
        if kwargs:
            raise SgbError(f"ChebConv: unsupported arguments {sorted(kwargs)}")
        self.in_channels, self.out_channels, self.normalization = in_channels, out_channels, normalization
        self.lins = nn.ModuleList([_PygLinear(in_channels, out_channels) for _ in range(K)])
        if bias:
            self.bias = nn.Parameter(torch.empty(out_channels))
        else:
            self.register_parameter("bias", None)
        self.reset_parameters()

    def reset_parameters(self) -> None:
        for lin in self.lins:
            lin.reset_parameters()
        if self.bias is not None:
            self.bias.data.zero_()

    def forward(self, x: Tensor, edge_index: Tensor, edge_weight: Optional[Tensor] = None,
                batch: Optional[Tensor] = None, lambda_max: Optional[Tensor] = None) -> Tensor:
        if edge_weight is not None or batch is not None or lambda_max is not None:
            raise SgbError("ChebConv: edge_weight / batch / lambda_max are not supported (lambda_max is 2.0 for 'sym')")
        _check_inputs(x, edge_index, self.in_channels)
        g = ops.graph_for(edge_index, x.shape[0], MODE_CHEB)
        return ConvBlockFn.apply(x, g, _BlockCfg("cheb", None, 1.0), self.bias, None, None, *[l.weight for l in self.lins])

    def __repr__(self) -> str:
        return f"ChebConv({self.in_channels}, {self.out_channels}, K={len(self.lins)}, normalization={self.normalization})"


# ----------------------------------------------------------------------------------------
# fused block: conv (+ BatchNorm1d + LeakyReLU)
# ----------------------------------------------------------------------------------------
class _BlockCfg:
    def __init__(self, kind: str, bn: Optional[nn.BatchNorm1d], slope: float, x_amax: Optional[Tensor] = None):
        self.kind, self.bn, self.slope = kind, bn, slope
        self.x_amax = x_amax          # in:  device scalar max|x| if the producer of x published one (ops.amax_of)
        self.out_amax = None          # out: device scalar max|z| of the block output (set by ConvBlockFn.forward)


class ConvBlockFn(torch.autograd.Function):
    """One autograd node for ``act(BN(conv(x)))`` (BN/act optional).

    GCN : aggregate-first when Cin <= Cout (P = S x, Y = P W^T + b), transform-first
          otherwise (H = x W^T, Y = S H + b): the SpMM always runs at the narrower width.
    Cheb: T0 = x, T1 = S x, Tk = 2 S T(k-1) - T(k-2);  Y = sum_k Tk Wk^T + b.
    BatchNorm statistics are accumulated in the epilogue of the kernel that writes Y.
    """

    @staticmethod
    def forward(ctx, x, graph, cfg, bias, gamma, beta, *weights):
        x = ops._f32c(x, "x")
        bn = cfg.bn
        training = bn is not None and (bn.training or bn.running_mean is None)
        want_stats = training
        saved_ops: List[Tensor] = []
        # device scalars max|operand| for the fp16-split GEMM engine, published by the kernels that write the operands
        am = ops.new_amax(x.device, 4)
        x_amax = cfg.x_amax
        if cfg.kind == "gcn":
            (w,) = weights
            cout, cin = w.shape
            agg_first = cin <= cout
            if agg_first:
                p = ops.spmm(graph, x, amax_out=am[0:1])
                r = ops.gemm(p, w, transb=True, bias=bias, want_stats=want_stats, a_amax=am[0:1])
                saved_ops = [p]
                ctx.op_amax = [am[0:1]]
            else:
                h = ops.gemm(x, w, transb=True, a_amax=x_amax)
                r = ops.spmm(graph, h, bias=bias, want_stats=want_stats)
                saved_ops = [x]
                ctx.op_amax = [x_amax]
            ctx.agg_first = agg_first
        else:
            # T1 .. T(K-1) live side by side in ONE [m, (K-1) cin] matrix: Y = x W0^T + [T1 | T2 | ..] [W1 | W2 | ..]^T is then two
            # transforms instead of K accumulating ones (each of those re-reads and re-writes Y), and the backward pass reads dY
            # twice instead of K times for the weight gradients and for dT.  The aggregation kernels write / read the column blocks
            # in place (row stride (K-1) cin); max|.| of the whole block accumulates in one slot over the K - 1 launches.
            nw = len(weights)
            cin = x.shape[1]
            tcat = None
            if nw > 1:
                tcat = torch.empty((x.shape[0], (nw - 1) * cin), dtype=torch.float32, device=x.device)
                tk = [x] + [tcat[:, (k - 1) * cin:k * cin] for k in range(1, nw)]
                ops.spmm(graph, x, out=tk[1], amax_out=am[0:1])
                for k in range(2, nw):
                    ops.spmm(graph, tk[k - 1], alpha=2.0, addend=tk[k - 2], beta=-1.0, out=tk[k], amax_out=am[0:1])
            one = nw == 1
            r = ops.gemm(x, weights[0], transb=True, bias=bias if one else None, want_stats=want_stats and one, a_amax=x_amax)
            if not one:
                wcat = weights[1] if nw == 2 else torch.cat(list(weights[1:]), dim=1)
                r = ops.gemm(tcat, wcat, transb=True, bias=bias, out=r, accumulate=True, want_stats=want_stats, a_amax=am[0:1])
            saved_ops = [x] if one else [x, tcat]
            ctx.op_amax = [x_amax, am[0:1]]
        y, partials = r if want_stats else (r, None)

        ctx.graph, ctx.cfg, ctx.nw, ctx.has_bias = graph, cfg, len(weights), bias is not None
        if bn is None:
            ctx.bn_mode = 0
            ctx.save_for_backward(*weights, *saved_ops)
            return y
        if training:
            momentum = 0.1 if bn.momentum is None else bn.momentum
            if bn.track_running_stats and bn.num_batches_tracked is not None:
                bn.num_batches_tracked.add_(1)
                if bn.momentum is None:
                    momentum = 1.0 / float(bn.num_batches_tracked)
            rm = bn.running_mean if bn.track_running_stats else None
            rv = bn.running_var if bn.track_running_stats else None
            mean, invstd, scale, shift = ops.bn_finalize(partials, graph.n_global, gamma, beta, bn.eps, momentum, rm, rv,
                                                         comm=graph.comm)
        else:
            mean, invstd, scale, shift = ops.eval_affine(bn.running_mean, bn.running_var, gamma, beta, bn.eps)
        z = ops.bn_act_apply(y, mean, scale, shift, cfg.slope, amax_out=am[3:4])
        cfg.out_amax = am[3:4]
        ctx.bn_mode = 2 if training else 1
        ctx.has_affine = gamma is not None
        ctx.save_for_backward(*weights, *saved_ops, y, scale, shift, mean, invstd)
        return z

    @staticmethod
    def backward(ctx, dz):
        cfg, graph, nw = ctx.cfg, ctx.graph, ctx.nw
        saved = ctx.saved_tensors
        weights = saved[:nw]
        dgamma = dbeta = None
        dz = dz.contiguous()
        am = ops.new_amax(dz.device, 2)
        dy_amax = None
        if ctx.bn_mode == 0:
            saved_ops = saved[nw:]
            dy = dz
        else:
            saved_ops = saved[nw:-5]
            y, scale, shift, mean, invstd = saved[-5:]
            dy_amax = am[0:1]
            dy, dgamma, dbeta = ops.bn_act_bwd(dz, y, scale, shift, mean, invstd, cfg.slope, ctx.bn_mode == 2,
                                               comm=graph.comm, n_global=graph.n_global, amax_out=dy_amax)
            if not ctx.has_affine:
                dgamma = dbeta = None
        need_x = ctx.needs_input_grad[0]
        db = None
        want_db = ctx.has_bias and ctx.needs_input_grad[3]
        # a bare conv (no BatchNorm behind it) receives a dY none of our kernels produced: its max |dY| (the scale of the
        # fp16-split GEMMs) is reduced ONCE -- in the bias-gradient pass when there is one -- and shared by the dW and
        # dX GEMMs, instead of each GEMM running its own reduction pass over dY
        dy_used_by_gemm = cfg.kind != "gcn" or ctx.agg_first
        if dy_amax is None and dy_used_by_gemm:
            if want_db and ctx.bn_mode != 2:
                dy_amax = am[0:1]
                db = ops.colsum(dy, amax_out=dy_amax)
                want_db = False
            elif ops.tensor_core_likely(dy.shape[0], weights[0].shape[1], dy.shape[1]):
                dy_amax = ops.amax(dy)
        if want_db:
            if ctx.bn_mode == 2:
                # training-mode BatchNorm removes the batch mean, so sum_rows dY is identically zero (dY = scale *
                # (dA - mean(dA) - xhat * mean(dA * xhat)) and sum xhat = 0): the gradient of a conv bias in front of
                # BatchNorm is exactly 0 (the reference holds fp32 rounding noise there).  No reduction pass needed.
                db = torch.zeros(dy.shape[1], dtype=dy.dtype, device=dy.device)
            else:
                db = ops.colsum(dy)
        dws: List[Optional[Tensor]] = [None] * nw
        dx = None
        if cfg.kind == "gcn":
            (w,) = weights
            if ctx.agg_first:
                (p,) = saved_ops
                dws[0] = ops.gemm_tn(dy, p, g_amax=dy_amax, a_amax=ctx.op_amax[0])
                if need_x:
                    dp = ops.gemm(dy, w, transb=False, a_amax=dy_amax)
                    dx = ops.spmm(graph, dp, transpose=True)
            else:
                (x,) = saved_ops
                dh = ops.spmm(graph, dy, transpose=True, amax_out=am[1:2])
                dws[0] = ops.gemm_tn(dh, x, g_amax=am[1:2], a_amax=ctx.op_amax[0])
                if need_x:
                    dx = ops.gemm(dh, w, transb=False, a_amax=am[1:2])
        else:
            x = saved_ops[0]
            cin = x.shape[1]
            dws[0] = ops.gemm_tn(dy, x, g_amax=dy_amax, a_amax=ctx.op_amax[0])
            if nw > 1:
                tcat = saved_ops[1]
                dwcat = ops.gemm_tn(dy, tcat, g_amax=dy_amax, a_amax=ctx.op_amax[1])          # [cout, (K-1) cin]
                for k in range(1, nw):
                    dws[k] = dwcat[:, (k - 1) * cin:k * cin]
            if need_x:
                # g_k = dY W_k + a_k S^T g_(k+1) - g_(k+2),  a_k = 2 for k >= 1, 1 for k = 0;  dY [W1 | W2 | ..] in one transform
                ds: List[Optional[Tensor]] = [ops.gemm(dy, weights[0], transb=False, a_amax=dy_amax)]
                if nw > 1:
                    wcat = weights[1] if nw == 2 else torch.cat(list(weights[1:]), dim=1)
                    dcat = ops.gemm(dy, wcat, transb=False, a_amax=dy_amax)
                    ds += [dcat[:, (k - 1) * cin:k * cin] for k in range(1, nw)]
                gs: List[Optional[Tensor]] = [None] * nw
                for k in range(nw - 1, -1, -1):
                    d = ds[k]
                    if k + 2 <= nw - 1:
                        d = torch.sub(d, gs[k + 2], out=d)
                    if k + 1 <= nw - 1:
                        d = ops.spmm(graph, gs[k + 1], transpose=True, alpha=2.0 if k >= 1 else 1.0, addend=d, beta=1.0)
                    gs[k] = d
                dx = gs[0]
        return (dx, None, None, db, dgamma, dbeta, *dws)


# ----------------------------------------------------------------------------------------
# MeshPool / MeshUnpool (reference util/meshnet.py:9-27), on the SpMM kernel
# ----------------------------------------------------------------------------------------
class _SparseHashModule(nn.Module):
    """Holds the reference's sparse "hash" matrix as a buffer of the same name (so ``state_dict`` / ``.to(device)`` behave
    as in util/meshnet.py) and, lazily per device, the CSR pair the SpMM kernel reads."""

    _buffer_name = "hash"

    def __init__(self, hash_matrix: Tensor):
        super().__init__()
        self.register_buffer(self._buffer_name, hash_matrix)
        self._op = None
        self._op_key = None

    def _row_scale(self, mat: Tensor) -> Optional[Tensor]:
        return None

    def _operator(self, device) -> "ops.SparseOp":
        mat = getattr(self, self._buffer_name)
        key = (str(device), mat._version if not mat.is_sparse else id(mat))
        if self._op is None or self._op_key != key:
            m = mat.to(device)
            self._op = ops.SparseOp(m, self._row_scale(m))
            self._op_key = key
        return self._op

    def forward(self, input: Tensor) -> Tensor:
        require_cuda(input)
        if input.dtype != torch.float32 or input.dim() != 2:
            raise SgbError(f"{type(self).__name__}: expected a float32 [N, C] tensor")
        return ops.sparse_mm(self._operator(input.device), input)


class MeshPool(_SparseHashModule):
    """util/meshnet.py:9-17: ``sparse.mm(pool_hash, x) / rowsum(pool_hash)`` -- the average of the fine vertices merged into
    each coarse vertex.  The reference densifies ``pool_hash`` [n_coarse, n_fine] on every forward to get the row sums;
    here they are folded into the CSR weights once (w = value / rowsum, so the result differs from the reference by the
    rounding of that one division, ~1e-7 relative).  A coarse vertex with an empty row gives 0 (reference: NaN)."""

    _buffer_name = "pool_hash"

    def __init__(self, pool_hash: Tensor):
        super().__init__(pool_hash)

    def _row_scale(self, mat: Tensor) -> Tensor:
        m = mat if mat.is_sparse else mat.to_sparse()
        v_sum = torch.sparse.sum(m, dim=1).to_dense().to(torch.float32)
        return torch.where(v_sum != 0, 1.0 / v_sum, torch.zeros_like(v_sum))


class MeshUnpool(_SparseHashModule):
    """util/meshnet.py:20-27: ``sparse.mm(unpool_hash, x)`` -- every fine vertex takes the row of its coarse vertex."""

    _buffer_name = "unpool_hash"

    def __init__(self, unpool_hash: Tensor):
        super().__init__(unpool_hash)


class Sequential(nn.Module):
    """torch_geometric.nn.Sequential(input_args, modules): children are registered as
    ``module_{i}``; entries are ``(module, "a, b -> c")`` or bare modules applied to the
    previous output."""

    def __init__(self, input_args: str, modules: list):
        super().__init__()
        self._in_names = [a.strip() for a in input_args.split(",")]
        self._descs: List[Tuple[List[str], List[str]]] = []
        for i, entry in enumerate(modules):
            if isinstance(entry, (tuple, list)):
                mod, desc = entry
                lhs, rhs = desc.split("->")
                ins = [a.strip() for a in lhs.split(",")]
                outs = [a.strip() for a in rhs.split(",")]
            else:
                mod = entry
                prev = self._descs[-1][1] if self._descs else self._in_names[:1]
                ins, outs = list(prev), list(prev)
            if isinstance(mod, nn.Module):
                setattr(self, f"module_{i}", mod)
            else:
                object.__setattr__(self, f"module_{i}", mod)     # plain callable
            self._descs.append((ins, outs))

    def __len__(self) -> int:
        return len(self._descs)

    def __getitem__(self, i: int):
        return getattr(self, f"module_{i}")

    def _fusable(self, i: int) -> Optional[Tuple[int, float]]:
        """If entries i.. form conv -> BatchNorm1d [-> LeakyReLU | ReLU], return (entries consumed, slope)."""
        conv = self[i]
        if not isinstance(conv, (GCNConv, ChebConv)) or i + 1 >= len(self):
            return None
        ins, outs = self._descs[i]
        if len(ins) != 2 or len(outs) != 1:
            return None
        bn = self[i + 1]
        if type(bn) is not nn.BatchNorm1d or self._descs[i + 1][0] != outs:
            return None
        if i + 2 < len(self) and self._descs[i + 2][0] == outs:
            act = self[i + 2]
            if type(act) is nn.LeakyReLU:
                return 3, float(act.negative_slope)
            if type(act) is nn.ReLU:
                return 3, 0.0
        return 2, 1.0

    def forward(self, *args):
        env = dict(zip(self._in_names, args))
        out = None
        i = 0
        while i < len(self):
            ins, outs = self._descs[i]
            mod = self[i]
            fuse = self._fusable(i)
            if fuse is not None:
                consumed, slope = fuse
                x, edge_index = env[ins[0]], env[ins[1]]
                bn = self[i + 1]
                _check_inputs(x, edge_index, mod.in_channels)
                if isinstance(mod, GCNConv):
                    g = ops.graph_for(edge_index, x.shape[0], MODE_GCN)
                    cfg = _BlockCfg("gcn", bn, slope, ops.amax_of(x))
                    out = ConvBlockFn.apply(x, g, cfg, mod.bias, bn.weight, bn.bias, mod.lin.weight)
                else:
                    g = ops.graph_for(edge_index, x.shape[0], MODE_CHEB)
                    cfg = _BlockCfg("cheb", bn, slope, ops.amax_of(x))
                    out = ConvBlockFn.apply(x, g, cfg, mod.bias, bn.weight, bn.bias, *[l.weight for l in mod.lins])
                if cfg.out_amax is not None:
                    ops.tag_amax(out, cfg.out_amax)      # the next block's transform reads it (same tensor object, same version)
                env[outs[0]] = out
                i += consumed
                continue
            vals = [env[k] for k in ins]
            if type(mod) is nn.Linear and len(vals) == 1 and vals[0].is_cuda and vals[0].dim() == 2:
                out = ops.linear(vals[0], mod.weight, mod.bias)
            elif (type(mod) is nn.BatchNorm1d and len(vals) == 1 and vals[0].is_cuda and vals[0].dim() == 2
                  and vals[0].dtype == torch.float32 and len(outs) == 1):
                # BatchNorm1d [-> LeakyReLU | ReLU] behind something that is not a conv (conv -> MeshPool -> BN -> act,
                # util/meshnet.py:44-47,106-109): our statistics / apply / backward kernels, one autograd node
                slope, consumed = 1.0, 1
                if i + 1 < len(self) and self._descs[i + 1][0] == outs:
                    act = self[i + 1]
                    if type(act) is nn.LeakyReLU:
                        slope, consumed = float(act.negative_slope), 2
                    elif type(act) is nn.ReLU:
                        slope, consumed = 0.0, 2
                out = ops.bn_act(vals[0], mod, slope)
                env[outs[0]] = out
                i += consumed
                continue
            else:
                out = mod(*vals)
            if len(outs) == 1:
                env[outs[0]] = out
            else:
                for k, v in zip(outs, out):
                    env[k] = v
            i += 1
        return out
