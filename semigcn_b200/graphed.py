"""Whole-train-step CUDA graphs for the launch-bound regime (the reference's own mesh sizes).

On the meshes SeMIGCN actually trains on (10-30 k vertices, BASELINE.json configs[0]/[1]) and on
the 100 k-vertex meshes of configs[4], one SGCN step is ~270 short kernels: the GPU work is
1-3 ms while the Python / launch path costs ~6.7 ms (measured, B200 box).  Every kernel behind the
C-ABI runs on the caller's stream with no host synchronisation (include/semigcn_b200.h), the CSR of
an ``edge_index`` is cached, and the step losses never read a value back -- so the step
``zero_grad -> posnet(data, dm) -> losses -> backward -> optimizer.step`` (sgcn.py:125-146) is
capturable as ONE CUDA graph and replayed with new masks / inputs copied into static buffers.

    step = GraphedTrainStep(net, loss_fn, opt, z1, x_pos, edge_index, dm0)
    for dm in masks:                      # sgcn.py:112-146
        loss = step(dm)                   # device scalar; loss.item() only when it is logged

The optimizer must be capture-safe (``torch.optim.Adam(..., capturable=True)``).  The graph holds the
``edge_index`` it was captured with (same mesh every step, as in the reference); a different mesh
needs its own ``GraphedTrainStep``.
"""
from __future__ import annotations

from typing import Callable, Optional

import torch
from torch import Tensor

from ._lib import SgbError, require_cuda
from .data import Data


class GraphedTrainStep:
    def __init__(self, net: torch.nn.Module, loss_fn: Callable[[Tensor], Tensor], optimizer: torch.optim.Optimizer,
                 z1: Tensor, x_pos: Tensor, edge_index: Tensor, dm: Tensor, warmup: int = 3, accumulate: int = 1):
        """``loss_fn(out) -> scalar`` closes over its (static) targets.  ``accumulate``: optimizer steps every that many
        calls (sgcn.py:40,146 steps every 5 masks): the graph then holds forward + backward only when it is > 1."""
        require_cuda(z1, x_pos, dm)          # edge_index may be None for networks that hold their graphs (MGCN)
        for g in optimizer.param_groups:
            if not g.get("capturable", False) and accumulate == 1:
                raise SgbError("GraphedTrainStep: the optimizer must be built with capturable=True")
        self.net, self.loss_fn, self.opt = net, loss_fn, optimizer
        self.accumulate, self._calls = int(accumulate), 0
        self.z1, self.x_pos, self.dm = z1.clone(), x_pos.clone(), dm.clone().to(torch.float32)
        self.data = Data(z1=self.z1, x_pos=self.x_pos, edge_index=edge_index)
        snap = self._snapshot()
        side = torch.cuda.Stream(device=z1.device)
        side.wait_stream(torch.cuda.current_stream(z1.device))
        with torch.cuda.stream(side):             # warm-up off the capture stream: CSR / topology caches, allocator pools, Adam state
            for _ in range(max(1, warmup)):
                self._eager_step(step_opt=accumulate == 1)
        torch.cuda.current_stream(z1.device).wait_stream(side)
        torch.cuda.synchronize(z1.device)
        if accumulate > 1:
            optimizer.zero_grad(set_to_none=False)
        from . import _lib
        self.graph = torch.cuda.CUDAGraph()
        n0 = _lib.LAUNCHES
        with torch.cuda.graph(self.graph):
            self.loss = self._eager_step(step_opt=accumulate == 1)
        self.launches_per_replay = _lib.LAUNCHES - n0     # kernels of ours inside one replay
        self.out = self._out
        self._restore(snap)                       # warm-up and capture must not count as training steps

    def _snapshot(self):
        params = [(p, p.detach().clone()) for p in self.net.parameters()]
        bufs = [(b, b.detach().clone()) for b in self.net.buffers()]
        state = {}
        for p, st in self.opt.state.items():
            state[p] = {k: (v.detach().clone() if torch.is_tensor(v) else v) for k, v in st.items()}
        return params, bufs, state

    def _restore(self, snap) -> None:
        params, bufs, state = snap
        with torch.no_grad():
            for p, v in params:
                p.copy_(v)
            for b, v in bufs:
                b.copy_(v)
            for p, st in self.opt.state.items():
                old = state.get(p)
                for k, v in st.items():
                    if torch.is_tensor(v):            # the graph holds these addresses: restore in place
                        if old is not None and torch.is_tensor(old.get(k)):
                            v.copy_(old[k])
                        else:
                            v.zero_()                 # state created by the warm-up: back to "never stepped"
                    elif old is not None and k in old:
                        st[k] = old[k]
            for p in self.net.parameters():
                if p.grad is not None:
                    p.grad.zero_()

    def _eager_step(self, step_opt: bool) -> Tensor:
        if step_opt:
            self.opt.zero_grad(set_to_none=True)
        self._out = self.net(self.data, self.dm)
        loss = self.loss_fn(self._out)
        loss.backward()
        if step_opt:
            self.opt.step()
        return loss.detach()

    def __call__(self, dm: Optional[Tensor] = None, z1: Optional[Tensor] = None, x_pos: Optional[Tensor] = None) -> Tensor:
        """Copy the new mask / inputs (device or pinned host tensors) into the static buffers and replay.  Returns the
        static loss tensor (overwritten by the next call)."""
        if dm is not None:
            self.dm.copy_(dm.reshape(self.dm.shape), non_blocking=True)
        if z1 is not None:
            self.z1.copy_(z1, non_blocking=True)
        if x_pos is not None:
            self.x_pos.copy_(x_pos, non_blocking=True)
        self.graph.replay()
        if self.accumulate > 1:
            self._calls += 1
            if self._calls % self.accumulate == 0:       # eager optimizer step on the accumulated gradients (sgcn.py:146-148)
                self.opt.step()
                self.opt.zero_grad(set_to_none=False)
        return self.loss
