"""Per-kernel-family CUDA-event timing for bench.py's roofline block.

When a ``KernelProfile`` is active every C-ABI wrapper in ops.py brackets its launch with two
CUDA events recorded on the launching stream (torch's current stream, which is the stream we
pass through the C-ABI) and logs the ALGORITHMIC bytes / flops of that launch
(SURVEY.md §8(d)).  Inactive by default: zero overhead on the product path.
"""
from __future__ import annotations

from collections import defaultdict
from typing import Dict, List, Optional

import torch

ACTIVE: Optional["KernelProfile"] = None


class KernelProfile:
    def __init__(self):
        self.records: List[tuple] = []

    def __enter__(self):
        global ACTIVE
        ACTIVE = self
        return self

    def __exit__(self, *exc):
        global ACTIVE
        ACTIVE = None

    def summary(self) -> Dict[str, dict]:
        torch.cuda.synchronize()
        agg = defaultdict(lambda: dict(launches=0, ms=0.0, bytes=0.0, flops=0.0))
        for family, nbytes, flops, e0, e1 in self.records:
            a = agg[family]
            a["launches"] += 1
            a["ms"] += e0.elapsed_time(e1)
            a["bytes"] += nbytes
            a["flops"] += flops
        return dict(agg)


class _Span:
    __slots__ = ("family", "nbytes", "flops", "e0")

    def __init__(self, family, nbytes, flops):
        self.family, self.nbytes, self.flops = family, nbytes, flops
        self.e0 = torch.cuda.Event(enable_timing=True)
        self.e0.record()

    def close(self):
        e1 = torch.cuda.Event(enable_timing=True)
        e1.record()
        if ACTIVE is not None:
            ACTIVE.records.append((self.family, self.nbytes, self.flops, self.e0, e1))


def span(family: str, nbytes: float, flops: float = 0.0) -> Optional[_Span]:
    return _Span(family, nbytes, flops) if ACTIVE is not None else None


class region:
    """``with profile.region("adam"):`` -- brackets any stretch of stream work (torch ops included) as one record of the
    active profile, so that the per-family times of a step add up to the step.  No-op when no profile is active."""

    def __init__(self, family: str, nbytes: float = 0.0, flops: float = 0.0):
        self.sp = span(family, nbytes, flops)

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        if self.sp is not None:
            self.sp.close()
