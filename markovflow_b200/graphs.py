"""CUDA-graph replay of a fixed computation.

A small problem (the reference's own configuration: ONE series of 1,000 states) is bound by launch
overhead, not by the GPU: log-likelihood + posterior + marginals are ~25 kernel launches and as many
ctypes / allocator round trips.  :class:`Graphed` records such a computation once into a CUDA graph
and replays it with one launch; inputs and outputs are the SAME tensors every time (update the inputs
in place, read the outputs after :meth:`__call__`).

The per-chain pivot checks need a device->host read, which a graph cannot contain: they are switched
off while recording and replaying (a failed factorisation shows up as NaNs in the outputs).
"""
from __future__ import annotations

from typing import Any, Callable

import torch

from . import config


class Graphed:
    """``g = Graphed(fn)``: run ``fn()`` a few times eagerly (lazy kernel configuration, allocator
    warm-up), capture it, then ``g()`` replays the captured launches and returns ``fn``'s outputs."""

    def __init__(self, fn: Callable[[], Any], warmup: int = 3) -> None:
        self._prev = config.check_numerics()
        config.set_check_numerics(False)
        try:
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                for _ in range(max(1, warmup)):
                    fn()
            torch.cuda.current_stream().wait_stream(side)
            torch.cuda.synchronize()
            self.graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self.graph):
                self.outputs = fn()
        finally:
            config.set_check_numerics(self._prev)

    def __call__(self):
        self.graph.replay()
        return self.outputs
