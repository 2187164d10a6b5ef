"""CUDA-graph capture of a whole training step (forward recurrence, loss head, BPTT, gradient all-reduce,
optimizer) so that the ~25 launches of a step replay with one ``cudaGraphLaunch``.

The C-ABI entry points are capture-safe: they allocate nothing, never synchronise the host, enqueue on the
stream they are given, and the TMA tensor maps they encode on the host are baked into the captured kernel
parameters (the buffers they point to come from the graph's private memory pool and stay put across replays).
"""
from __future__ import annotations

from typing import Callable, Optional

import torch

from . import _lib


class CapturedStep:
    """``step = CapturedStep(fn)`` runs ``fn`` a few times eagerly on a side stream, captures one call, and
    replays it on every ``step()``.  ``fn`` must read its inputs from fixed tensors (copy new batches into
    them) and must not synchronise with the host.  ``launches`` is the number of library kernels in one step."""

    def __init__(self, fn: Callable[[], None], warmup: int = 3, pool: Optional[tuple] = None):
        if not torch.cuda.is_available():
            raise RuntimeError("CapturedStep needs a CUDA device")
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(max(warmup, 1)):
                fn()
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        self.graph = torch.cuda.CUDAGraph()
        n0 = _lib.launch_count()
        with torch.cuda.graph(self.graph, pool=pool):
            fn()
        self.launches = _lib.launch_count() - n0

    def __call__(self) -> None:
        self.graph.replay()
