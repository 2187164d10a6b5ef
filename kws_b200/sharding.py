"""Multi-GPU plumbing for the FastGRNN path: one process per GPU, ``torch.distributed``
(NCCL over NVLink 5 / NVSwitch on the box, gloo in the CPU tests).

The recurrence shards only along the batch: rows are independent in
forward and in backward up to the parameter-gradient sum; time is strictly sequential.

* inference -- contiguous batch slices per rank, weights replicated, **no collective**
  (outputs stay sharded; ``gather_states`` is an optional convenience).
* training  -- data parallel: every rank runs forward+BPTT on its slice and the parameter
  gradients, which live in ONE flat fp32 bucket (20 738 floats for I=32, H=128), are summed
  with a single all-reduce and divided by the world size.  The payload is ~83 KB, i.e.
  latency-bound; it is issued on the compute stream right after the reduce kernel that
  writes the bucket.

The reference has no distributed code at all; this is new work behind the same module API.
"""
from __future__ import annotations

from typing import Iterable, List, Optional, Sequence, Tuple

import torch
import torch.distributed as dist


def shard_bounds(total: int, world_size: int, rank: int) -> Tuple[int, int]:
    """[begin, end) of ``rank``'s contiguous slice of ``total`` rows; the first ``total % world``
    ranks get one extra row, so slices differ by at most one row and cover [0,total) exactly."""
    if world_size < 1 or not (0 <= rank < world_size):
        raise ValueError("bad rank %d / world %d" % (rank, world_size))
    base, rem = divmod(int(total), world_size)
    begin = rank * base + min(rank, rem)
    return begin, begin + base + (1 if rank < rem else 0)


def shard_batch(x: torch.Tensor, batch_dim: int, world_size: Optional[int] = None,
                rank: Optional[int] = None) -> torch.Tensor:
    """This rank's slice (a view) of a replicated/global batch tensor."""
    world_size = dist.get_world_size() if world_size is None else world_size
    rank = dist.get_rank() if rank is None else rank
    b, e = shard_bounds(x.shape[batch_dim], world_size, rank)
    return x.narrow(batch_dim, b, e - b)


def gather_states(local: torch.Tensor, batch_dim: int, total: int, group=None) -> torch.Tensor:
    """All-gather sharded hidden states back to the global batch (optional; inference itself
    needs no communication).  Handles the ragged last slices by padding to the largest one."""
    world = dist.get_world_size(group)
    sizes = [shard_bounds(total, world, r) for r in range(world)]
    maxn = max(e - b for b, e in sizes)
    pad_shape = list(local.shape)
    pad_shape[batch_dim] = maxn
    buf = local.new_zeros(pad_shape)
    buf.narrow(batch_dim, 0, local.shape[batch_dim]).copy_(local)
    parts = [torch.empty_like(buf) for _ in range(world)]
    dist.all_gather(parts, buf.contiguous(), group=group)
    return torch.cat([p.narrow(batch_dim, 0, e - b) for p, (b, e) in zip(parts, sizes)], dim=batch_dim)


class GradBucket:
    """One flat fp32 bucket holding the gradients of ``params``; ``param.grad`` are views into
    it, so autograd accumulates in place and a single all-reduce covers every parameter."""

    def __init__(self, params: Iterable[torch.nn.Parameter]):
        self.params: List[torch.nn.Parameter] = [p for p in params if p.requires_grad]
        if not self.params:
            raise ValueError("no trainable parameters")
        dev = self.params[0].device
        for p in self.params:
            if p.device != dev or p.dtype != torch.float32:
                raise ValueError("all bucketed parameters must be fp32 on one device")
        self.numel = sum(p.numel() for p in self.params)
        self.flat = torch.zeros(self.numel, dtype=torch.float32, device=dev)
        self.views: List[torch.Tensor] = []
        off = 0
        for p in self.params:
            v = self.flat[off:off + p.numel()].view(p.shape)
            off += p.numel()
            self.views.append(v)
        self.attach()

    def attach(self) -> None:
        for p, v in zip(self.params, self.views):
            p.grad = v

    def zero(self) -> None:
        self.flat.zero_()
        self.attach()

    def all_reduce_mean(self, group=None, async_op: bool = False):
        """sum over ranks, then divide by the world size (mean of per-rank mean-losses ==
        the loss over the concatenated batch when slices have equal size)."""
        world = dist.get_world_size(group)
        if world == 1:
            return None
        work = dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, group=group, async_op=async_op)
        if async_op:
            return _MeanWork(work, self.flat, world)
        self.flat.div_(world)
        return None

    def all_reduce_weighted(self, local_rows: int, total_rows: int, group=None) -> None:
        """For ragged slices: each rank holds the gradient of its *mean* loss over ``local_rows``;
        weight by local_rows/total_rows so the sum equals the gradient of the global mean."""
        self.flat.mul_(float(local_rows) / float(total_rows))
        if dist.get_world_size(group) > 1:
            dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, group=group)


class _MeanWork:
    """Handle returned by ``GradBucket.all_reduce_mean(async_op=True)``: ``wait()`` completes the
    sum AND applies the division by the world size, so a caller that waits and steps the optimizer
    sees the same mean gradient as the blocking call."""

    def __init__(self, work, flat: torch.Tensor, world: int):
        self._work, self._flat, self._world, self._done = work, flat, world, False

    def wait(self) -> bool:
        if not self._done:
            self._work.wait()
            self._flat.div_(self._world)
            self._done = True
        return True

    def is_completed(self) -> bool:
        return self._done or self._work.is_completed()


def broadcast_parameters(params: Sequence[torch.Tensor], src: int = 0, group=None) -> None:
    """Make every rank start from rank ``src``'s weights (one flat broadcast)."""
    params = list(params)
    if not params or dist.get_world_size(group) == 1:
        return
    flat = torch.cat([p.detach().reshape(-1) for p in params])
    dist.broadcast(flat, src=src, group=group)
    off = 0
    with torch.no_grad():
        for p in params:
            p.copy_(flat[off:off + p.numel()].view(p.shape))
            off += p.numel()


def max_over_ranks(value: float, device: torch.device, group=None) -> float:
    """Timing helper: device-measured milliseconds, max over ranks."""
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
    return float(t.item())
