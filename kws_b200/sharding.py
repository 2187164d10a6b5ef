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

* training, one node -- ``PeerReducer``: the all-reduce FUSED with the SGD step in one kernel of ours that loads the
  peers' buckets over NVLink peer memory (``csrc/fgrnn_peer.cu``); NCCL stays the fallback (several nodes, no P2P).

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


class _RawDeviceArray:
    """A float32 view of raw device memory for ``torch.as_tensor`` (``__cuda_array_interface__``)."""

    def __init__(self, ptr: int, numel: int):
        self.__cuda_array_interface__ = {"shape": (int(numel),), "typestr": "<f4", "data": (int(ptr), False), "version": 3,
                                         "strides": None}


class PeerReducer:
    """The gradient bucket of one rank in NVLink peer memory, and the fused all-reduce + SGD step over it
    (``fgrnn_sgd_allreduce_peer``, include/fastgrnn_b200.h).

    ``bucket`` is where the backward pass writes this rank's gradients (a float32 tensor view of a ``fgrnn_peer_alloc``
    region whose CUDA IPC handle every other rank has opened).  ``step(params, lr, reduced=...)`` pushes it into every
    peer's receive area, sums all ranks' gradients in rank order and applies ``params -= lr / world * sum`` in one launch
    -- every rank computes the very same bits, there is no NCCL call on the path, and the launch can be captured in a
    CUDA graph.

    Construction is a collective over ``group`` (handles travel through ``all_gather_object``).  All ranks must be
    processes of ONE node whose GPUs have P2P access; otherwise construction raises and the caller keeps NCCL."""

    def __init__(self, numel: int, device: torch.device, group=None):
        import ctypes as C
        from . import _lib
        if device.type != "cuda":
            raise RuntimeError("PeerReducer needs a CUDA device")
        lib = _lib.load()
        self.group, self.device, self.numel = group, device, int(numel)
        self.world, self.rank = dist.get_world_size(group), dist.get_rank(group)
        if self.world > _lib.PEER_MAX_RANKS:
            raise RuntimeError("PeerReducer: %d ranks, at most %d (one NVSwitch node)" % (self.world, _lib.PEER_MAX_RANKS))
        self._index = device.index if device.index is not None else torch.cuda.current_device()
        self._grad_bytes = (4 * self.numel + 255) // 256 * 256
        total = self._grad_bytes + int(lib.fgrnn_peer_recv_bytes(self.numel, self.world))
        # Construction is a collective: a rank that fails locally still takes part in both object all-gathers (carrying its
        # error), so that every rank raises together and none is left waiting for the others
        ptr = C.c_void_p()
        handle = C.create_string_buffer(_lib.PEER_HANDLE_BYTES)
        self._local, self._opened, self.bucket = 0, [], None
        err = None
        try:
            _lib.check(lib.fgrnn_peer_alloc(total, self._index, C.byref(ptr), handle), "peer_alloc")
            self._local = int(ptr.value)
        except RuntimeError as e:
            err = str(e)
        infos = [None] * self.world
        dist.all_gather_object(infos, (bytes(handle.raw), self._index, _node_id(), err), group=group)
        self.ptrs = [0] * self.world
        if err is None:
            try:
                failed = [(r, info[3]) for r, info in enumerate(infos) if info[3] is not None]
                if failed:
                    raise RuntimeError("rank %d could not allocate its region: %s" % failed[0])
                if len({info[2] for info in infos}) != 1:
                    raise RuntimeError("ranks are on different hosts")
                for r, info in enumerate(infos):
                    if r == self.rank:
                        self.ptrs[r] = self._local
                        continue
                    out = C.c_void_p()
                    _lib.check(lib.fgrnn_peer_open(info[0], self._index, C.byref(out)), "peer_open(rank %d)" % r)
                    self.ptrs[r] = int(out.value)
                    self._opened.append(int(out.value))
            except RuntimeError as e:
                err = str(e)
        oks = [None] * self.world
        dist.all_gather_object(oks, err, group=group)
        bad = [(r, e) for r, e in enumerate(oks) if e is not None]
        if bad:
            self.close()
            raise RuntimeError("PeerReducer: peer memory is not available (rank %d: %s)" % bad[0])
        self.bucket = torch.as_tensor(_RawDeviceArray(self._local, self.numel), device=device)
        self.state = torch.zeros(int(lib.fgrnn_peer_state_bytes()) // 4, dtype=torch.int32, device=device)
        self.steps = 0
        torch.cuda.synchronize(device)

    def step(self, params: torch.Tensor, lr: float, reduced: Optional[torch.Tensor] = None, stream: Optional[int] = None) -> None:
        """``params -= lr / world * (sum over ranks of bucket)``; ``reduced`` (optional, [numel]) receives the sum."""
        import ctypes as C
        from . import _lib, engine
        if params.dtype != torch.float32 or not params.is_contiguous() or params.numel() != self.numel:
            raise RuntimeError("PeerReducer.step: params must be a contiguous float32 buffer of %d elements" % self.numel)
        if reduced is not None and (reduced.dtype != torch.float32 or not reduced.is_contiguous() or reduced.numel() < self.numel):
            raise RuntimeError("PeerReducer.step: reduced must be a contiguous float32 buffer of %d elements" % self.numel)
        d = _lib.FgrnnPeerStep()
        d.abi_version, d.device, d.world, d.rank = _lib.ABI_VERSION, self._index, self.world, self.rank
        d.params = params.data_ptr()
        d.reduced = reduced.data_ptr() if reduced is not None else None
        d.bucket = self._local
        for r in range(self.world):
            d.recv[r] = self.ptrs[r] + self._grad_bytes
        d.state = self.state.data_ptr()
        d.n, d.lr, d.grad_scale = self.numel, float(lr), 1.0 / self.world
        with torch.cuda.device(self.device):
            _lib.check(_lib.load().fgrnn_sgd_allreduce_peer(C.byref(d), engine._stream(self.device) if stream is None else stream),
                       "sgd_allreduce_peer")
        self.steps += 1

    def check(self) -> None:
        """After a synchronisation: raise if a peer failed to arrive in some step (the kernel gives up after 4 s)."""
        flag = int(self.state[-1].item())
        if flag:
            raise RuntimeError("PeerReducer: the gradients of rank %d did not arrive in a fused all-reduce step" % (flag - 1))

    def close(self) -> None:
        from . import _lib
        lib = _lib.load()
        for p in self._opened:
            lib.fgrnn_peer_close(p, self._index)
        self._opened = []
        if self._local:
            self.bucket = None
            lib.fgrnn_peer_free(self._local, self._index)
            self._local = 0


def _node_id() -> str:
    import socket
    try:
        with open("/proc/sys/kernel/random/boot_id") as f:
            return f.read().strip()
    except OSError:
        return socket.gethostname()


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
