"""The keyword spotter's training step on the last hidden state, in six launches.

trainClassifier.py:225-240 does, per batch: ``optimizer.zero_grad(); out = model(x); loss = NLLLoss(out, y);
loss.backward(); optimizer.step()`` where model.py:200-231 runs the FastGRNN layer and applies
``hidden2keyword`` + ``log_softmax`` to ``out[-1]`` only.  Through autograd that is the recurrence, ~15 small head
launches, a dense all-zero ``[T,B,H]`` gradient for the ``out[-1]`` slice (written, then read back by BPTT), the reverse
recurrence and a handful of optimizer launches.  Here the same arithmetic is

    1. forward recurrence (saves z_t / c_t)                        ``fgrnn_forward``
    2. head forward + loss + head backward, one kernel             ``fgrnn_head_nll``  -> loss, dW_head, db_head, dh_T
    3. BPTT from the LAST state's gradient only (``grad_t0=T-1``)  ``fgrnn_backward``  -> reverse recurrence,
       contraction, reduce; parameter gradients land in ONE flat bucket
    4. + 5. (data parallel, one node) all-reduce FUSED with SGD    ``fgrnn_sgd_allreduce_peer`` -- one kernel that loads the
       peers' buckets over NVLink peer memory, sums them in rank order and updates the parameters (csrc/fgrnn_peer.cu)
    4. (data parallel, NCCL fallback) one all-reduce of the bucket
    5. SGD over the flat parameter buffer                          ``fgrnn_sgd_flat``

with no autograd graph.  Parameters of the layer and the head are re-pointed at views of one flat buffer, so the
modules stay usable (``state_dict``, inference, the autograd path) and see every update.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional

import torch
import torch.distributed as dist

from . import _lib, engine


def head_nll(h_last: torch.Tensor, weight: torch.Tensor, bias: torch.Tensor, labels: torch.Tensor, *,
             dW: Optional[torch.Tensor] = None, db: Optional[torch.Tensor] = None, want_dh: bool = True,
             want_logp: bool = False, workspace: Optional[torch.Tensor] = None):
    """``hidden2keyword`` + ``log_softmax`` + mean NLL loss and their gradients in one launch (model.py:227-231,
    trainClassifier.py:236).  ``h_last`` [B,H] (rows may be strided, e.g. ``out[-1]`` or ``out[:, -1]``), ``weight``
    [C,H], ``bias`` [C], ``labels`` int64 [B].  Returns ``(loss, dW, db, dh, logp)``; dh / logp are None unless asked."""
    lib = _lib.load()
    if not h_last.is_cuda:
        raise RuntimeError("head_nll: h_last must be a CUDA tensor")
    B, H = h_last.shape
    Cn = weight.shape[0]
    dev = h_last.device
    if h_last.dtype != torch.float32 or (H > 1 and h_last.stride(1) != 1):
        h_last = h_last.float().contiguous()
    if labels.dtype != torch.int64 or not labels.is_contiguous():
        labels = labels.long().contiguous()
    if tuple(weight.shape) != (Cn, H) or tuple(bias.shape) != (Cn,) or tuple(labels.shape) != (B,):
        raise RuntimeError("head_nll: weight %s / bias %s / labels %s do not match h_last %s"
                           % (tuple(weight.shape), tuple(bias.shape), tuple(labels.shape), (B, H)))
    weight, bias = weight.contiguous(), bias.contiguous()
    loss = torch.empty((), dtype=torch.float32, device=dev)
    dW = torch.empty((Cn, H), dtype=torch.float32, device=dev) if dW is None else dW
    db = torch.empty((Cn,), dtype=torch.float32, device=dev) if db is None else db
    dh = torch.empty((B, H), dtype=torch.float32, device=dev) if want_dh else None
    logp = torch.empty((B, Cn), dtype=torch.float32, device=dev) if want_logp else None
    need = lib.fgrnn_head_workspace_bytes(B, H, Cn)
    if workspace is None:
        workspace = torch.zeros(need, dtype=torch.uint8, device=dev)
    elif workspace.numel() < need:
        raise RuntimeError("head_nll: workspace has %d bytes, needs %d" % (workspace.numel(), need))
    with torch.cuda.device(dev):
        _lib.check(lib.fgrnn_head_nll(h_last.data_ptr(), h_last.stride(0), weight.data_ptr(), bias.data_ptr(), labels.data_ptr(),
                                      loss.data_ptr(), dW.data_ptr(), db.data_ptr(), dh.data_ptr() if want_dh else None,
                                      dh.stride(0) if want_dh else 0, logp.data_ptr() if want_logp else None,
                                      workspace.data_ptr(), workspace.numel(), B, H, Cn,
                                      dev.index if dev.index is not None else torch.cuda.current_device(),
                                      engine._stream(dev)), "fastgrnn head")
    return loss, dW, db, dh, logp


def sgd_flat(params: torch.Tensor, grads: torch.Tensor, lr: float, grad_scale: float = 1.0) -> None:
    """``params -= lr * grad_scale * grads`` over two flat fp32 buffers, one launch."""
    lib = _lib.load()
    if params.dtype != torch.float32 or grads.dtype != torch.float32 or not params.is_contiguous() or not grads.is_contiguous():
        raise RuntimeError("sgd_flat: flat contiguous float32 buffers required")
    n = params.numel()
    if grads.numel() < n:
        raise RuntimeError("sgd_flat: %d gradients for %d parameters" % (grads.numel(), n))
    dev = params.device
    with torch.cuda.device(dev):
        _lib.check(lib.fgrnn_sgd_flat(params.data_ptr(), grads.data_ptr(), n, C.c_float(lr), C.c_float(grad_scale),
                                      dev.index if dev.index is not None else torch.cuda.current_device(),
                                      engine._stream(dev)), "fastgrnn sgd")


class LastStateTrainStep:
    """``step(x, labels) -> loss`` for one FastGRNN layer (``kws_b200.rnn.FastGRNN``) followed by the
    ``hidden2keyword`` linear head, plain SGD.  ``group``: process group for data-parallel training (gradients of the
    ranks' mean losses are averaged, i.e. the loss over the concatenated batch when the slices have equal size).
    ``collective``: "peer" = the fused all-reduce + SGD kernel over NVLink peer memory (``sharding.PeerReducer``; raises if
    the ranks cannot map each other's memory), "nccl" = ``dist.all_reduce`` + ``fgrnn_sgd_flat``, "auto" = peer when it can
    be set up, else nccl (``self.collective`` tells which)."""

    def __init__(self, layer, head: torch.nn.Linear, lr: float, group=None, data_parallel: Optional[bool] = None,
                 collective: str = "auto"):
        if type(layer).__name__ != "FastGRNN" or getattr(layer, "_bidirectional", False):
            raise RuntimeError("LastStateTrainStep drives one unidirectional kws_b200.rnn.FastGRNN layer")
        cell = layer.cell
        self.layer, self.cell, self.head, self.lr, self.group = layer, cell, head, float(lr), group
        self.layout = "IH"
        self.batch_first = bool(layer._batch_first)
        self.gate_nl, self.update_nl = cell.gate_nonlinearity, cell.update_nonlinearity
        names = [k for k in engine._BUCKET_ORDER if engine._present(getattr(cell, k, None))]
        plist = [getattr(cell, k) for k in names] + [head.weight, head.bias]
        dev = plist[0].device
        if dev.type != "cuda":
            raise RuntimeError("LastStateTrainStep: the modules must be on a CUDA device (no CPU fallback)")
        self.device = dev
        self.n_cell = sum(p.numel() for p in plist[:-2])
        total = sum(p.numel() for p in plist)
        self.flat_params = torch.empty(total, dtype=torch.float32, device=dev)
        self.flat_grads = torch.zeros(total, dtype=torch.float32, device=dev)
        off = 0
        with torch.no_grad():
            for p in plist:
                view = self.flat_params[off:off + p.numel()].view(p.shape)
                view.copy_(p)
                p.data = view                       # the module's parameter now IS a slice of the flat buffer
                p.grad = self.flat_grads[off:off + p.numel()].view(p.shape)
                off += p.numel()
        self.params = {k: getattr(cell, k).data for k in names}
        H, Cn = head.weight.shape[1], head.weight.shape[0]
        self._head_ws = None
        self._H, self._C = H, Cn
        self.world = 1
        if data_parallel is None:
            data_parallel = dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1
        if data_parallel:
            self.world = dist.get_world_size(group)
        # where this rank's gradients are written: the flat bucket itself, or -- fused peer step -- a bucket in NVLink peer
        # memory (flat_grads, i.e. every p.grad, then receives the all-reduced SUM from the same kernel that updates the parameters)
        self.peer = None
        self.collective = "none" if self.world == 1 else "nccl"
        if collective not in ("auto", "peer", "nccl"):
            raise ValueError("collective must be 'auto', 'peer' or 'nccl'")
        if self.world > 1 and collective != "nccl":
            from . import sharding
            try:
                self.peer = sharding.PeerReducer(total, dev, group)
                self.collective = "peer"
            except RuntimeError:
                if collective == "peer":
                    raise
        self.bucket = self.peer.bucket if self.peer is not None else self.flat_grads
        self.dW_head = self.bucket[self.n_cell:self.n_cell + head.weight.numel()].view(head.weight.shape)
        self.db_head = self.bucket[self.n_cell + head.weight.numel():total].view(head.bias.shape)

    def compute(self, x: torch.Tensor, labels: torch.Tensor) -> torch.Tensor:
        """Forward, loss and every gradient (into ``flat_grads``); no collective, no update."""
        out, z_s, c_s, _ = engine.forward(x, self.params, None, layout=self.layout, batch_first=self.batch_first,
                                          gate_nl=self.gate_nl, update_nl=self.update_nl, save_for_backward=True)
        T = out.shape[1] if self.batch_first else out.shape[0]
        B = out.shape[0] if self.batch_first else out.shape[1]
        h_last = out[:, -1] if self.batch_first else out[-1]
        need = _lib.load().fgrnn_head_workspace_bytes(B, self._H, self._C)
        if self._head_ws is None or self._head_ws.numel() < need:
            self._head_ws = torch.zeros(need, dtype=torch.uint8, device=self.device)
        loss, _, _, dh, _ = head_nll(h_last, self.head.weight.data, self.head.bias.data, labels, dW=self.dW_head,
                                     db=self.db_head, want_dh=True, workspace=self._head_ws)
        gh = dh.unsqueeze(1) if self.batch_first else dh.unsqueeze(0)
        engine.backward(gh, x, out, z_s, c_s, self.params, None, layout=self.layout, batch_first=self.batch_first,
                        gate_nl=self.gate_nl, update_nl=self.update_nl, need_dx=False, need_dh0=False,
                        grad_bucket=self.bucket, grad_t0=T - 1)
        return loss

    def update(self) -> None:
        if self.peer is not None:
            self.peer.step(self.flat_params, self.lr, reduced=self.flat_grads)
            return
        if self.world > 1:
            dist.all_reduce(self.flat_grads, op=dist.ReduceOp.SUM, group=self.group)
        sgd_flat(self.flat_params, self.flat_grads, self.lr, 1.0 / self.world)

    def __call__(self, x: torch.Tensor, labels: torch.Tensor) -> torch.Tensor:
        loss = self.compute(x, labels)
        self.update()
        return loss
