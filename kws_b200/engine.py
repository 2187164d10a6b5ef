"""Tensor-level host API over the C ABI: fills the descriptors from torch tensors,
allocates outputs/workspace through torch's caching allocator and enqueues the
work on torch's current stream.  PyTorch is plumbing here (device memory,
streams); all arithmetic happens in ``libfastgrnn_b200.so``.

Layouts follow the reference's two conventions:
``layout="IH"`` = ``FastGRNNCell`` parameters (rnn.py:246-256), ``layout="HI"``
= ``FastGRNNCUDA`` parameters (rnn.py:782-805).
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, Optional, Tuple

import torch

from . import _lib

_MATS = ("W", "U", "W1", "W2", "U1", "U2")


def _present(t) -> bool:
    return t is not None and t.numel() > 0


def _ptr(t) -> Optional[int]:
    return t.data_ptr() if _present(t) else None


def _check_param(name: str, t: torch.Tensor, device: torch.device, shape) -> torch.Tensor:
    if not t.is_cuda:
        raise RuntimeError("%s must be a CUDA tensor" % name)            # cuda/fastgrnn_cuda.cpp:69
    if t.device != device:
        raise RuntimeError("%s is on %s but input is on %s" % (name, t.device, device))
    if t.dtype != torch.float32:
        raise RuntimeError("%s must be float32 (got %s); the recurrent state is fp32" % (name, t.dtype))
    if tuple(t.shape) != tuple(shape):
        raise RuntimeError("%s has shape %s, expected %s" % (name, tuple(t.shape), tuple(shape)))
    if not t.is_contiguous():
        raise RuntimeError("%s must be contiguous" % name)               # cuda/fastgrnn_cuda.cpp:70
    return t


def _nl(name_or_int) -> int:
    if isinstance(name_or_int, int):
        if name_or_int not in _lib.NL_NAMES:
            raise ValueError("unknown nonlinearity enum %r" % (name_or_int,))
        return name_or_int
    if name_or_int not in _lib.NL:
        # rnn.py:62-66 raises ValueError for unknown names; callables cannot cross the C ABI
        raise ValueError("nonlinearity is either a callable or a value "
                         "['tanh', 'sigmoid', 'relu', 'quantTanh', 'quantSigm', 'quantSigm4']; "
                         "the CUDA engine supports the named ones only, got %r" % (name_or_int,))
    return _lib.NL[name_or_int]


class _Problem:
    """Validated view of one call: fills FgrnnProblem and keeps the tensors alive."""

    def __init__(self, x: torch.Tensor, params: Dict[str, torch.Tensor], h0: Optional[torch.Tensor],
                 layout: str, batch_first: bool, gate_nl, update_nl, force_path: int = -1):
        if not isinstance(x, torch.Tensor) or not x.is_cuda:
            raise RuntimeError("input must be a CUDA tensor")            # cuda/fastgrnn_cuda.cpp:69
        if x.dim() != 3:
            raise RuntimeError("input must be 3-D ([T,B,F] or [B,T,F]), got %s" % (tuple(x.shape),))
        if x.dtype not in (torch.float32, torch.bfloat16):
            raise RuntimeError("input must be float32 or bfloat16, got %s" % x.dtype)
        if x.stride(2) != 1 and x.shape[2] > 1:
            # features are not contiguous: the trainer's `audio.permute(2, 0, 1)` view of a (B,F,T) batch
            # (trainClassifier.py:203-204) and its relatives.  One tiled pass through the ingest kernel puts it into
            # (B,T,F); the result is a batch-first tensor, viewed back to the caller's dimension order without a copy.
            x = ingest_features_last(x, batch_first)
        self.x = x
        self.device = x.device
        self.batch_first = bool(batch_first)
        self.B = x.shape[0] if batch_first else x.shape[1]
        self.T = x.shape[1] if batch_first else x.shape[0]
        self.I = x.shape[2]
        bg = params["bias_gate"]
        self.H = H = int(bg.shape[-1])
        self.layout = {"IH": _lib.LAYOUT_IH, "HI": _lib.LAYOUT_HI}[layout]
        ih = layout == "IH"
        W1, U1 = params.get("W1"), params.get("U1")
        self.rW = int(W1.shape[1] if ih else W1.shape[0]) if _present(W1) else 0
        self.rU = int(U1.shape[1] if ih else U1.shape[0]) if _present(U1) else 0
        I, rW, rU = self.I, self.rW, self.rU
        shapes = {
            "W": (I, H) if ih else (H, I), "U": (H, H),
            "W1": (I, rW) if ih else (rW, I), "W2": (rW, H) if ih else (H, rW),
            "U1": (H, rU) if ih else (rU, H), "U2": (rU, H) if ih else (H, rU),
        }
        need = (["W"] if rW == 0 else ["W1", "W2"]) + (["U"] if rU == 0 else ["U1", "U2"])
        self.t = {}
        for k in need:
            if not _present(params.get(k)):
                raise RuntimeError("%s must be a CUDA tensor" % k.lower())
            self.t[k] = _check_param(k.lower(), params[k], self.device, shapes[k])
        self.t["bias_gate"] = _check_param("bias_gate", bg, self.device, (1, H))
        self.t["bias_update"] = _check_param("bias_update", params["bias_update"], self.device, (1, H))
        self.t["zeta"] = _check_param("zeta", params["zeta"], self.device, (1, 1))
        self.t["nu"] = _check_param("nu", params["nu"], self.device, (1, 1))
        # optional per-unit factors on the pre-activations (folded eval-mode BatchNorm, rnn.py:402-408)
        for k in ("gate_scale", "update_scale"):
            if _present(params.get(k)):
                self.t[k] = _check_param(k, params[k], self.device, (1, H))
        if h0 is not None:
            h0 = _check_param("old_h", h0, self.device, (self.B, H))
        self.h0 = h0
        self.gate_nl = _nl(gate_nl)
        self.update_nl = _nl(update_nl)
        self.force_path = int(force_path)

    def strides(self, t: torch.Tensor) -> Tuple[int, int]:
        """(stride_b, stride_t) in elements for a [T,B,*] or [B,T,*] tensor."""
        return (t.stride(0), t.stride(1)) if self.batch_first else (t.stride(1), t.stride(0))

    def fill(self, p: _lib.FgrnnProblem) -> None:
        p.abi_version = _lib.ABI_VERSION
        p.device = self.device.index if self.device.index is not None else torch.cuda.current_device()
        p.B, p.T, p.I, p.H, p.rW, p.rU = self.B, self.T, self.I, self.H, self.rW, self.rU
        p.gate_nl, p.update_nl = self.gate_nl, self.update_nl
        p.weight_layout = self.layout
        p.x_dtype = _lib.BF16 if self.x.dtype == torch.bfloat16 else _lib.F32
        p.force_path = self.force_path
        for k in _MATS:
            setattr(p, k, _ptr(self.t.get(k)))
        for k in ("bias_gate", "bias_update", "zeta", "nu"):
            setattr(p, k, self.t[k].data_ptr())
        p.x = self.x.data_ptr() if self.x.numel() else None
        p.x_stride_b, p.x_stride_t = self.strides(self.x)
        p.h0 = _ptr(self.h0)
        p.gate_scale = _ptr(self.t.get("gate_scale"))
        p.update_scale = _ptr(self.t.get("update_scale"))


def ingest_features_last(x: torch.Tensor, batch_first: bool, mean: Optional[torch.Tensor] = None,
                         std: Optional[torch.Tensor] = None) -> torch.Tensor:
    """x [T,B,F] (or [B,T,F]) with any strides -> a tensor of the same shape whose feature stride is 1, produced by
    ``fgrnn_ingest_bft`` (fp32) in one pass at HBM speed; other dtypes fall back to ``contiguous()``.  With ``mean`` /
    ``std`` (F values each, e.g. the loaders' (1,F,1) arrays) the pass also applies ``(x - mean) / std``
    (preprocessing.py:76) with the reference's rounding, so raw features can be fed."""
    norm = mean is not None
    if norm:
        mean = mean.reshape(-1).to(x.device, torch.float32).contiguous()
        std = std.reshape(-1).to(x.device, torch.float32).contiguous()
        if mean.numel() != x.shape[2] or std.numel() != x.shape[2]:
            raise RuntimeError("mean / std must have %d entries" % x.shape[2])
    if x.dtype != torch.float32 or not x.is_cuda or x.numel() == 0:
        if norm:
            x = (x.float() - mean) / std
        return x.contiguous()
    lib = _lib.load()
    B, T = (x.shape[0], x.shape[1]) if batch_first else (x.shape[1], x.shape[0])
    F = x.shape[2]
    sb, st = (x.stride(0), x.stride(1)) if batch_first else (x.stride(1), x.stride(0))
    if B > 65535:
        return ((x - mean) / std).contiguous() if norm else x.contiguous()
    dst = torch.empty((B, T, F), dtype=torch.float32, device=x.device)
    with torch.cuda.device(x.device):
        _lib.check(lib.fgrnn_ingest_bft(x.data_ptr(), sb, x.stride(2), st, dst.data_ptr(),
                                        mean.data_ptr() if norm else None, std.data_ptr() if norm else None, B, F, T,
                                        x.device.index if x.device.index is not None else torch.cuda.current_device(),
                                        _stream(x.device)), "fastgrnn ingest")
    return dst if batch_first else dst.transpose(0, 1)


def fold_input_normalization(params: Dict[str, torch.Tensor], mean: torch.Tensor, std: torch.Tensor, *,
                             layout: str = "IH") -> Dict[str, torch.Tensor]:
    """Fold the loaders' per-feature standardisation ``(x - mean) / std`` (data_pipeline/preprocessing.py:60-76; mean /
    std of shape (1,F,1), ``model_batchnorm/mean.npy``) into the first layer, so that RAW features can be fed:

        ((x - m) / s) . W  =  x . (diag(1/s) W)  -  (m / s) . W

    i.e. the rows of W (W1 for a low-rank layer) are divided by std and both biases absorb ``-(mean/std) . W``.
    Returns a new parameter dict (same keys, same layout); the inputs are not modified."""
    ih = layout == "IH"
    m = mean.reshape(-1).to(params["bias_gate"])
    s = std.reshape(-1).to(params["bias_gate"])
    out = dict(params)
    with torch.no_grad():
        inv = 1.0 / s
        if _present(params.get("W")):
            W = params["W"] if ih else params["W"].t()                    # [I,H]
            shift = -((m * inv).unsqueeze(0) @ W)                         # [1,H]
            Wn = W * inv.unsqueeze(1)
            out["W"] = (Wn if ih else Wn.t()).contiguous()
        else:
            W1 = params["W1"] if ih else params["W1"].t()                 # [I,r]
            W2 = params["W2"] if ih else params["W2"].t()                 # [r,H]
            shift = -(((m * inv).unsqueeze(0) @ W1) @ W2)
            W1n = W1 * inv.unsqueeze(1)
            out["W1"] = (W1n if ih else W1n.t()).contiguous()
        out["bias_gate"] = (params["bias_gate"] + shift).contiguous()
        out["bias_update"] = (params["bias_update"] + shift).contiguous()
    return out


def _workspace(nbytes: int, device: torch.device) -> Optional[torch.Tensor]:
    if nbytes <= 0:
        return None
    return torch.empty(nbytes, dtype=torch.uint8, device=device)


def _stream(device: torch.device) -> int:
    return torch.cuda.current_stream(device).cuda_stream


def forward(x: torch.Tensor, params: Dict[str, torch.Tensor], h0: Optional[torch.Tensor] = None, *,
            layout: str = "HI", batch_first: bool = False, gate_nl="sigmoid", update_nl="tanh",
            save_for_backward: bool = False, want_states: bool = True, want_last: bool = False,
            force_path: int = -1, out: Optional[torch.Tensor] = None):
    """Run the recurrence over all T steps.

    Returns ``(hidden_states, z_s, c_s, h_last)``: hidden_states in the input's
    layout ([T,B,H] or [B,T,H]); z_s / c_s [T,B,H] when ``save_for_backward``
    (what ``forward_unroll`` returns, cu:414) else None; h_last [B,H] when
    ``want_last`` else None.  ``out`` may be a caller-owned destination (any [T,B,H]/[B,T,H]
    view with unit hidden stride), e.g. a slice of a staging buffer.
    """
    lib = _lib.load()
    pr = _Problem(x, params, h0, layout, batch_first, gate_nl, update_nl, force_path)
    dev, B, T, H = pr.device, pr.B, pr.T, pr.H
    with torch.cuda.device(dev):
        d = _lib.FgrnnForward()
        pr.fill(d.p)
        if not want_states:
            out = None
        if want_states:
            shape = (B, T, H) if pr.batch_first else (T, B, H)
            if out is None:
                out = torch.empty(shape, dtype=torch.float32, device=dev)
            elif (tuple(out.shape) != shape or out.dtype != torch.float32 or out.device != dev
                  or (H > 1 and out.stride(2) != 1)):
                raise RuntimeError("out must be a float32 %s tensor on %s with unit hidden stride" % (shape, dev))
            d.out = out.data_ptr() if out.numel() else None
            d.out_stride_b, d.out_stride_t = pr.strides(out)
        h_last = None
        if want_last or not want_states:
            h_last = torch.empty((B, H), dtype=torch.float32, device=dev)
            if T == 0:
                h_last.copy_(pr.h0) if pr.h0 is not None else h_last.zero_()
            d.h_last = h_last.data_ptr() if h_last.numel() else None
        z_s = c_s = None
        if save_for_backward:
            z_s = torch.empty((T, B, H), dtype=torch.float32, device=dev)
            c_s = torch.empty((T, B, H), dtype=torch.float32, device=dev)
            d.save_z = z_s.data_ptr() if z_s.numel() else None
            d.save_c = c_s.data_ptr() if c_s.numel() else None
        if B * T > 0:
            nbytes = lib.fgrnn_forward_workspace_bytes(C.byref(d))
            ws = _workspace(nbytes, dev)
            d.workspace = ws.data_ptr() if ws is not None else None
            d.workspace_bytes = nbytes
            _lib.check(lib.fgrnn_forward(C.byref(d), _stream(dev)), "fastgrnn forward")
    return out, z_s, c_s, h_last


def forward_plan(x, params, h0=None, *, layout="HI", batch_first=False, gate_nl="sigmoid",
                 update_nl="tanh", force_path=-1) -> str:
    """Name of the kernel family a call with these arguments selects."""
    lib = _lib.load()
    pr = _Problem(x, params, h0, layout, batch_first, gate_nl, update_nl, force_path)
    d = _lib.FgrnnForward()
    pr.fill(d.p)
    d.out = 0x10000  # aligned non-NULL marker; nothing is launched
    d.out_stride_b, d.out_stride_t = (pr.T * pr.H, pr.H) if pr.batch_first else (pr.H, pr.B * pr.H)
    return _lib.PATH_NAMES.get(lib.fgrnn_forward_plan(C.byref(d)), "invalid")


def backward(grad_h: torch.Tensor, x: torch.Tensor, hs: torch.Tensor, z_s: torch.Tensor, c_s: torch.Tensor,
             params: Dict[str, torch.Tensor], h0: Optional[torch.Tensor] = None, *,
             layout: str = "HI", batch_first: bool = False, gate_nl="sigmoid", update_nl="tanh",
             need_dx: bool = True, need_dh0: bool = True, need_params: bool = True,
             grad_bucket: Optional[torch.Tensor] = None, force_path: int = -1,
             grad_t0: int = 0) -> Dict[str, torch.Tensor]:
    """Backward-through-time.  Returns a dict with ``x``, ``h0`` and one entry per
    parameter (same layout/shape as the parameter).  When ``grad_bucket`` (a flat
    fp32 tensor of ``grad_bucket_numel`` floats) is given, the parameter gradients
    are written straight into it (views returned) so one all-reduce covers them.
    ``grad_t0``: ``grad_h`` holds the steps t >= grad_t0 only (T - grad_t0 of them) and the
    earlier steps have no upstream gradient; ``grad_t0 = T - 1`` is the keyword spotter's case
    (model.py:227-231 consumes ``out[-1]``), ``grad_h`` then being one [1,B,H] / [B,1,H] slab."""
    lib = _lib.load()
    pr = _Problem(x, params, h0, layout, batch_first, gate_nl, update_nl, force_path)
    dev, B, T, H, I = pr.device, pr.B, pr.T, pr.H, pr.I
    if not grad_h.is_cuda:
        raise RuntimeError("grad_h must be a CUDA tensor")
    if grad_h.dtype != torch.float32:
        grad_h = grad_h.float()
    if (grad_h.stride(2) != 1 and H > 1) or 0 in grad_h.stride()[:2]:
        grad_h = grad_h.contiguous()         # also materialises expanded gradients (stride 0 from sum/mean)
    if hs.stride(2) != 1 and H > 1:
        hs = hs.contiguous()
    exp = (B, T, H) if pr.batch_first else (T, B, H)
    grad_t0 = int(grad_t0)
    if T > 0 and not (0 <= grad_t0 < T):
        raise RuntimeError("grad_t0 = %d must be in [0, T = %d)" % (grad_t0, T))
    gexp = (B, T - grad_t0, H) if pr.batch_first else (T - grad_t0, B, H)
    if tuple(grad_h.shape) != gexp or tuple(hs.shape) != exp:
        raise RuntimeError("grad_h must have shape %s and hidden_states %s" % (gexp, exp))
    for name, t in (("z", z_s), ("h_prime", c_s)):
        if not t.is_cuda:
            raise RuntimeError("%s must be a CUDA tensor" % name)
        if tuple(t.shape) != (T, B, H) or not t.is_contiguous():
            raise RuntimeError("%s must be a contiguous [T,B,H] tensor" % name)
    with torch.cuda.device(dev):
        g = _lib.FgrnnBackward()
        pr.fill(g.p)
        g.grad_h = grad_h.data_ptr() if grad_h.numel() else None
        g.grad_stride_b, g.grad_stride_t = pr.strides(grad_h)
        g.grad_t0 = grad_t0
        if T - grad_t0 == 1:                 # a single step: its stride is never used, keep it well formed for the TMA map
            g.grad_stride_t = max(int(g.grad_stride_t), B * H)
        g.hs = hs.data_ptr() if hs.numel() else None
        g.hs_stride_b, g.hs_stride_t = pr.strides(hs)
        g.z_s = z_s.data_ptr() if z_s.numel() else None
        g.c_s = c_s.data_ptr() if c_s.numel() else None
        res: Dict[str, torch.Tensor] = {}
        if need_dx:
            dx = torch.empty((B, T, I) if pr.batch_first else (T, B, I), dtype=torch.float32, device=dev)
            g.d_x = dx.data_ptr() if dx.numel() else None
            g.dx_stride_b, g.dx_stride_t = pr.strides(dx)
            res["x"] = dx
        if need_dh0:
            res["h0"] = torch.empty((B, H), dtype=torch.float32, device=dev)
            g.d_h0 = res["h0"].data_ptr() if res["h0"].numel() else None
        if need_params:
            names = (["W"] if pr.rW == 0 else ["W1", "W2"]) + (["U"] if pr.rU == 0 else ["U1", "U2"]) \
                + ["bias_gate", "bias_update", "zeta", "nu"]
            if grad_bucket is not None:
                views = bucket_views(grad_bucket, {k: pr.t[k] for k in names})
            for k in names:
                res[k] = views[k] if grad_bucket is not None else torch.empty_like(pr.t[k])
                setattr(g, "d_" + k, res[k].data_ptr())
        nbytes = lib.fgrnn_backward_workspace_bytes(C.byref(g))
        ws = _workspace(nbytes, dev)
        g.workspace = ws.data_ptr() if ws is not None else None
        g.workspace_bytes = nbytes
        _lib.check(lib.fgrnn_backward(C.byref(g), _stream(dev)), "fastgrnn backward")
    return res


_BUCKET_ORDER = ("W", "W1", "W2", "U", "U1", "U2", "bias_gate", "bias_update", "zeta", "nu")


def grad_bucket_numel(params: Dict[str, torch.Tensor]) -> int:
    return sum(int(params[k].numel()) for k in _BUCKET_ORDER if _present(params.get(k)))


def bucket_views(bucket: torch.Tensor, params: Dict[str, torch.Tensor]) -> Dict[str, torch.Tensor]:
    """Views into one flat fp32 bucket in the fixed order of ``fgrnn_grad_bucket_layout``
    ({W|W1,W2, U|U1,U2, bias_gate, bias_update, zeta, nu})."""
    if bucket.dtype != torch.float32 or not bucket.is_contiguous() or bucket.dim() != 1:
        raise RuntimeError("grad bucket must be a flat contiguous float32 tensor")
    need = grad_bucket_numel(params)
    if bucket.numel() < need:
        raise RuntimeError("grad bucket has %d floats, needs %d" % (bucket.numel(), need))
    out, off = {}, 0
    for k in _BUCKET_ORDER:
        t = params.get(k)
        if _present(t):
            out[k] = bucket[off:off + t.numel()].view(t.shape)
            off += t.numel()
    return out
