"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the reference FastGRNN path.

Parity status: parity unpinned by the reference's own tests (it has none);
pinned instead against the reference code run in the build container, see
``oracle/__init__.py``.

Every function cites the reference lines (relative to ``/root/reference``) it
follows.  The restatement uses the same ATen ops in the same order as the
reference so that, on one machine, it is bit-identical to the shimmed
reference (asserted in ``tests/test_oracle_vs_reference.py``).

Parameter layout here is the reference's *oracle* (``FastGRNNCell``) layout,
``rnn.py:246-261``:  W [I,H] | W1 [I,rW], W2 [rW,H];  U [H,H] | U1 [H,rU],
U2 [rU,H];  bias_gate, bias_update [1,H];  zeta, nu [1,1].  The CUDA-module
layout (``rnn.py:494-517``) is the transpose of every matrix; see
``to_cuda_layout`` / ``from_cuda_layout``.
"""
from __future__ import annotations

from dataclasses import dataclass, fields
from typing import Dict, Optional

import torch

NONLINEARITIES = ("sigmoid", "tanh", "relu", "quantTanh", "quantSigm", "quantSigm4")


def nonlinearity(A: torch.Tensor, name: str) -> torch.Tensor:
    """rnn.py:40-67 ``gen_nonlinearity``.

    ``relu`` is unrunnable in the reference (``torch.relu(A, 0.0)`` raises,
    rnn.py:52); it is restated as the evident intent
    ``max(A, 0)`` and has no reference-run pin.
    """
    if name == "tanh":
        return torch.tanh(A)                                             # rnn.py:47
    if name == "sigmoid":
        return torch.sigmoid(A)                                          # rnn.py:49
    if name == "relu":
        return torch.relu(A)                                             # rnn.py:51-52 (intent)
    if name == "quantTanh":                                              # rnn.py:53-54
        return torch.max(torch.min(A, torch.ones_like(A)), -1.0 * torch.ones_like(A))
    if name == "quantSigm":                                              # rnn.py:55-57
        A = (A + 1.0) / 2.0
        return torch.max(torch.min(A, torch.ones_like(A)), torch.zeros_like(A))
    if name == "quantSigm4":                                             # rnn.py:58-60
        A = (A + 2.0) / 4.0
        return torch.max(torch.min(A, torch.ones_like(A)), torch.zeros_like(A))
    raise ValueError("unknown nonlinearity %r" % (name,))                # rnn.py:62-66


@dataclass
class Params:
    """Parameter container in the oracle layout (rnn.py:246-261)."""
    bias_gate: torch.Tensor
    bias_update: torch.Tensor
    zeta: torch.Tensor
    nu: torch.Tensor
    W: Optional[torch.Tensor] = None
    U: Optional[torch.Tensor] = None
    W1: Optional[torch.Tensor] = None
    W2: Optional[torch.Tensor] = None
    U1: Optional[torch.Tensor] = None
    U2: Optional[torch.Tensor] = None

    @property
    def hidden_size(self) -> int:
        return int(self.bias_gate.shape[1])

    @property
    def input_size(self) -> int:
        return int((self.W if self.W is not None else self.W1).shape[0])

    def tensors(self) -> Dict[str, torch.Tensor]:
        return {f.name: getattr(self, f.name) for f in fields(self)
                if getattr(self, f.name) is not None}

    def map(self, fn) -> "Params":
        return Params(**{k: fn(v) for k, v in self.tensors().items()})

    def requires_grad_(self, flag: bool = True) -> "Params":
        for v in self.tensors().values():
            v.requires_grad_(flag)
        return self


def init_params(input_size: int, hidden_size: int, wRank=None, uRank=None,
                zetaInit: float = 1.0, nuInit: float = -4.0,
                generator: Optional[torch.Generator] = None) -> Params:
    """rnn.py:246-261: matrices 0.1*randn in the order W|W1,W2 then U|U1,U2;
    biases ones; zeta/nu constants.  Under the same ``torch.manual_seed`` this
    draws the same numbers as ``rnn.FastGRNNCell.__init__``."""
    def rn(*shape):
        return 0.1 * torch.randn(list(shape), generator=generator)
    kw = {}
    if wRank is None:
        kw["W"] = rn(input_size, hidden_size)                            # rnn.py:247
    else:
        kw["W1"] = rn(input_size, wRank)                                 # rnn.py:249
        kw["W2"] = rn(wRank, hidden_size)                                # rnn.py:250
    if uRank is None:
        kw["U"] = rn(hidden_size, hidden_size)                           # rnn.py:253
    else:
        kw["U1"] = rn(hidden_size, uRank)                                # rnn.py:255
        kw["U2"] = rn(uRank, hidden_size)                                # rnn.py:256
    return Params(bias_gate=torch.ones([1, hidden_size]),                # rnn.py:258
                  bias_update=torch.ones([1, hidden_size]),              # rnn.py:259
                  zeta=zetaInit * torch.ones([1, 1]),                    # rnn.py:260
                  nu=nuInit * torch.ones([1, 1]), **kw)                  # rnn.py:261


def cell_step(x: torch.Tensor, h: torch.Tensor, p: Params,
              gate_nl: str = "sigmoid", update_nl: str = "tanh") -> torch.Tensor:
    """One recurrence step, rnn.py:273-297 (``FastGRNNCell.forward``)."""
    if p.W is not None:
        wComp = torch.matmul(x, p.W)                                     # rnn.py:278
    else:
        wComp = torch.matmul(torch.matmul(x, p.W1), p.W2)                # rnn.py:280-281
    if p.U is not None:
        uComp = torch.matmul(h, p.U)                                     # rnn.py:284
    else:
        uComp = torch.matmul(torch.matmul(h, p.U1), p.U2)                # rnn.py:286-287
    pre_comp = wComp + uComp                                             # rnn.py:289
    z = nonlinearity(pre_comp + p.bias_gate, gate_nl)                    # rnn.py:290-291
    c = nonlinearity(pre_comp + p.bias_update, update_nl)                # rnn.py:292-293
    new_h = z * h + (torch.sigmoid(p.zeta) * (1.0 - z)                   # rnn.py:294-295
                     + torch.sigmoid(p.nu)) * c
    return new_h


def bn_eval(a: torch.Tensor, bn: Dict[str, torch.Tensor]) -> torch.Tensor:
    """``nn.BatchNorm1d`` in eval mode on a [B,H] tensor (what ``self.bn_w.eval()(wComp)`` computes, rnn.py:395):
    running statistics, affine weight / bias.  ``bn`` = dict(weight, bias, mean, var, eps)."""
    return torch.nn.functional.batch_norm(a, bn["mean"], bn["var"], bn["weight"], bn["bias"], False, 0.0, float(bn["eps"]))


def cell_step_bn(x: torch.Tensor, h: torch.Tensor, p: Params, bns: Dict[str, Dict[str, torch.Tensor]],
                 gate_nl: str = "sigmoid", update_nl: str = "tanh") -> torch.Tensor:
    """One step of ``FastGRNNBatchNormCell.forward`` with ``training=False`` (rnn.py:377-410).
    ``bns`` has the four layers "w", "u", "gate", "update" (rnn.py:363-366)."""
    if p.W is not None:
        wComp = torch.matmul(x, p.W)                                     # rnn.py:385
    else:
        wComp = torch.matmul(torch.matmul(x, p.W1), p.W2)                # rnn.py:387
    if p.U is not None:
        uComp = torch.matmul(h, p.U)                                     # rnn.py:391
    else:
        uComp = torch.matmul(torch.matmul(h, p.U1), p.U2)                # rnn.py:393
    wComp = bn_eval(wComp, bns["w"])                                     # rnn.py:396
    uComp = bn_eval(uComp, bns["u"])                                     # rnn.py:397
    pre_gate = wComp + uComp + p.bias_gate                               # rnn.py:400
    pre_update = wComp + uComp + p.bias_update                           # rnn.py:401
    pre_gate = bn_eval(pre_gate, bns["gate"])                            # rnn.py:404
    pre_update = bn_eval(pre_update, bns["update"])                      # rnn.py:405
    z = nonlinearity(pre_gate, gate_nl)                                  # rnn.py:408
    c = nonlinearity(pre_update, update_nl)                              # rnn.py:409
    return z * h + (torch.sigmoid(p.zeta) * (1.0 - z) + torch.sigmoid(p.nu)) * c      # rnn.py:412-413


def unroll_bn(x: torch.Tensor, p: Params, bns, h0: Optional[torch.Tensor] = None, batch_first: bool = False,
              gate_nl: str = "sigmoid", update_nl: str = "tanh") -> torch.Tensor:
    """All T hidden states of the eval-mode BatchNorm variant (``FastGRNNBatchNorm`` -> ``BaseRNN.forward``,
    rnn.py:709-734, :620-622 / :658-660), dtype-generic.  h0 is [B,H]."""
    xs = x.transpose(0, 1) if batch_first else x
    T, B = xs.shape[0], xs.shape[1]
    h = torch.zeros(B, p.hidden_size, dtype=x.dtype) if h0 is None else h0
    outs = []
    for t in range(T):
        h = cell_step_bn(xs[t], h, p, bns, gate_nl, update_nl)
        outs.append(h)
    out = torch.stack(outs, 0)
    return out.transpose(0, 1) if batch_first else out


def unroll(x: torch.Tensor, p: Params, h0: Optional[torch.Tensor] = None,
           batch_first: bool = False, gate_nl: str = "sigmoid",
           update_nl: str = "tanh") -> torch.Tensor:
    """All T hidden states in the input's layout, rnn.py:574-668
    (``BaseRNN.forward``, unidirectional FastGRNN branch :620-622 / :658-660).

    ``h0`` follows the reference ``hiddenState`` convention: shape [1,B,H]
    (rnn.py:588-591) and it is **mutated in place** like the reference's; a
    [B,H] tensor is accepted and wrapped.  Output dtype follows the reference
    quirk D12: the state buffers are default-dtype (rnn.py:579-591) unless the
    caller passes fp64 inputs *and* an fp64 ``h0``; for fp64 truth runs use
    ``unroll_functional``.
    """
    H = p.hidden_size
    hiddenStates = torch.zeros([x.shape[0], x.shape[1], H])              # rnn.py:579-581
    if h0 is None:
        hiddenState = torch.zeros([1, x.shape[0] if batch_first else x.shape[1], H])  # :588-591
    else:
        hiddenState = h0 if h0.dim() == 3 else h0.unsqueeze(0)
    if batch_first:
        for i in range(0, x.shape[1]):                                   # rnn.py:620
            hiddenState[0] = cell_step(x[:, i, :], hiddenState[0].clone(), p, gate_nl, update_nl)  # :621
            hiddenStates[:, i, :] = hiddenState[0]                       # rnn.py:622
    else:
        for i in range(0, x.shape[0]):                                   # rnn.py:658
            hiddenState[0] = cell_step(x[i, :, :], hiddenState[0].clone(), p, gate_nl, update_nl)  # :659
            hiddenStates[i, :, :] = hiddenState[0]                       # rnn.py:660
    return hiddenStates                                                  # rnn.py:628 / :666


def unroll_functional(x: torch.Tensor, p: Params, h0: Optional[torch.Tensor] = None,
                      batch_first: bool = False, gate_nl: str = "sigmoid",
                      update_nl: str = "tanh") -> torch.Tensor:
    """Same math as ``unroll`` without the in-place buffers, dtype-generic
    (used for the fp64 "truth" runs).  h0 is [B,H]."""
    xs = x.transpose(0, 1) if batch_first else x
    T, B = xs.shape[0], xs.shape[1]
    h = torch.zeros(B, p.hidden_size, dtype=x.dtype) if h0 is None else h0
    outs = []
    for t in range(T):
        h = cell_step(xs[t], h, p, gate_nl, update_nl)
        outs.append(h)
    out = torch.stack(outs, 0)
    return out.transpose(0, 1) if batch_first else out


def head_logits(out_time_major_like: torch.Tensor, weight: torch.Tensor,
                bias: torch.Tensor, apply_softmax: bool = True) -> torch.Tensor:
    """model.py:227-231: ``hidden2keyword(model_output[-1, :, :])`` followed by
    ``log_softmax(dim=1)``.  Note the reference indexes dim 0 regardless of
    ``batch_first``; this restates that faithfully."""
    y = torch.nn.functional.linear(out_time_major_like[-1, :, :], weight, bias)  # model.py:228
    if apply_softmax:
        y = torch.nn.functional.log_softmax(y, dim=1)                    # model.py:230
    return y


def _dgate(z: torch.Tensor, pre_b: torch.Tensor, name: str) -> torch.Tensor:
    """Derivative of the gate/update nonlinearity expressed on its output
    (cu:27-40 for sigmoid/relu/tanh) or on its input for the clamps."""
    if name == "sigmoid":
        return z * (1.0 - z)                                             # cu:28-30
    if name == "tanh":
        return 1.0 - z * z                                               # cu:38-40
    if name == "relu":
        return (z != 0).to(z.dtype)                                      # cu:33-35
    if name == "quantTanh":
        return ((pre_b > -1.0) & (pre_b < 1.0)).to(z.dtype)
    if name == "quantSigm":
        return ((pre_b > -1.0) & (pre_b < 1.0)).to(z.dtype) * 0.5
    if name == "quantSigm4":
        return ((pre_b > -2.0) & (pre_b < 2.0)).to(z.dtype) * 0.25
    raise ValueError(name)


def bptt_closed_form(x_tm: torch.Tensor, p: Params, h0: torch.Tensor, grad_h_tm: torch.Tensor,
                     gate_nl: str = "sigmoid", update_nl: str = "tanh") -> Dict[str, torch.Tensor]:
    """Closed-form backward-through-time for time-major x [T,B,I], grad [T,B,H].

    Elementwise part: cuda/fastgrnn_cuda_kernel.cu:109-118; matrix part
    :537-545; low-rank chain rule :546-555 -- but with the *correct* tanh-gate
    derivative (the reference instantiates d_sigmoid for the tanh gate in the
    unrolled path, cu:519-521) and in the oracle [I,H] layout.
    Cross-checked against autograd of ``unroll`` in the tests; the autograd
    result is the gradient oracle, this is the kernel-shaped restatement.
    """
    T, B, _ = x_tm.shape
    dt = x_tm.dtype
    s_z, s_n = torch.sigmoid(p.zeta), torch.sigmoid(p.nu)
    hs, zs, cs, pres = [], [], [], []
    h = h0
    for t in range(T):
        w = x_tm[t] @ p.W if p.W is not None else (x_tm[t] @ p.W1) @ p.W2
        u = h @ p.U if p.U is not None else (h @ p.U1) @ p.U2
        pre = w + u
        z = nonlinearity(pre + p.bias_gate, gate_nl)
        c = nonlinearity(pre + p.bias_update, update_nl)
        hs.append(h); zs.append(z); cs.append(c); pres.append(pre)
        h = z * h + (s_z * (1.0 - z) + s_n) * c
    H, I = p.hidden_size, p.input_size
    g = {k: torch.zeros_like(v) for k, v in p.tensors().items()}
    d_x = torch.zeros_like(x_tm)
    delta = torch.zeros(B, H, dtype=dt)
    d_zeta = torch.zeros((), dtype=dt); d_nu = torch.zeros((), dtype=dt)
    Wfull = p.W if p.W is not None else p.W1 @ p.W2
    Ufull = p.U if p.U is not None else p.U1 @ p.U2
    for t in range(T - 1, -1, -1):
        G = grad_h_tm[t] + delta                                         # cu:474
        z, c, hp, pre = zs[t], cs[t], hs[t], pres[t]
        dc = (s_z * (1.0 - z) + s_n) * _dgate(c, pre + p.bias_update, update_nl) * G    # cu:111
        dz = (hp - s_z * c) * _dgate(z, pre + p.bias_gate, gate_nl) * G  # cu:112
        dpre = dc + dz                                                   # cu:115
        g["bias_update"] += dc.sum(0, keepdim=True)                      # cu:113,543
        g["bias_gate"] += dz.sum(0, keepdim=True)                        # cu:114,542
        d_zeta = d_zeta + ((1.0 - z) * c * G).sum()                      # cu:116
        d_nu = d_nu + (c * G).sum()                                      # cu:117
        if p.W is not None:
            g["W"] += x_tm[t].t() @ dpre                                 # cu:539 (transposed layout)
        else:
            g["W2"] += (x_tm[t] @ p.W1).t() @ dpre
            g["W1"] += x_tm[t].t() @ (dpre @ p.W2.t())
        if p.U is not None:
            g["U"] += hp.t() @ dpre                                      # cu:540
        else:
            g["U2"] += (hp @ p.U1).t() @ dpre
            g["U1"] += hp.t() @ (dpre @ p.U2.t())
        d_x[t] = dpre @ Wfull.t()                                        # cu:538
        delta = z * G + dpre @ Ufull.t()                                 # cu:110,537
    g["zeta"] = (d_zeta * s_z * (1.0 - s_z)).reshape(1, 1)               # cu:116,544
    g["nu"] = (d_nu * s_n * (1.0 - s_n)).reshape(1, 1)                   # cu:117,545
    g["x"] = d_x
    g["h0"] = delta
    return g


def autograd_grads(x: torch.Tensor, p: Params, h0: Optional[torch.Tensor], grad_out: torch.Tensor,
                   batch_first: bool = False, gate_nl: str = "sigmoid",
                   update_nl: str = "tanh") -> Dict[str, torch.Tensor]:
    """The gradient oracle: torch.autograd through ``unroll`` on CPU.
    Returns grads for every parameter plus ``x`` and ``h0`` ([B,H])."""
    x = x.detach().clone().requires_grad_(True)
    q = p.map(lambda v: v.detach().clone().requires_grad_(True))
    B = x.shape[0] if batch_first else x.shape[1]
    h0_leaf = (torch.zeros(B, q.hidden_size) if h0 is None
               else h0.detach().reshape(B, q.hidden_size).clone()).requires_grad_(True)
    # the reference mutates hiddenState in place (rnn.py:621); feed a non-leaf view
    out = unroll(x, q, (h0_leaf * 1.0).unsqueeze(0), batch_first, gate_nl, update_nl)
    out.backward(grad_out)
    g = {k: v.grad for k, v in q.tensors().items()}
    g["x"] = x.grad
    g["h0"] = h0_leaf.grad
    return g


def to_cuda_layout(p: Params) -> Dict[str, torch.Tensor]:
    """Oracle layout -> FastGRNNCUDA layout (rnn.py:782-805): every matrix
    transposed: W [H,I], U [H,H]ᵀ, W1 [rW,I], W2 [H,rW], U1 [rU,H], U2 [H,rU]."""
    out = {}
    for k, v in p.tensors().items():
        out[k] = v.t().contiguous() if k in ("W", "U", "W1", "W2", "U1", "U2") else v.clone()
    return out


def from_cuda_layout(d: Dict[str, torch.Tensor]) -> Params:
    kw = {}
    for k, v in d.items():
        if v is None or v.numel() == 0:
            continue
        kw[k] = v.t().contiguous() if k in ("W", "U", "W1", "W2", "U1", "U2") else v.clone()
    return Params(**kw)


def tolerance_ratio(got: torch.Tensor, ref: torch.Tensor, rtol: float, atol: float) -> float:
    """max |got-ref| / (atol + rtol*|ref|): <=1 passes ``allclose``."""
    got = got.double(); ref = ref.double()
    return float(((got - ref).abs() / (atol + rtol * ref.abs())).max())


def grad_tolerance_ratio(got: torch.Tensor, ref: torch.Tensor, rtol: float = 1e-4) -> float:
    """Gradient criterion: rtol 1e-4
    with the absolute floor atol = rtol * max|ref| per tensor, because a pure
    elementwise rtol fails on near-zero entries even for fp32-vs-fp64 autograd
    of the oracle itself."""
    atol = rtol * float(ref.abs().max()) if ref.numel() else 0.0
    if atol == 0.0:
        atol = 1e-12
    return tolerance_ratio(got, ref, rtol, atol)
