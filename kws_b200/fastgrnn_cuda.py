"""Drop-in for the reference's ``fastgrnn_cuda`` extension module.

Same four function names and positional argument orders as the pybind module in
``/root/reference/cuda/fastgrnn_cuda.cpp:235-240`` -- ``forward`` (:73-86),
``backward`` (:109-123), ``forward_unroll`` (:147-160), ``backward_unroll``
(:182-197) -- same return lists, same ``RuntimeError`` wording for the
``CHECK_CUDA`` / ``CHECK_CONTIGUOUS`` conditions (:69-71).  All tensors are in
the ``FastGRNNCUDA`` layout (rnn.py:782-805): ``w [H,I]``, ``u [H,H]``,
``w1 [rW,I]``, ``w2 [H,rW]``, ``u1 [rU,H]``, ``u2 [H,rU]``; unused slots are
``torch.empty(0)`` and low rank is selected by ``w1.size(0) != 0`` /
``u1.size(0) != 0`` (cuda/fastgrnn_cuda.cpp:88-99).

Differences by design: work runs on torch's *current* stream (the reference
launches on the legacy default stream); the tanh-gate gradient is the
correct ``1 - z^2`` (the reference's unrolled kernel uses the sigmoid
derivative, cu:519-521); the whole T-step loop is one persistent
kernel instead of ~6 launches per step (cu:367-413).
"""
from __future__ import annotations

from typing import List

import torch

from . import engine


def _check_input(name: str, t: torch.Tensor) -> None:
    # CHECK_INPUT = CHECK_CUDA + CHECK_CONTIGUOUS (cuda/fastgrnn_cuda.cpp:69-71)
    if not t.is_cuda:
        raise RuntimeError("%s must be a CUDA tensor" % name)
    if not t.is_contiguous():
        raise RuntimeError("%s must be contiguous" % name)


def _check_weights(w, u, w1, w2, u1, u2) -> None:
    if w1.size(0) == 0:                                                  # cpp:88-93
        _check_input("w", w)
    else:
        _check_input("w1", w1)
        _check_input("w2", w2)
    if u1.size(0) == 0:                                                  # cpp:94-99
        _check_input("u", u)
    else:
        _check_input("u1", u1)
        _check_input("u2", u2)


def _params(w, u, bias_gate, bias_update, zeta, nu, w1, w2, u1, u2):
    return {"W": w, "U": u, "W1": w1, "W2": w2, "U1": u1, "U2": u2,
            "bias_gate": bias_gate, "bias_update": bias_update, "zeta": zeta, "nu": nu}


def _empty() -> torch.Tensor:
    return torch.empty(0)


def _grad_list(g, low_w: bool, low_u: bool) -> List[torch.Tensor]:
    # order of cuda/fastgrnn_cuda_kernel.cu:317 and :556
    return [g["x"], g["bias_gate"], g["bias_update"], g["zeta"], g["nu"], g["h0"],
            _empty() if low_w else g["W"], _empty() if low_u else g["U"],
            g["W1"] if low_w else _empty(), g["W2"] if low_w else _empty(),
            g["U1"] if low_u else _empty(), g["U2"] if low_u else _empty()]


def forward(input, w, u, bias_gate, bias_update, zeta, nu, old_h, z_non_linearity,
            w1, w2, u1, u2) -> List[torch.Tensor]:
    """One step (cuda/fastgrnn_cuda.cpp:73-107): returns ``[new_h, z, h_prime]``, each [B,H]."""
    _check_input("input", input)
    _check_weights(w, u, w1, w2, u1, u2)
    for n, t in (("bias_gate", bias_gate), ("bias_update", bias_update), ("zeta", zeta),
                 ("nu", nu), ("old_h", old_h)):
        _check_input(n, t)
    out, z, c, _ = engine.forward(input.unsqueeze(0), _params(w, u, bias_gate, bias_update, zeta, nu, w1, w2, u1, u2),
                                  old_h, layout="HI", batch_first=False, gate_nl=int(z_non_linearity),
                                  update_nl="tanh", save_for_backward=True)
    return [out[0], z[0], c[0]]


def backward(grad_h, input, old_h, zeta, nu, w, u, z, h_prime, w1, w2, u1, u2,
             z_non_linearity) -> List[torch.Tensor]:
    """One-step gradients (cuda/fastgrnn_cuda.cpp:109-145): 12 tensors in the order
    ``d_input, d_bias_z, d_bias_h_prime, d_zeta, d_nu, d_old_h, d_w, d_u, d_w1, d_w2, d_u1, d_u2``
    (cu:317).  The bias values are not needed by the backward formulas (cu:81-87) and, as in the
    reference, are not arguments."""
    for n, t in (("grad_h", grad_h), ("input", input), ("old_h", old_h), ("zeta", zeta), ("nu", nu),
                 ("z", z), ("h_prime", h_prime)):
        _check_input(n, t)
    _check_weights(w, u, w1, w2, u1, u2)
    H = old_h.shape[1]
    dummy_bias = torch.empty((1, H), dtype=torch.float32, device=old_h.device)
    g = engine.backward(grad_h.unsqueeze(0), input.unsqueeze(0), z.unsqueeze(0), z.unsqueeze(0), h_prime.unsqueeze(0),
                        _params(w, u, dummy_bias, dummy_bias, zeta, nu, w1, w2, u1, u2), old_h,
                        layout="HI", batch_first=False, gate_nl=int(z_non_linearity), update_nl="tanh")
    g["x"] = g["x"][0]
    return _grad_list(g, w1.size(0) != 0, u1.size(0) != 0)


def forward_unroll(input, w, u, bias_gate, bias_update, zeta, nu, initial_h, z_non_linearity,
                   w1, w2, u1, u2) -> List[torch.Tensor]:
    """T steps (cuda/fastgrnn_cuda.cpp:147-180): ``input [T,B,I]`` ->
    ``[hidden_states, z_s, h_prime_s]``, each [T,B,H] (cu:414)."""
    _check_input("input", input)
    _check_weights(w, u, w1, w2, u1, u2)
    for n, t in (("bias_gate", bias_gate), ("bias_update", bias_update), ("initial_h", initial_h),
                 ("zeta", zeta), ("nu", nu)):
        _check_input(n, t)
    out, z_s, c_s, _ = engine.forward(input, _params(w, u, bias_gate, bias_update, zeta, nu, w1, w2, u1, u2),
                                      initial_h, layout="HI", batch_first=False,
                                      gate_nl=int(z_non_linearity), update_nl="tanh", save_for_backward=True)
    return [out, z_s, c_s]


def backward_unroll(grad_h, input, hidden_states, zeta, nu, w, u, z, h_prime, initial_h,
                    w1, w2, u1, u2, z_non_linearity) -> List[torch.Tensor]:
    """BPTT (cuda/fastgrnn_cuda.cpp:182-232): 12 tensors in the order of cu:556; ``d_old_h`` is the
    gradient of ``initial_h``."""
    for n, t in (("grad_h", grad_h), ("input", input), ("hidden_states", hidden_states), ("z", z),
                 ("h_prime", h_prime)):
        _check_input(n, t)
    _check_weights(w, u, w1, w2, u1, u2)
    for n, t in (("zeta", zeta), ("nu", nu), ("initial_h", initial_h)):
        _check_input(n, t)
    H = initial_h.shape[1]
    dummy_bias = torch.empty((1, H), dtype=torch.float32, device=initial_h.device)
    g = engine.backward(grad_h, input, hidden_states, z, h_prime,
                        _params(w, u, dummy_bias, dummy_bias, zeta, nu, w1, w2, u1, u2), initial_h,
                        layout="HI", batch_first=False, gate_nl=int(z_non_linearity), update_nl="tanh")
    return _grad_list(g, w1.size(0) != 0, u1.size(0) != 0)
