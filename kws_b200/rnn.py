"""``rnn``-compatible operator surface on the B200 engine.

Drop-in for the reference's ``rnn.py`` for the FastGRNN path: same class and
function names, constructor keywords, parameter names/shapes (so reference
checkpoints load), call signatures and error behaviour -- but every T-step
unroll is ONE persistent CUDA kernel launch through ``libfastgrnn_b200.so``
instead of a per-timestep Python loop (rnn.py:620-630 / :658-668) or a host C++
loop of ~6 launches per step (cuda/fastgrnn_cuda_kernel.cu:367-413).

Mirrored reference symbols (file:line in /root/reference):
  gen_nonlinearity rnn.py:40-67 | RNNCell rnn.py:69-190 | FastGRNNCell rnn.py:192-313
  FastGRNNCUDACell rnn.py:454-549 | BaseRNN rnn.py:551-668 | FastGRNN rnn.py:670-707
  FastGRNNCUDA rnn.py:738-889 | FastGRNNFunction rnn.py:891-905
  FastGRNNUnrollFunction rnn.py:907-972 | onnx_exportable_rnn rnn.py:19-38
  FastGRNNBatchNormCell rnn.py:316-452 | FastGRNNBatchNorm rnn.py:709-734 (eval mode folded into the recurrence)

There is no CPU path: parameters/inputs must live on a CUDA device, otherwise the
forward raises ``RuntimeError``.  Reference call-site defects that this module
does not reproduce: D1 (stray ``training=`` kwarg), D2 (``self.W`` with wRank),
D4 (``fastgrnn_cuda`` unbound), D6 (tanh-gate gradient), D10 (sparsify no-op).
"""
from __future__ import annotations

import torch
import torch.nn as nn
from torch.autograd import Function
from torch.autograd.function import once_differentiable

from . import engine, fastgrnn_cuda, ref_utils
from . import ref_utils as utils   # the reference module refers to ``utils.findCUDA`` (rnn.py:476)

__all__ = ["gen_nonlinearity", "RNNCell", "FastGRNNCell", "FastGRNNCUDACell", "BaseRNN", "FastGRNN",
           "FastGRNNCUDA", "FastGRNNBatchNorm", "FastGRNNBatchNormCell", "FastGRNNFunction", "FastGRNNUnrollFunction",
           "onnx_exportable_rnn", "fastgrnn_cuda"]

NON_LINEARITY = {"sigmoid": 0, "relu": 1, "tanh": 2}                      # rnn.py:478, rnn.py:751


def onnx_exportable_rnn(input, fargs, cell, output):
    """rnn.py:19-38: symbolic-only autograd Function used by the (disabled) ONNX export."""
    class RNNSymbolic(Function):
        @staticmethod
        def symbolic(g, *fargs):
            return g.op(cell.name, *fargs, outputs=1, hidden_size_i=cell.state_size,
                        wRank_i=cell.wRank, uRank_i=cell.uRank,
                        gate_nonlinearity_s=cell.gate_nonlinearity,
                        update_nonlinearity_s=cell.update_nonlinearity)

        @staticmethod
        def forward(ctx, *fargs):
            return output

        @staticmethod
        def backward(ctx, *gargs, **gkwargs):
            raise RuntimeError("FIXME: Traced RNNs don't support backward")

    return RNNSymbolic.apply(input, *fargs)


def gen_nonlinearity(A, nonlinearity):
    """rnn.py:40-67 (elementwise helper kept for API parity; the engine applies the same
    functions inside its kernels).  ``relu`` raises in the reference (rnn.py:52, D3); here it works."""
    if nonlinearity == "tanh":
        return torch.tanh(A)
    if nonlinearity == "sigmoid":
        return torch.sigmoid(A)
    if nonlinearity == "relu":
        return torch.relu(A)
    if nonlinearity == "quantTanh":
        return torch.clamp(A, -1.0, 1.0)
    if nonlinearity == "quantSigm":
        return torch.clamp((A + 1.0) / 2.0, 0.0, 1.0)
    if nonlinearity == "quantSigm4":
        return torch.clamp((A + 2.0) / 4.0, 0.0, 1.0)
    if not callable(nonlinearity):
        raise ValueError("nonlinearity is either a callable or a value " +
                         "['tanh', 'sigmoid', 'relu', 'quantTanh', " + "'quantSigm'")
    return nonlinearity(A)


# ----------------------------------------------------------------------------------------------
# the one autograd op everything routes through
# ----------------------------------------------------------------------------------------------
_PARAM_ORDER = ("bias_gate", "bias_update", "zeta", "nu", "W", "U", "W1", "W2", "U1", "U2")


class _Recurrence(Function):
    """T-step FastGRNN recurrence as a single differentiable op (any layout)."""

    @staticmethod
    def forward(ctx, x, h0, bias_gate, bias_update, zeta, nu, W, U, W1, W2, U1, U2, cfg):
        layout, batch_first, gate_nl, update_nl = cfg
        params = dict(zip(_PARAM_ORDER, (bias_gate, bias_update, zeta, nu, W, U, W1, W2, U1, U2)))
        need = any(ctx.needs_input_grad)
        out, z_s, c_s, _ = engine.forward(x, params, h0, layout=layout, batch_first=batch_first,
                                          gate_nl=gate_nl, update_nl=update_nl, save_for_backward=need)
        if need:
            ctx.cfg = cfg
            ctx.has = [t is not None and t.numel() > 0 for t in (W, U, W1, W2, U1, U2)]
            ctx.save_for_backward(x, out, z_s, c_s, h0, bias_gate, bias_update, zeta, nu, W, U, W1, W2, U1, U2)
        return out

    @staticmethod
    @once_differentiable
    def backward(ctx, grad_out):
        layout, batch_first, gate_nl, update_nl = ctx.cfg
        x, out, z_s, c_s, h0, bias_gate, bias_update, zeta, nu, W, U, W1, W2, U1, U2 = ctx.saved_tensors
        params = dict(zip(_PARAM_ORDER, (bias_gate, bias_update, zeta, nu, W, U, W1, W2, U1, U2)))
        ng = ctx.needs_input_grad
        g = engine.backward(grad_out, x, out, z_s, c_s, params, h0, layout=layout, batch_first=batch_first,
                            gate_nl=gate_nl, update_nl=update_nl, need_dx=ng[0], need_dh0=ng[1],
                            need_params=any(ng[2:12]))
        dx = g.get("x")
        if dx is not None and dx.dtype != x.dtype:
            dx = dx.to(x.dtype)
        grads = [dx, g.get("h0") if h0 is not None else None]
        for i, k in enumerate(_PARAM_ORDER):
            grads.append(g.get(k) if ng[2 + i] else None)
        return tuple(grads) + (None,)


def _mat(t):
    return t if (t is not None and t.numel() > 0) else None


def _recurrence(x, h0, owner, layout, batch_first, gate_nl, update_nl):
    return _Recurrence.apply(x, h0, owner.bias_gate, owner.bias_update, owner.zeta, owner.nu,
                             _mat(getattr(owner, "W", None)), _mat(getattr(owner, "U", None)),
                             _mat(getattr(owner, "W1", None)), _mat(getattr(owner, "W2", None)),
                             _mat(getattr(owner, "U1", None)), _mat(getattr(owner, "U2", None)),
                             (layout, bool(batch_first), gate_nl, update_nl))


def _require_cuda(device, what):
    if device.type != "cuda":
        raise RuntimeError("%s: parameters are on %s; the kws_b200 FastGRNN path runs on CUDA (sm_100a) "
                           "only -- there is no CPU fallback. Move the module to a CUDA device." % (what, device))


# ----------------------------------------------------------------------------------------------
# cells
# ----------------------------------------------------------------------------------------------
class RNNCell(nn.Module):
    """rnn.py:69-190: common bookkeeping of the EdgeML cells."""

    def __init__(self, input_size, hidden_size, gate_nonlinearity, update_nonlinearity,
                 num_W_matrices, num_U_matrices, num_biases, wRank=None, uRank=None,
                 wSparsity=1.0, uSparsity=1.0):
        super(RNNCell, self).__init__()
        self._input_size = input_size
        self._hidden_size = hidden_size
        self._gate_nonlinearity = gate_nonlinearity
        self._update_nonlinearity = update_nonlinearity
        self._num_W_matrices = num_W_matrices
        self._num_U_matrices = num_U_matrices
        self._num_biases = num_biases
        self._num_weight_matrices = [self._num_W_matrices, self._num_U_matrices, self._num_biases]
        self._wRank = wRank
        self._uRank = uRank
        self._wSparsity = wSparsity
        self._uSparsity = uSparsity
        self.oldmats = []

    state_size = property(lambda self: self._hidden_size)
    input_size = property(lambda self: self._input_size)
    output_size = property(lambda self: self._hidden_size)
    gate_nonlinearity = property(lambda self: self._gate_nonlinearity)
    update_nonlinearity = property(lambda self: self._update_nonlinearity)
    wRank = property(lambda self: self._wRank)
    uRank = property(lambda self: self._uRank)
    num_W_matrices = property(lambda self: self._num_W_matrices)
    num_U_matrices = property(lambda self: self._num_U_matrices)
    num_weight_matrices = property(lambda self: self._num_weight_matrices)

    @property
    def name(self):
        raise NotImplementedError()

    def forward(self, input, state):
        raise NotImplementedError()

    def getVars(self):
        raise NotImplementedError()

    def get_model_size(self):
        """rnn.py:143-163: 4 bytes x (2 + nnz of every matrix at its target sparsity + dense biases)."""
        return _model_size(self.getVars(), self._num_W_matrices, self._num_U_matrices,
                           self._wSparsity, self._uSparsity)

    def copy_previous_UW(self):
        """rnn.py:165-172: snapshot the W/U matrices (support for sparsifyWithSupport)."""
        mats = self.getVars()
        n = self._num_W_matrices + self._num_U_matrices
        self.oldmats = [mats[i].detach().clone() for i in range(n)]

    def sparsify(self):
        """rnn.py:174-183: hard-threshold W/U to their target sparsity, in place on the device
        (the reference round-trips through numpy and only writes back full-rank W/U)."""
        _sparsify(self)

    def sparsifyWithSupport(self):
        """rnn.py:185-189: re-apply the support of the last ``sparsify`` snapshot."""
        _sparsify_with_support(self)


def _model_size(mats, nW, nU, wSparsity, uSparsity):
    totalnnz = 2   # zeta and nu (rnn.py:151)
    for i in range(0, nW):
        totalnnz += ref_utils.countNNZ(mats[i], wSparsity)
    for i in range(nW, nW + nU):
        totalnnz += ref_utils.countNNZ(mats[i], uSparsity)
    for i in range(nW + nU, len(mats)):
        totalnnz += ref_utils.countNNZ(mats[i], False)
    return totalnnz * 4


def _sparsify(owner):
    mats = owner.getVars()
    endW = owner._num_W_matrices
    endU = endW + owner._num_U_matrices
    for i in range(0, endW):
        ref_utils.hard_threshold_(mats[i].data, owner._wSparsity)
    for i in range(endW, endU):
        ref_utils.hard_threshold_(mats[i].data, owner._uSparsity)
    owner.copy_previous_UW()


def _sparsify_with_support(owner):
    mats = owner.getVars()
    endU = owner._num_W_matrices + owner._num_U_matrices
    if len(owner.oldmats) != endU:
        raise RuntimeError("sparsifyWithSupport called before sparsify/copy_previous_UW")
    for i in range(0, endU):
        ref_utils.copy_support_(owner.oldmats[i], mats[i].data)


class FastGRNNCell(RNNCell):
    """rnn.py:192-313.  Parameters in the oracle layout: W [I,H] | W1 [I,rW], W2 [rW,H];
    U [H,H] | U1 [H,rU], U2 [rU,H]; bias_gate/bias_update [1,H]; zeta/nu [1,1].

        z_t = gate_nl(W x_t + U h_{t-1} + B_g);  h~_t = update_nl(W x_t + U h_{t-1} + B_h)
        h_t = z_t*h_{t-1} + (sigmoid(zeta)(1-z_t) + sigmoid(nu))*h~_t
    """

    def __init__(self, input_size, hidden_size, gate_nonlinearity="sigmoid",
                 update_nonlinearity="tanh", wRank=None, uRank=None,
                 wSparsity=1.0, uSparsity=1.0, zetaInit=1.0, nuInit=-4.0, name="FastGRNN"):
        super(FastGRNNCell, self).__init__(input_size, hidden_size, gate_nonlinearity,
                                           update_nonlinearity, 1, 1, 2, wRank, uRank,
                                           wSparsity, uSparsity)
        self._zetaInit = zetaInit
        self._nuInit = nuInit
        if wRank is not None:
            self._num_W_matrices += 1
            self._num_weight_matrices[0] = self._num_W_matrices
        if uRank is not None:
            self._num_U_matrices += 1
            self._num_weight_matrices[1] = self._num_U_matrices
        self._name = name
        # same draw order as rnn.py:246-261 so a seeded construction gives identical parameters
        if wRank is None:
            self.W = nn.Parameter(0.1 * torch.randn([input_size, hidden_size]))
        else:
            self.W1 = nn.Parameter(0.1 * torch.randn([input_size, wRank]))
            self.W2 = nn.Parameter(0.1 * torch.randn([wRank, hidden_size]))
        if uRank is None:
            self.U = nn.Parameter(0.1 * torch.randn([hidden_size, hidden_size]))
        else:
            self.U1 = nn.Parameter(0.1 * torch.randn([hidden_size, uRank]))
            self.U2 = nn.Parameter(0.1 * torch.randn([uRank, hidden_size]))
        self.bias_gate = nn.Parameter(torch.ones([1, hidden_size]))
        self.bias_update = nn.Parameter(torch.ones([1, hidden_size]))
        self.zeta = nn.Parameter(self._zetaInit * torch.ones([1, 1]))
        self.nu = nn.Parameter(self._nuInit * torch.ones([1, 1]))

    @property
    def name(self):
        return self._name

    @property
    def cellType(self):
        return "FastGRNN"

    def _device(self):
        return self.bias_gate.device

    def forward(self, input, state, training=True):
        """One step (rnn.py:273-297): input [B,I], state [B,H] -> new_h [B,H]."""
        device = self._device()
        _require_cuda(device, "FastGRNNCell.forward")
        input = input.to(device)                                         # rnn.py:275-276
        state = state.to(device)
        out = _recurrence(input.unsqueeze(0), state.contiguous(), self, "IH", False,
                          self._gate_nonlinearity, self._update_nonlinearity)
        return out[0]

    def unroll(self, input, h0, batch_first):
        """All T states for a whole sequence in one launch (what BaseRNN.forward uses)."""
        device = self._device()
        _require_cuda(device, "FastGRNN.forward")
        return _recurrence(input.to(device), h0, self, "IH", batch_first,
                           self._gate_nonlinearity, self._update_nonlinearity)

    def getVars(self):
        Vars = []
        if self._num_W_matrices == 1:
            Vars.append(self.W)
        else:
            Vars.extend([self.W1, self.W2])
        if self._num_U_matrices == 1:
            Vars.append(self.U)
        else:
            Vars.extend([self.U1, self.U2])
        Vars.extend([self.bias_gate, self.bias_update])
        Vars.extend([self.zeta, self.nu])
        return Vars


def _make_cuda_params(mod, input_size, hidden_size, wRank, uRank, zetaInit, nuInit, device):
    """Parameter set of the CUDA-layout modules (rnn.py:494-517 / :782-805): transposed matrices,
    unused rank slots are plain CPU ``torch.empty(0)`` attributes, created directly on the GPU."""
    if wRank is None:
        mod.W = nn.Parameter(0.1 * torch.randn([hidden_size, input_size], device=device))
        mod.W1 = torch.empty(0)
        mod.W2 = torch.empty(0)
    else:
        mod.W = torch.empty(0)
        mod.W1 = nn.Parameter(0.1 * torch.randn([wRank, input_size], device=device))
        mod.W2 = nn.Parameter(0.1 * torch.randn([hidden_size, wRank], device=device))
    if uRank is None:
        mod.U = nn.Parameter(0.1 * torch.randn([hidden_size, hidden_size], device=device))
        mod.U1 = torch.empty(0)
        mod.U2 = torch.empty(0)
    else:
        mod.U = torch.empty(0)
        mod.U1 = nn.Parameter(0.1 * torch.randn([uRank, hidden_size], device=device))
        mod.U2 = nn.Parameter(0.1 * torch.randn([hidden_size, uRank], device=device))
    mod.bias_gate = nn.Parameter(torch.ones([1, hidden_size], device=device))
    mod.bias_update = nn.Parameter(torch.ones([1, hidden_size], device=device))
    mod.zeta = nn.Parameter(zetaInit * torch.ones([1, 1], device=device))
    mod.nu = nn.Parameter(nuInit * torch.ones([1, 1], device=device))


def _cuda_vars(mod):
    Vars = []
    if mod._num_W_matrices == 1:
        Vars.append(mod.W)
    else:
        Vars.extend([mod.W1, mod.W2])
    if mod._num_U_matrices == 1:
        Vars.append(mod.U)
    else:
        Vars.extend([mod.U1, mod.U2])
    Vars.extend([mod.bias_gate, mod.bias_update, mod.zeta, mod.nu])
    return Vars


def _require_gpu_available():
    # rnn.py:476-477 / :749-750 raise exactly this when no CUDA is found
    if utils.findCUDA() is None or not torch.cuda.is_available():
        raise Exception('FastGRNNCUDA is supported only on GPU devices.')


class FastGRNNCUDACell(RNNCell):
    """rnn.py:454-549: single-step module in the CUDA layout (W [H,I], U [H,H], W1 [rW,I],
    W2 [H,rW], U1 [rU,H], U2 [H,rU]); update nonlinearity fixed to tanh (cu:57)."""

    def __init__(self, input_size, hidden_size, gate_nonlinearity="sigmoid",
                 update_nonlinearity="tanh", wRank=None, uRank=None, zetaInit=1.0, nuInit=-4.0,
                 wSparsity=1.0, uSparsity=1.0, name="FastGRNNCUDACell"):
        super(FastGRNNCUDACell, self).__init__(input_size, hidden_size, gate_nonlinearity,
                                               update_nonlinearity, 1, 1, 2, wRank, uRank,
                                               wSparsity, uSparsity)
        _require_gpu_available()
        self._zetaInit = zetaInit
        self._nuInit = nuInit
        self._name = name
        self.device = torch.device("cuda")
        if wRank is not None:
            self._num_W_matrices += 1
            self._num_weight_matrices[0] = self._num_W_matrices
        if uRank is not None:
            self._num_U_matrices += 1
            self._num_weight_matrices[1] = self._num_U_matrices
        _make_cuda_params(self, input_size, hidden_size, wRank, uRank, zetaInit, nuInit, self.device)
        self._gate_non_linearity = NON_LINEARITY[gate_nonlinearity]

    @property
    def name(self):
        return self._name

    @property
    def cellType(self):
        return "FastGRNNCUDACell"

    def forward(self, input, state):
        dev = self.bias_gate.device
        _require_cuda(dev, "FastGRNNCUDACell.forward")
        if not input.is_cuda:
            input = input.to(dev)     # the reference discards this result (rnn.py:529-532, D9)
        if not state.is_cuda:
            state = state.to(dev)
        return FastGRNNFunction.apply(input, self.bias_gate, self.bias_update, self.zeta, self.nu, state,
                                      self.W, self.U, self.W1, self.W2, self.U1, self.U2,
                                      self._gate_non_linearity)

    def getVars(self):
        return _cuda_vars(self)


# ----------------------------------------------------------------------------------------------
# unrollers / modules
# ----------------------------------------------------------------------------------------------
class BaseRNN(nn.Module):
    """rnn.py:551-668: ``static_rnn``-style unroller.  For FastGRNN cells the whole loop
    (rnn.py:620-630 batch-first, :658-668 time-major) is one engine call; input
    [T,B,F] by default, [B,T,F] with ``batch_first``.  ``hiddenState`` is [num_directions,B,H]
    (zeros if None, rnn.py:588-591) and, like the reference's in-place loop, holds the final
    state(s) on return."""

    def __init__(self, cell: RNNCell, batch_first=False, cell_reverse: RNNCell = None, bidirectional=False):
        super(BaseRNN, self).__init__()
        self.RNNCell = cell
        self._batch_first = batch_first
        self._bidirectional = bidirectional
        if cell_reverse is not None:
            self.RNNCell_reverse = cell_reverse
        elif self._bidirectional:
            self.RNNCell_reverse = cell

    def getVars(self):
        return self.RNNCell.getVars()

    def forward(self, input, hiddenState=None, cellState=None, training=True):
        cell = self.RNNCell
        if not isinstance(cell, FastGRNNCell):
            raise NotImplementedError("kws_b200.BaseRNN unrolls FastGRNN cells only (got %s); other "
                                      "EdgeML cells are outside the accelerated path" % type(cell).__name__)
        self.device = input.device
        self.num_directions = 2 if self._bidirectional else 1
        tdim = 1 if self._batch_first else 0
        B = input.shape[0] if self._batch_first else input.shape[1]
        if hiddenState is not None:
            if hiddenState.dim() != 3 or hiddenState.shape[0] < self.num_directions:
                raise RuntimeError("hiddenState must be [num_directions, batch, hidden] (rnn.py:588-591), got %s"
                                   % (tuple(hiddenState.shape),))
            if hiddenState.shape[1] != B:
                raise RuntimeError("hiddenState batch %d != input batch %d" % (hiddenState.shape[1], B))
        dev = cell._device()

        def h0_of(d):
            # clone like the reference's ``hiddenState[0].clone()`` (rnn.py:621): the caller's tensor is
            # overwritten below and must not alias what autograd saved
            return None if hiddenState is None else hiddenState[d].to(dev, torch.float32).clone().contiguous()

        kw = {"training": training} if isinstance(cell, FastGRNNBatchNormCell) else {}
        out = cell.unroll(input, h0_of(0), self._batch_first, **kw)
        outs = [out]
        if self._bidirectional:
            # the reverse direction consumes input[T-1-i] at step i and stores its state at index i
            # (rnn.py:623-626 / :661-664), i.e. states are kept in processing order
            outs.append(self.RNNCell_reverse.unroll(input.flip(tdim), h0_of(1), self._batch_first))
        if hiddenState is not None and input.shape[tdim] > 0:
            with torch.no_grad():
                for d, o in enumerate(outs):
                    hiddenState[d].copy_(o.select(tdim, o.shape[tdim] - 1))   # rnn.py:621 mutates in place
        return outs[0] if not self._bidirectional else torch.cat(outs, -1)


class FastGRNN(nn.Module):
    """rnn.py:670-707 ("Equivalent to nn.FastGRNN using FastGRNNCell")."""

    def __init__(self, input_size, hidden_size, gate_nonlinearity="sigmoid",
                 update_nonlinearity="tanh", wRank=None, uRank=None,
                 wSparsity=1.0, uSparsity=1.0, zetaInit=1.0, nuInit=-4.0,
                 batch_first=False, bidirectional=False, is_shared_bidirectional=True):
        super(FastGRNN, self).__init__()
        self._bidirectional = bidirectional
        self._batch_first = batch_first
        self._is_shared_bidirectional = is_shared_bidirectional
        self.cell = FastGRNNCell(input_size, hidden_size, gate_nonlinearity=gate_nonlinearity,
                                 update_nonlinearity=update_nonlinearity, wRank=wRank, uRank=uRank,
                                 wSparsity=wSparsity, uSparsity=uSparsity, zetaInit=zetaInit, nuInit=nuInit)
        self.unrollRNN = BaseRNN(self.cell, batch_first=self._batch_first, bidirectional=self._bidirectional)
        self.training = True
        if self._bidirectional is True and self._is_shared_bidirectional is False:
            self.cell_reverse = FastGRNNCell(input_size, hidden_size, gate_nonlinearity=gate_nonlinearity,
                                             update_nonlinearity=update_nonlinearity, wRank=wRank, uRank=uRank,
                                             wSparsity=wSparsity, uSparsity=uSparsity,
                                             zetaInit=zetaInit, nuInit=nuInit)
            # the reference passes cell_reverse into the batch_first slot here (rnn.py:697, D11)
            self.unrollRNN = BaseRNN(self.cell, batch_first=self._batch_first, cell_reverse=self.cell_reverse,
                                     bidirectional=self._bidirectional)

    def getVars(self):
        return self.unrollRNN.getVars()

    def forward(self, input, hiddenState=None, cellState=None):
        return self.unrollRNN(input, hiddenState, cellState, training=self.training)

    def train(self, mode=True):
        self.training = mode
        super(FastGRNN, self).train(mode)
        return self


class FastGRNNBatchNormCell(FastGRNNCell):
    """rnn.py:316-452: FastGRNN cell with four ``nn.BatchNorm1d`` layers -- on W x (``bn_w``), on U h (``bn_u``) and on the
    two pre-activations (``bn_gate``, ``bn_update``).  Same parameter names / creation order as the reference, so its
    checkpoints (``model_batchnorm/FastGRNNBatchNorm_KeywordSpotter.pt``-shaped state_dicts) load.

    EVAL MODE runs on the engine: with running statistics every BatchNorm layer is a per-unit affine map
    ``a*v + b`` (a = weight/sqrt(var+eps), b = bias - mean*a), so

        pre       = x.(W diag(a_w)) + h.(U diag(a_u))                      columns of W / U scaled
        z         = gate  (a_g * pre + [a_g (b_w + b_u + bias_gate)   + b_g])
        c         = update(a_c * pre + [a_c (b_w + b_u + bias_update) + b_c])

    i.e. the plain recurrence with folded weights, folded biases and the per-unit ``gate_scale`` / ``update_scale`` of the
    C ABI.  Training mode (per-time-step BATCH statistics, rnn.py:395-405 with ``training=True``) is a different
    computation -- a cross-batch reduction inside every step -- and is not accelerated: it raises."""

    def __init__(self, input_size, hidden_size, gate_nonlinearity="sigmoid", update_nonlinearity="tanh",
                 wRank=None, uRank=None, wSparsity=1.0, uSparsity=1.0, zetaInit=1.0, nuInit=-4.0,
                 name="FastGRNNBatchNorm"):
        super(FastGRNNBatchNormCell, self).__init__(input_size, hidden_size, gate_nonlinearity, update_nonlinearity,
                                                    wRank, uRank, wSparsity, uSparsity, zetaInit, nuInit, name)
        self.bn_w = nn.BatchNorm1d(hidden_size)                          # rnn.py:363-366
        self.bn_u = nn.BatchNorm1d(hidden_size)
        self.bn_gate = nn.BatchNorm1d(hidden_size)
        self.bn_update = nn.BatchNorm1d(hidden_size)

    @property
    def cellType(self):
        return "FastGRNNBatchNorm"

    @staticmethod
    def _affine(bn):
        a = bn.weight.detach().float() * torch.rsqrt(bn.running_var.float() + bn.eps)
        return a, bn.bias.detach().float() - bn.running_mean.float() * a

    def folded_params(self):
        """The eval-mode cell as plain FastGRNN parameters + pre-activation scales (all on the cell's device)."""
        with torch.no_grad():
            aw, bw = self._affine(self.bn_w)
            au, bu = self._affine(self.bn_u)
            ag, bg = self._affine(self.bn_gate)
            ac, bc = self._affine(self.bn_update)
            p = {}
            if self._wRank is None:
                p["W"] = (self.W.detach() * aw).contiguous()
            else:
                p["W1"] = self.W1.detach().contiguous(); p["W2"] = (self.W2.detach() * aw).contiguous()
            if self._uRank is None:
                p["U"] = (self.U.detach() * au).contiguous()
            else:
                p["U1"] = self.U1.detach().contiguous(); p["U2"] = (self.U2.detach() * au).contiguous()
            shift = bw + bu
            p["bias_gate"] = (ag * (shift + self.bias_gate.detach()[0]) + bg).unsqueeze(0).contiguous()
            p["bias_update"] = (ac * (shift + self.bias_update.detach()[0]) + bc).unsqueeze(0).contiguous()
            p["gate_scale"] = ag.unsqueeze(0).contiguous()
            p["update_scale"] = ac.unsqueeze(0).contiguous()
            p["zeta"] = self.zeta.detach().contiguous(); p["nu"] = self.nu.detach().contiguous()
        return p

    def _eval_mode(self, training):
        # the reference picks eval statistics when called with training=False (``self.bn_w.eval()(wComp)``) and otherwise
        # follows the BatchNorm modules' own mode (rnn.py:395-396)
        return (not training) or not (self.bn_w.training or self.bn_u.training or self.bn_gate.training or self.bn_update.training)

    def unroll(self, input, h0, batch_first, training=True):
        device = self._device()
        _require_cuda(device, "FastGRNNBatchNorm.forward")
        if not self._eval_mode(training):
            raise NotImplementedError("FastGRNNBatchNorm in training mode normalises every time step with BATCH statistics "
                                      "(rnn.py:395-405), which this engine does not accelerate; call with training=False "
                                      "or put the module in eval() mode (running statistics fold into the recurrence)")
        if not training:
            for bn in (self.bn_w, self.bn_u, self.bn_gate, self.bn_update):
                bn.eval()                                                # the reference's ``self.bn_w.eval()(...)`` has the same side effect
        out, _, _, _ = engine.forward(input.to(device), self.folded_params(), h0, layout="IH", batch_first=batch_first,
                                      gate_nl=self._gate_nonlinearity, update_nl=self._update_nonlinearity)
        return out

    def forward(self, input, state, training=True):
        """One step (rnn.py:377-410): input [B,I], state [B,H] -> new_h [B,H]."""
        device = self._device()
        return self.unroll(input.to(device).unsqueeze(0), state.to(device).contiguous(), False, training=training)[0]


class FastGRNNBatchNorm(nn.Module):
    """rnn.py:709-734: "FastGRNN with Batch Normalization wrapper class".  ``forward(input, hiddenState=None,
    training=True)``; eval mode is accelerated (see FastGRNNBatchNormCell), training mode raises."""

    def __init__(self, input_size, hidden_size, gate_nonlinearity="sigmoid", update_nonlinearity="tanh",
                 wRank=None, uRank=None, wSparsity=1.0, uSparsity=1.0, zetaInit=1.0, nuInit=-4.0, batch_first=False):
        super(FastGRNNBatchNorm, self).__init__()
        self.cell = FastGRNNBatchNormCell(input_size, hidden_size, gate_nonlinearity=gate_nonlinearity,
                                          update_nonlinearity=update_nonlinearity, wRank=wRank, uRank=uRank,
                                          wSparsity=wSparsity, uSparsity=uSparsity, zetaInit=zetaInit, nuInit=nuInit)
        self.unrollRNN = BaseRNN(self.cell, batch_first=batch_first)
        self.training = True

    def getVars(self):
        return self.unrollRNN.getVars()

    def forward(self, input, hiddenState=None, training=True):
        return self.unrollRNN(input, hiddenState, training=training)

    def train(self, mode=True):
        self.training = mode
        super(FastGRNNBatchNorm, self).train(mode)
        return self


class FastGRNNCUDA(nn.Module):
    """rnn.py:738-889: unrolled CUDA-layout module.  ``forward(input, hiddenState=None,
    cell_state=None)``: input [T,B,F] (or [B,T,F] with ``batch_first``), hiddenState [B,H]."""

    def __init__(self, input_size, hidden_size, gate_nonlinearity="sigmoid",
                 update_nonlinearity="tanh", wRank=None, uRank=None,
                 wSparsity=1.0, uSparsity=1.0, zetaInit=1.0, nuInit=-4.0,
                 batch_first=False, name="FastGRNNCUDA"):
        super(FastGRNNCUDA, self).__init__()
        _require_gpu_available()
        self.cell = FastGRNNCUDACell(input_size, hidden_size, gate_nonlinearity=gate_nonlinearity,
                                     update_nonlinearity=update_nonlinearity, wRank=wRank, uRank=uRank,
                                     wSparsity=wSparsity, uSparsity=uSparsity,
                                     zetaInit=zetaInit, nuInit=nuInit)
        self._input_size = input_size
        self._hidden_size = hidden_size
        self._zetaInit = zetaInit
        self._nuInit = nuInit
        self._name = name
        self._num_W_matrices = 1
        self._num_U_matrices = 1
        self._num_biases = 2
        self._num_weight_matrices = [self._num_W_matrices, self._num_U_matrices, self._num_biases]
        self._wRank = wRank
        self._uRank = uRank
        self._wSparsity = wSparsity
        self._uSparsity = uSparsity
        self.oldmats = []
        self.device = torch.device("cuda")
        self.batch_first = batch_first
        if wRank is not None:
            self._num_W_matrices += 1
            self._num_weight_matrices[0] = self._num_W_matrices
        if uRank is not None:
            self._num_U_matrices += 1
            self._num_weight_matrices[1] = self._num_U_matrices
        _make_cuda_params(self, input_size, hidden_size, wRank, uRank, zetaInit, nuInit, self.device)
        self._gate_non_linearity = NON_LINEARITY[gate_nonlinearity]

    @property
    def name(self):
        return self._name

    def forward(self, input, hiddenState=None, cell_state=None):
        dev = self.bias_gate.device
        _require_cuda(dev, "FastGRNNCUDA.forward")
        if not input.is_cuda:
            input = input.to(dev)
        # The reference transposes a batch-first input to a contiguous [T,B,F] copy and transposes the
        # result back (rnn.py:812-813, :823-824); the engine reads/writes [B,T,*] through strides, so
        # the view below costs nothing and the returned tensor has the same shape and values.
        if self.batch_first is True:
            input = input.transpose(0, 1)
        if hiddenState is None:
            hiddenState = torch.zeros([input.shape[1], self._hidden_size], device=dev)   # rnn.py:816-818
        if not hiddenState.is_cuda:
            hiddenState = hiddenState.to(dev)
        result = FastGRNNUnrollFunction.apply(input, self.bias_gate, self.bias_update, self.zeta, self.nu,
                                              hiddenState, self.W, self.U, self.W1, self.W2, self.U1, self.U2,
                                              self._gate_non_linearity)
        if self.batch_first is True:
            return result.transpose(0, 1)
        return result

    def getVars(self):
        return _cuda_vars(self)

    def get_model_size(self):
        return _model_size(self.getVars(), self._num_W_matrices, self._num_U_matrices,
                           self._wSparsity, self._uSparsity)

    def copy_previous_UW(self):
        mats = self.getVars()
        n = self._num_W_matrices + self._num_U_matrices
        self.oldmats = [mats[i].detach().clone() for i in range(n)]

    def sparsify(self):
        """rnn.py:875-884 thresholds copies and never writes the parameters (D10); this applies the
        intended in-place hard threshold on whatever device the parameters are on."""
        _sparsify(self)

    def sparsifyWithSupport(self):
        _sparsify_with_support(self)


# ----------------------------------------------------------------------------------------------
# reference-named autograd Functions (fixed 13-argument signatures, rnn.py:891-972)
# ----------------------------------------------------------------------------------------------
class FastGRNNFunction(Function):
    """Single step (rnn.py:891-905): ``apply(input[B,I], bias_gate, bias_update, zeta, nu, old_h[B,H],
    w, u, w1, w2, u1, u2, gate_non_linearity:int) -> new_h[B,H]``."""

    @staticmethod
    def forward(ctx, input, bias_gate, bias_update, zeta, nu, old_h, w, u, w1, w2, u1, u2, gate_non_linearity):
        outputs = fastgrnn_cuda.forward(input.contiguous(), w, u, bias_gate, bias_update, zeta, nu,
                                        old_h.contiguous(), gate_non_linearity, w1, w2, u1, u2)
        new_h = outputs[0]
        ctx.save_for_backward(input, old_h, zeta, nu, w, u, outputs[1], outputs[2], w1, w2, u1, u2)
        ctx.non_linearity = gate_non_linearity
        return new_h

    @staticmethod
    @once_differentiable
    def backward(ctx, grad_h):
        input, old_h, zeta, nu, w, u, z, h_prime, w1, w2, u1, u2 = ctx.saved_tensors
        outputs = fastgrnn_cuda.backward(grad_h.contiguous(), input.contiguous(), old_h.contiguous(), zeta, nu,
                                         w, u, z, h_prime, w1, w2, u1, u2, ctx.non_linearity)
        return tuple(outputs + [None])


class FastGRNNUnrollFunction(Function):
    """T steps (rnn.py:907-972): ``apply(input[T,B,I], bias_gate, bias_update, zeta, nu, old_h[B,H],
    w, u, w1, w2, u1, u2, gate_non_linearity:int) -> hidden_states[T,B,H]``; backward returns
    ``(d_input, d_bias_gate, d_bias_update, d_zeta, d_nu, d_old_h, d_w, d_u, d_w1, d_w2, d_u1, d_u2, None)``.

    ``input`` may be any [T,B,I] view whose feature stride is 1 (the reference calls
    ``.contiguous()``, rnn.py:910; the engine takes strides instead).  When no input needs a
    gradient (inference) the z/h~ save buffers are not written at all."""

    @staticmethod
    def forward(ctx, input, bias_gate, bias_update, zeta, nu, old_h, w, u, w1, w2, u1, u2, gate_non_linearity):
        params = {"W": w, "U": u, "W1": w1, "W2": w2, "U1": u1, "U2": u2,
                  "bias_gate": bias_gate, "bias_update": bias_update, "zeta": zeta, "nu": nu}
        need = any(ctx.needs_input_grad)
        hidden_states, z_s, h_prime_s, _ = engine.forward(
            input, params, old_h.contiguous(), layout="HI", batch_first=False,
            gate_nl=int(gate_non_linearity), update_nl="tanh", save_for_backward=need)
        if need:
            ctx.save_for_backward(input, hidden_states, zeta, nu, w, u, z_s, h_prime_s, old_h, w1, w2, u1, u2)
            ctx.gate_non_linearity = gate_non_linearity
        return hidden_states

    @staticmethod
    @once_differentiable
    def backward(ctx, grad_h):
        input, hidden_states, zeta, nu, w, u, z_s, h_prime_s, old_h, w1, w2, u1, u2 = ctx.saved_tensors
        H = old_h.shape[1]
        dummy_bias = torch.empty((1, H), dtype=torch.float32, device=old_h.device)
        params = {"W": w, "U": u, "W1": w1, "W2": w2, "U1": u1, "U2": u2,
                  "bias_gate": dummy_bias, "bias_update": dummy_bias, "zeta": zeta, "nu": nu}
        ng = ctx.needs_input_grad
        g = engine.backward(grad_h, input, hidden_states, z_s, h_prime_s, params, old_h.contiguous(),
                            layout="HI", batch_first=False, gate_nl=int(ctx.gate_non_linearity),
                            update_nl="tanh", need_dx=ng[0], need_dh0=ng[5])
        low_w, low_u = w1.size(0) != 0, u1.size(0) != 0
        e = torch.empty(0)
        dx = g.get("x")
        if dx is not None and dx.dtype != input.dtype:
            dx = dx.to(input.dtype)
        # order of cuda/fastgrnn_cuda_kernel.cu:556 (+ None for the int argument, rnn.py:972)
        return (dx, g["bias_gate"], g["bias_update"], g["zeta"], g["nu"], g.get("h0"),
                e if low_w else g["W"], e if low_u else g["U"],
                g["W1"] if low_w else e, g["W2"] if low_w else e,
                g["U1"] if low_u else e, g["U2"] if low_u else e, None)
