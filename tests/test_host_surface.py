"""CPU: the rnn / fastgrnn_cuda operator surface mirrors the reference (names, constructor
keywords, parameter names and shapes, state_dict keys, error behaviour) -- no compute."""
import inspect
import os
import sys

import pytest
import torch

from conftest import ROOT, load_golden
from oracle import ref_shim

from kws_b200 import fastgrnn_cuda as ext
from kws_b200 import rnn as krnn


def test_exported_names():
    for n in ("FastGRNN", "FastGRNNCUDA", "FastGRNNBatchNorm", "onnx_exportable_rnn", "FastGRNNCell", "FastGRNNCUDACell",
              "BaseRNN", "RNNCell", "gen_nonlinearity", "FastGRNNFunction", "FastGRNNUnrollFunction", "fastgrnn_cuda"):
        assert hasattr(krnn, n), n
    for n in ("forward", "backward", "forward_unroll", "backward_unroll"):     # cuda/fastgrnn_cuda.cpp:235-240
        assert callable(getattr(ext, n))


def test_extension_argument_orders():
    # positional orders of cuda/fastgrnn_cuda.cpp:73-86, :109-123, :147-160, :182-197
    assert list(inspect.signature(ext.forward).parameters) == [
        "input", "w", "u", "bias_gate", "bias_update", "zeta", "nu", "old_h", "z_non_linearity", "w1", "w2", "u1", "u2"]
    assert list(inspect.signature(ext.backward).parameters) == [
        "grad_h", "input", "old_h", "zeta", "nu", "w", "u", "z", "h_prime", "w1", "w2", "u1", "u2", "z_non_linearity"]
    assert list(inspect.signature(ext.forward_unroll).parameters) == [
        "input", "w", "u", "bias_gate", "bias_update", "zeta", "nu", "initial_h", "z_non_linearity", "w1", "w2", "u1", "u2"]
    assert list(inspect.signature(ext.backward_unroll).parameters) == [
        "grad_h", "input", "hidden_states", "zeta", "nu", "w", "u", "z", "h_prime", "initial_h", "w1", "w2", "u1", "u2",
        "z_non_linearity"]


@pytest.mark.parametrize("wR,uR", [(None, None), (16, 32), (8, None), (None, 8)])
def test_fastgrnn_parameters_match_reference_layout(wR, uR):
    torch.manual_seed(3)
    m = krnn.FastGRNN(32, 128, wRank=wR, uRank=uR, batch_first=True)
    shapes = {k: tuple(v.shape) for k, v in m.cell.named_parameters()}
    exp = {"bias_gate": (1, 128), "bias_update": (1, 128), "zeta": (1, 1), "nu": (1, 1)}
    exp.update({"W": (32, 128)} if wR is None else {"W1": (32, wR), "W2": (wR, 128)})
    exp.update({"U": (128, 128)} if uR is None else {"U1": (128, uR), "U2": (uR, 128)})
    assert shapes == exp
    # every parameter appears under both prefixes, like the reference
    keys = set(m.state_dict().keys())
    assert keys == {p + k for k in exp for p in ("cell.", "unrollRNN.RNNCell.")}
    assert m.cell.zeta.item() == 1.0 and m.cell.nu.item() == -4.0 and torch.all(m.cell.bias_gate == 1)
    assert m.cell.cellType == "FastGRNN" and m.cell.name == "FastGRNN"
    assert m.cell.num_weight_matrices == [1 if wR is None else 2, 1 if uR is None else 2, 2]
    assert len(m.getVars()) == len(exp)
    # rnn.py:151-162 counts zeta and nu twice (2 up front + once more as trailing getVars entries)
    assert m.cell.get_model_size() == 4 * (sum(v.numel() for v in m.cell.parameters()) + 2)


@pytest.mark.skipif(not ref_shim.available(), reason="/root/reference not present")
@pytest.mark.parametrize("wR,uR", [(None, None), (16, 32)])
def test_seeded_construction_equals_reference(wR, uR):
    rnn_ref, _ = ref_shim.load()
    torch.manual_seed(5)
    a = rnn_ref.FastGRNN(32, 64, wRank=wR, uRank=uR)
    torch.manual_seed(5)
    b = krnn.FastGRNN(32, 64, wRank=wR, uRank=uR)
    sa, sb = a.state_dict(), b.state_dict()
    assert list(sa.keys()) == list(sb.keys())
    for k in sa:
        assert torch.equal(sa[k], sb[k]), k
    b.load_state_dict(sa)        # reference checkpoints load
    assert inspect.signature(rnn_ref.FastGRNN.__init__).parameters.keys() == inspect.signature(krnn.FastGRNN.__init__).parameters.keys()
    assert inspect.signature(rnn_ref.FastGRNNCUDA.__init__).parameters.keys() == inspect.signature(krnn.FastGRNNCUDA.__init__).parameters.keys()
    assert inspect.signature(rnn_ref.FastGRNNCell.__init__).parameters.keys() == inspect.signature(krnn.FastGRNNCell.__init__).parameters.keys()


def test_no_cpu_fallback_and_gpu_only_modules():
    m = krnn.FastGRNN(8, 16)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        m(torch.randn(3, 2, 8))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        m.cell(torch.randn(2, 8), torch.zeros(2, 16))
    if not torch.cuda.is_available():
        with pytest.raises(Exception, match="FastGRNNCUDA is supported only on GPU devices."):   # rnn.py:749-750
            krnn.FastGRNNCUDA(8, 16)
        with pytest.raises(Exception, match="FastGRNNCUDA is supported only on GPU devices."):   # rnn.py:476-477
            krnn.FastGRNNCUDACell(8, 16)
    bn = krnn.FastGRNNBatchNorm(8, 16)                       # constructs on the CPU like the reference; runs on CUDA only
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        bn(torch.randn(3, 2, 8), training=False)
    with pytest.raises(RuntimeError, match="input must be a CUDA tensor"):                       # cpp:69
        e = torch.empty(0)
        ext.forward_unroll(torch.randn(3, 2, 8), torch.randn(16, 8), torch.randn(16, 16), torch.ones(1, 16),
                           torch.ones(1, 16), torch.ones(1, 1), torch.ones(1, 1), torch.zeros(2, 16), 0, e, e, e, e)


def test_gen_nonlinearity_matches_oracle():
    from oracle.fastgrnn_oracle import nonlinearity
    a = torch.linspace(-3, 3, 101)
    for n in ("tanh", "sigmoid", "quantTanh", "quantSigm", "quantSigm4", "relu"):
        assert torch.equal(krnn.gen_nonlinearity(a, n), nonlinearity(a, n)), n
    assert torch.equal(krnn.gen_nonlinearity(a, torch.abs), a.abs())
    with pytest.raises(ValueError):
        krnn.gen_nonlinearity(a, "nope")


def test_sparsify_thresholds_in_place():
    torch.manual_seed(0)
    m = krnn.FastGRNN(16, 32, wSparsity=0.25, uSparsity=0.5)
    m.cell.sparsify()
    assert abs((m.cell.W != 0).float().mean().item() - 0.25) < 0.02
    assert abs((m.cell.U != 0).float().mean().item() - 0.5) < 0.02
    with torch.no_grad():
        m.cell.W.add_(1.0)
    m.cell.sparsifyWithSupport()
    assert abs((m.cell.W != 0).float().mean().item() - 0.25) < 0.02
    assert m.cell.get_model_size() < 4 * sum(v.numel() for v in m.cell.parameters())


@pytest.mark.skipif(not ref_shim.available(), reason="/root/reference not present")
def test_unmodified_model_py_builds_on_compat_rnn():
    """The reference's model.py imports `rnn` by name (model.py:6); with kws_b200/compat first on
    sys.path it constructs its layers from our classes, with the reference's state_dict keys."""
    import importlib
    g = load_golden("model_2layer_256_128")
    saved = {k: sys.modules.pop(k, None) for k in ("rnn", "model", "fastgrnn_cuda")}
    sys.path.insert(0, ref_shim.REF_ROOT)
    sys.path.insert(0, os.path.join(ROOT, "kws_b200", "compat"))
    try:
        model = importlib.import_module("model")
        import rnn as bound
        assert bound.FastGRNN is krnn.FastGRNN
        Model = model.get_model_class()
        mdl = Model("FastGRNN", 32, 2, [256, 128], [None, None], [None, None], [1.0, 1.0], [1.0, 1.0],
                    "sigmoid", "tanh", num_classes=13, linear=True, batch_first=False, apply_softmax=True)
        assert sorted(mdl.state_dict().keys()) == sorted(str(k) for k in g["state_dict_keys"])
        assert isinstance(mdl.rnn_list[0], krnn.FastGRNN)
        assert mdl.get_model_size() > 0
        mdl.tracking = False
        with pytest.raises(RuntimeError, match="no CPU fallback"):   # the forward reaches our engine
            mdl(torch.randn(5, 2, 32))
    finally:
        sys.path.remove(ref_shim.REF_ROOT)
        sys.path.remove(os.path.join(ROOT, "kws_b200", "compat"))
        for k in ("rnn", "model", "fastgrnn_cuda"):
            sys.modules.pop(k, None)
            if saved[k] is not None:
                sys.modules[k] = saved[k]
