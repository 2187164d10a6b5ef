"""CPU: pin the oracle -- against the committed golden fixtures (minted from the unmodified
reference, oracle/make_golden.py) and, where /root/reference exists, bit-for-bit against the
reference code itself."""
import numpy as np
import pytest
import torch

from conftest import GRAD_RTOL, golden_case_names, load_golden, params_from_golden
from oracle import fastgrnn_oracle as O
from oracle import ref_shim


def _replay(g, double=False):
    p = params_from_golden(g)
    x = torch.from_numpy(g["x"].copy())
    bf = bool(g["batch_first"])
    h0 = torch.from_numpy(g["h0"].copy()) if "h0" in g else None
    gate, update = str(g["gate"]), str(g["update"])
    return p, x, bf, h0, gate, update


@pytest.mark.parametrize("name", golden_case_names())
def test_oracle_forward_matches_golden(name):
    g = load_golden(name)
    p, x, bf, h0, gate, update = _replay(g)
    out = O.unroll(x, p, None if h0 is None else h0.clone().unsqueeze(0), bf, gate, update)
    tdim = 1 if bf else 0
    # same machine/torch build => bit-identical; other builds may differ in the last ulp
    tol = dict(rtol=2e-6, atol=2e-7)
    if "out" in g:
        np.testing.assert_allclose(out.numpy(), g["out"], **tol)
    else:
        keep = torch.from_numpy(g["keep_t"])
        np.testing.assert_allclose(out.index_select(tdim, keep).numpy(), g["out_keep"], **tol)
    np.testing.assert_allclose(out.select(tdim, x.shape[tdim] - 1).numpy(), g["out_last"], **tol)
    assert abs(out.double().sum().item() - float(g["out_sum"])) <= 1e-6 * max(1.0, abs(float(g["out_sum"])))
    assert abs((out.double() ** 2).sum().item() - float(g["out_sumsq"])) <= 1e-6 * float(g["out_sumsq"])


@pytest.mark.parametrize("name", [n for n in golden_case_names() if n.startswith("small") or n in ("odd_b37_h64", "single_step", "h256_full")])
def test_oracle_grads_match_golden(name):
    g = load_golden(name)
    p, x, bf, h0, gate, update = _replay(g)
    go = torch.from_numpy(g["grad_out"].copy())
    grads = O.autograd_grads(x, p, h0, go, bf, gate, update)
    for k in p.tensors():
        assert O.grad_tolerance_ratio(grads[k], torch.from_numpy(g["g_" + k]), 1e-5) <= 1.0, k
    assert O.grad_tolerance_ratio(grads["x"], torch.from_numpy(g["g_x"]), 1e-5) <= 1.0
    if "g_h0" in g:
        assert O.grad_tolerance_ratio(grads["h0"], torch.from_numpy(g["g_h0"]), 1e-5) <= 1.0


@pytest.mark.parametrize("name", ["small_full_sigmoid_tm", "small_full_tanh_bf", "small_lr_both_sigmoid_tm",
                                  "small_lr_u_tanh_bf", "small_quant_a", "small_quant_b", "small_update_sigmoid"])
def test_closed_form_bptt_matches_autograd_fp64(name):
    """The kernel-shaped closed form (cu:109-118,537-556 with the correct tanh-gate
    derivative) equals autograd of the restated reference in fp64."""
    g = load_golden(name)
    p, x, bf, h0, gate, update = _replay(g)
    p64 = p.map(lambda v: v.double())
    x64 = x.double()
    B = x.shape[0] if bf else x.shape[1]
    h064 = h0.double() if h0 is not None else torch.zeros(B, p.hidden_size, dtype=torch.float64)
    go = torch.from_numpy(g["grad_out"].copy()).double()
    x_tm = x64.transpose(0, 1) if bf else x64
    go_tm = go.transpose(0, 1) if bf else go
    cf = O.bptt_closed_form(x_tm, p64, h064, go_tm, gate, update)
    # fp64 autograd through the functional unroll (D12: the buffered unroll is fp32-only)
    xq = x64.clone().requires_grad_(True)
    q = p64.map(lambda v: v.clone().requires_grad_(True))
    hq = h064.clone().requires_grad_(True)
    out = O.unroll_functional(xq, q, hq, bf, gate, update)
    out.backward(go)
    for k, v in q.tensors().items():
        np.testing.assert_allclose(cf[k].numpy(), v.grad.numpy(), rtol=1e-9, atol=1e-11, err_msg=k)
    gx = xq.grad.transpose(0, 1) if bf else xq.grad
    np.testing.assert_allclose(cf["x"].numpy(), gx.numpy(), rtol=1e-9, atol=1e-11)
    np.testing.assert_allclose(cf["h0"].numpy(), hq.grad.numpy(), rtol=1e-9, atol=1e-11)


def test_reference_tanh_gate_unrolled_backward_is_wrong():
    """Documents SURVEY D6: using z(1-z) for a tanh gate (cu:519-521) is off by O(1)."""
    g = load_golden("small_full_tanh_bf")
    p, x, bf, h0, gate, update = _replay(g)
    assert gate == "tanh"
    z = torch.tensor([0.5])
    assert abs(float(O._dgate(z, z, "tanh")) - float(O._dgate(z, z, "sigmoid"))) > 0.4


needs_ref = pytest.mark.skipif(not ref_shim.available(), reason="/root/reference not present (GPU box)")


@needs_ref
@pytest.mark.parametrize("wR,uR,gate,update,bf", [
    (None, None, "sigmoid", "tanh", True), (None, None, "tanh", "tanh", False),
    (3, 4, "sigmoid", "tanh", False), (3, None, "sigmoid", "tanh", True), (None, 4, "tanh", "tanh", True),
    (None, None, "quantSigm", "quantTanh", False), (None, None, "quantSigm4", "sigmoid", True)])
def test_oracle_bit_identical_to_reference(wR, uR, gate, update, bf):
    torch.manual_seed(11)
    m = ref_shim.make_fastgrnn(6, 16, gate_nonlinearity=gate, update_nonlinearity=update, wRank=wR, uRank=uR, batch_first=bf)
    torch.manual_seed(11)
    p2 = O.init_params(6, 16, wR, uR)
    p = ref_shim.params_of(m)
    for k, v in p.tensors().items():
        assert torch.equal(v.detach(), p2.tensors()[k]), "init order differs for " + k
    x = torch.randn(5, 7, 6) if bf else torch.randn(7, 5, 6)
    h0 = torch.randn(1, 5, 16)
    go = torch.randn(5, 7, 16) if bf else torch.randn(7, 5, 16)
    xr = x.clone().requires_grad_(True)
    h0l = h0.clone().requires_grad_(True)
    ref = m(xr, h0l * 1.0)
    ref.backward(go)
    mine = O.unroll(x, p.map(lambda v: v.detach()), h0.clone(), bf, gate, update)
    assert torch.equal(ref.detach(), mine)
    grads = O.autograd_grads(x, p.map(lambda v: v.detach()), h0[0], go, bf, gate, update)
    for k, v in p.tensors().items():
        assert torch.equal(v.grad, grads[k]), k
    assert torch.equal(xr.grad, grads["x"]) and torch.equal(h0l.grad[0], grads["h0"])


@needs_ref
def test_oracle_bit_identical_to_reference_c1_shape():
    """BASELINE config 1: B=64, T=99, I=32, H=128, batch-first, reference init."""
    torch.manual_seed(0)
    m = ref_shim.make_fastgrnn(32, 128, batch_first=True)
    x = torch.randn(64, 99, 32)
    with torch.no_grad():
        ref = m(x)
        mine = O.unroll(x, ref_shim.params_of(m).map(lambda v: v.detach()), None, True)
    assert torch.equal(ref, mine)
    assert 0.3 < float(ref.abs().mean()) < 0.5


@needs_ref
def test_reference_defects_still_present():
    """The shims exist because of D1/D2; if the reference ever changes, revisit them."""
    import importlib, sys
    rnn, _ = ref_shim.load()
    src = open(ref_shim.REF_ROOT + "/rnn.py").read()
    assert "training=training" in src and "def forward(self, input, state):" in src      # D1
    assert "device = self.W.device" in src                                                 # D2
    assert "#import fastgrnn_cuda" in src                                                  # D4
