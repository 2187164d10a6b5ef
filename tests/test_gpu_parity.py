"""GPU parity tests: the CUDA path (through the C ABI) against the committed golden vectors and
the CPU oracle on the same seeded inputs.  Tolerances are the north star's: hidden states /
logits rtol 1e-5 atol 1e-6; gradients rtol 1e-4 with atol = 1e-4*max|g_ref| per tensor."""
import numpy as np
import os

import pytest
import torch

from conftest import golden_case_names, golden_model_names, load_golden, params_from_golden
from gpu_helpers import (assert_state_parity, check_out_against_golden, dev, fastgrnn_cuda_from_golden, fastgrnn_from_golden,
                         golden_grad_out, grad_ratio, load_cell_params, state_ratio)
from oracle import fastgrnn_oracle as O

pytestmark = pytest.mark.gpu

CUDA_LAYOUT_CASES = [n for n in golden_case_names() if n not in ("small_quant_a", "small_quant_b", "small_update_sigmoid")]
GRAD_CASES = [n for n in golden_case_names() if n not in ("long_t1000", "bf16_input")]


def test_native_library_is_loaded():
    from kws_b200 import _lib
    _lib.load()
    maps = open("/proc/self/maps").read()
    assert os.path.basename(_lib.LIB_PATH) in maps        # libfastgrnn_b200.so (or the developer build KWS_B200_LIB names)


@pytest.mark.parametrize("name", golden_case_names())
def test_forward_golden_fastgrnn_module(name):
    """rnn.FastGRNN (oracle layout, 3-D hiddenState) vs reference outputs."""
    g = load_golden(name)
    m, d = fastgrnn_from_golden(g)
    x = torch.from_numpy(g["x"]).to(dev())
    hs = torch.from_numpy(g["h0"]).to(dev()).unsqueeze(0) if "h0" in g else None
    with torch.no_grad():
        out = m(x, hs)
    assert out.shape == (x.shape[0], x.shape[1], d["H"])
    assert check_out_against_golden(out, g, d) <= 1.0
    if hs is not None:   # BaseRNN leaves the final state in the caller's hiddenState (rnn.py:621)
        assert state_ratio(hs[0], g["out_last"]) <= 1.0


@pytest.mark.parametrize("name", CUDA_LAYOUT_CASES)
def test_forward_golden_fastgrnn_cuda_module(name):
    """rnn.FastGRNNCUDA (transposed layout, 2-D hiddenState) vs reference outputs."""
    g = load_golden(name)
    m, d = fastgrnn_cuda_from_golden(g)
    x = torch.from_numpy(g["x"]).to(dev())
    h0 = torch.from_numpy(g["h0"]).to(dev()) if "h0" in g else None
    with torch.no_grad():
        out = m(x, h0)
    assert check_out_against_golden(out, g, d) <= 1.0


@pytest.mark.parametrize("name", GRAD_CASES)
@pytest.mark.parametrize("kind", ["FastGRNN", "FastGRNNCUDA"])
def test_backward_golden(name, kind):
    """All gradient slots vs CPU autograd of the reference (golden)."""
    g = load_golden(name)
    if kind == "FastGRNNCUDA" and name not in CUDA_LAYOUT_CASES:
        pytest.skip("CUDA-layout module fixes update_nonlinearity=tanh")
    m, d = (fastgrnn_from_golden if kind == "FastGRNN" else fastgrnn_cuda_from_golden)(g)
    x = torch.from_numpy(g["x"]).to(dev()).requires_grad_(True)
    h0 = torch.from_numpy(g["h0"]).to(dev()).requires_grad_(True) if "h0" in g else None
    if kind == "FastGRNN":
        out = m(x, (h0 * 1.0).unsqueeze(0) if h0 is not None else None)
        owner, tr = m.cell, False
    else:
        out = m(x, h0)
        owner, tr = m, True
    go = golden_grad_out(g, tuple(out.shape)).to(dev())
    out.backward(go)
    for k in [k[2:] for k in g if k.startswith("p_")]:
        got = getattr(owner, k).grad
        ref = torch.from_numpy(g["g_" + k])
        if tr and k in ("W", "U", "W1", "W2", "U1", "U2"):
            ref = ref.t()
        assert got.shape == ref.shape, k
        assert grad_ratio(got, ref) <= 1.0, (k, grad_ratio(got, ref))
    tdim = 1 if d["batch_first"] else 0
    if "g_x" in g:
        assert grad_ratio(x.grad, g["g_x"]) <= 1.0
    else:
        assert grad_ratio(x.grad.select(tdim, d["T"] - 1), g["g_x_last"]) <= 1.0
        assert grad_ratio(x.grad.select(tdim, 0), g["g_x_first"]) <= 1.0
    if "g_h0" in g:
        assert grad_ratio(h0.grad, g["g_h0"]) <= 1.0


@pytest.mark.parametrize("B,T,I,H,wR,uR,gate,bf", [
    (1, 1, 1, 1, None, None, "sigmoid", False),
    (3, 5, 7, 10, None, None, "tanh", True),           # nothing a multiple of 4
    (9, 4, 32, 64, None, None, "sigmoid", False),
    (130, 11, 32, 128, None, None, "sigmoid", True),   # ragged last tile
    (17, 6, 64, 256, None, None, "sigmoid", False),
    (33, 7, 32, 256, 16, 32, "sigmoid", True),         # C4 ranks
    (12, 5, 20, 48, 5, None, "tanh", False),
    (12, 5, 20, 48, None, 7, "relu", True),
    (20, 9, 256, 128, None, None, "sigmoid", False),   # layer 2 of the trainer's default stack
])
def test_forward_backward_vs_oracle_seeded(B, T, I, H, wR, uR, gate, bf):
    from kws_b200 import rnn
    torch.manual_seed(1234 + B * 7 + T)
    p = O.init_params(I, H, wR, uR)
    p.bias_gate.add_(0.2 * torch.randn(1, H)); p.bias_update.add_(0.2 * torch.randn(1, H))
    x = torch.randn(B, T, I) if bf else torch.randn(T, B, I)
    h0 = 0.5 * torch.randn(B, H)
    go = torch.randn(B, T, H) if bf else torch.randn(T, B, H)
    ref = O.unroll(x, p, h0.clone().unsqueeze(0), bf, gate, "tanh")
    gref = O.autograd_grads(x, p, h0, go, bf, gate, "tanh")
    m = rnn.FastGRNN(I, H, gate_nonlinearity=gate, wRank=wR, uRank=uR, batch_first=bf)
    load_cell_params(m.cell, {k: v for k, v in p.tensors().items()}, False)
    m = m.to(dev())
    xg = x.to(dev()).requires_grad_(True)
    h0g = h0.to(dev()).requires_grad_(True)
    out = m(xg, (h0g * 1.0).unsqueeze(0))
    assert_state_parity(out.detach(), x, p, h0, bf, gate)
    out.backward(go.to(dev()))
    for k in p.tensors():
        assert grad_ratio(getattr(m.cell, k).grad, gref[k]) <= 1.0, k
    assert grad_ratio(xg.grad, gref["x"]) <= 1.0
    assert grad_ratio(h0g.grad, gref["h0"]) <= 1.0


def test_extension_shim_unroll_signatures():
    """fastgrnn_cuda.forward_unroll / backward_unroll: reference argument order and return lists
    (cuda/fastgrnn_cuda.cpp:147-232, cu:414, cu:556)."""
    from kws_b200 import fastgrnn_cuda as ext
    g = load_golden("small_lr_both_sigmoid_tm")
    p = params_from_golden(g)
    cu = {k: v.to(dev()) for k, v in O.to_cuda_layout(p).items()}
    e = torch.empty(0)
    x = torch.from_numpy(g["x"]).to(dev()); h0 = torch.from_numpy(g["h0"]).to(dev())
    outs = ext.forward_unroll(x, e, e, cu["bias_gate"], cu["bias_update"], cu["zeta"], cu["nu"], h0, 0,
                              cu["W1"], cu["W2"], cu["U1"], cu["U2"])
    assert len(outs) == 3 and all(o.shape == (x.shape[0], x.shape[1], 24) for o in outs)
    assert state_ratio(outs[0], g["out"]) <= 1.0
    assert float(outs[1].min()) >= 0 and float(outs[1].max()) <= 1 and float(outs[2].abs().max()) <= 1
    go = torch.from_numpy(g["grad_out"]).to(dev())
    grads = ext.backward_unroll(go, x, outs[0], cu["zeta"], cu["nu"], e, e, outs[1], outs[2], h0,
                                cu["W1"], cu["W2"], cu["U1"], cu["U2"], 0)
    assert len(grads) == 12
    names = ["x", "bias_gate", "bias_update", "zeta", "nu", "h0", "W", "U", "W1", "W2", "U1", "U2"]
    for n, t in zip(names, grads):
        if n in ("W", "U"):
            assert t.numel() == 0          # unused rank slots are torch.empty(0) (cu:546-555)
            continue
        ref = torch.from_numpy(g["g_" + n])
        if n in ("W1", "W2", "U1", "U2"):
            ref = ref.t()
        assert grad_ratio(t, ref) <= 1.0, n


def test_extension_shim_single_step_signatures():
    """fastgrnn_cuda.forward / backward and the FastGRNNCUDACell module (rnn.py:454-549)."""
    from kws_b200 import fastgrnn_cuda as ext
    from kws_b200 import rnn
    g = load_golden("single_step")
    p = params_from_golden(g)
    cu = {k: v.to(dev()) for k, v in O.to_cuda_layout(p).items()}
    e = torch.empty(0)
    x = torch.from_numpy(g["x"][0]).to(dev()); h0 = torch.from_numpy(g["h0"]).to(dev())
    new_h, z, c = ext.forward(x, cu["W"], cu["U"], cu["bias_gate"], cu["bias_update"], cu["zeta"], cu["nu"], h0, 0, e, e, e, e)
    assert state_ratio(new_h, g["out"][0]) <= 1.0
    go = torch.from_numpy(g["grad_out"][0]).to(dev())
    grads = ext.backward(go, x, h0, cu["zeta"], cu["nu"], cu["W"], cu["U"], z, c, e, e, e, e, 0)
    assert len(grads) == 12 and all(t.numel() == 0 for t in grads[8:])
    assert grad_ratio(grads[0], g["g_x"][0]) <= 1.0
    assert grad_ratio(grads[6], g["g_W"].T) <= 1.0 and grad_ratio(grads[7], g["g_U"].T) <= 1.0
    assert grad_ratio(grads[5], g["g_h0"]) <= 1.0
    cell = rnn.FastGRNNCUDACell(32, 128)
    load_cell_params(cell, p.tensors(), True)
    xr = x.clone().requires_grad_(True)
    out = cell(xr, h0)
    assert state_ratio(out, g["out"][0]) <= 1.0
    out.backward(go)
    assert grad_ratio(cell.W.grad, g["g_W"].T) <= 1.0 and grad_ratio(xr.grad, g["g_x"][0]) <= 1.0
    # FastGRNNCell single step (oracle layout)
    c2 = rnn.FastGRNNCell(32, 128)
    load_cell_params(c2, p.tensors(), False)
    c2 = c2.to(dev())
    assert state_ratio(c2(x, h0), g["out"][0]) <= 1.0


def test_error_behaviour_matches_reference_wording():
    from kws_b200 import fastgrnn_cuda as ext
    from kws_b200 import rnn
    e = torch.empty(0)
    H, I = 16, 6
    cu = dict(W=torch.randn(H, I, device=dev()), U=torch.randn(H, H, device=dev()),
              bg=torch.ones(1, H, device=dev()), bu=torch.ones(1, H, device=dev()),
              z=torch.ones(1, 1, device=dev()), n=torch.ones(1, 1, device=dev()))
    x = torch.randn(4, 3, I, device=dev()); h0 = torch.zeros(3, H, device=dev())
    with pytest.raises(RuntimeError, match="input must be a CUDA tensor"):
        ext.forward_unroll(x.cpu(), cu["W"], cu["U"], cu["bg"], cu["bu"], cu["z"], cu["n"], h0, 0, e, e, e, e)
    with pytest.raises(RuntimeError, match="w must be contiguous"):
        ext.forward_unroll(x, torch.randn(I, H, device=dev()).t(), cu["U"], cu["bg"], cu["bu"], cu["z"], cu["n"], h0, 0, e, e, e, e)
    with pytest.raises(RuntimeError, match="initial_h must be a CUDA tensor"):
        ext.forward_unroll(x, cu["W"], cu["U"], cu["bg"], cu["bu"], cu["z"], cu["n"], h0.cpu(), 0, e, e, e, e)
    with pytest.raises((RuntimeError, ValueError), match="unknown enum|nonlinearity"):
        ext.forward_unroll(x, cu["W"], cu["U"], cu["bg"], cu["bu"], cu["z"], cu["n"], h0, 7, e, e, e, e)
    m = rnn.FastGRNN(I, H)      # parameters left on the CPU: no silent fallback
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        m(x)


def test_strided_and_permuted_inputs():
    """The trainer feeds audio.permute(2,0,1) views of (B,F,T) tensors (trainClassifier.py:203-204)."""
    from kws_b200 import rnn
    torch.manual_seed(5)
    B, F, T, H = 10, 32, 13, 64
    audio = torch.randn(B, F, T, device=dev())
    x = audio.permute(2, 0, 1)                       # (T,B,F), feature stride != 1
    m = rnn.FastGRNNCUDA(F, H)
    with torch.no_grad():
        a = m(x)
        b = m(x.contiguous())
        # batch-first view of a time-major buffer and vice versa
        c = rnn.FastGRNNCUDA(F, H, batch_first=True)
        c.load_state_dict(m.state_dict())
        d = c(x.contiguous().transpose(0, 1))
    assert torch.equal(a, b)
    assert d.shape == (B, T, H) and torch.equal(d.transpose(0, 1), a)


def test_bf16_input_fp32_state():
    g = load_golden("bf16_input")
    m, d = fastgrnn_from_golden(g)
    x = torch.from_numpy(g["x"]).to(dev())
    assert torch.equal(x.bfloat16().float(), x)        # fixture inputs are exactly bf16-representable
    with torch.no_grad():
        out = m(x.bfloat16())
    assert out.dtype == torch.float32
    assert check_out_against_golden(out, g, d) <= 1.0


@pytest.mark.parametrize("name", golden_model_names())
@pytest.mark.parametrize("kind", ["FastGRNN", "FastGRNNCUDA"])
def test_model_level_logits_and_grads(name, kind):
    """Logits / loss gradients through a stack of layers + Linear + log_softmax, following the call
    pattern of model.py:185-231 (time-major, hidden2keyword(out[-1])), vs the unmodified reference
    model (golden)."""
    from kws_b200 import rnn
    g = load_golden(name)
    hidden = [int(h) for h in g["hidden"]]
    x = torch.from_numpy(g["x"]).to(dev())
    T, B, I = x.shape
    layers = []
    for l, H in enumerate(hidden):
        pl = {k[len("l%d_p_" % l):]: v for k, v in g.items() if k.startswith("l%d_p_" % l)}
        wR = pl["W1"].shape[1] if "W1" in pl else None
        uR = pl["U1"].shape[1] if "U1" in pl else None
        inp = I if l == 0 else hidden[l - 1]
        if kind == "FastGRNN":
            m = rnn.FastGRNN(inp, H, wRank=wR, uRank=uR)
            load_cell_params(m.cell, pl, False)
            m = m.to(dev())
        else:
            m = rnn.FastGRNNCUDA(inp, H, wRank=wR, uRank=uR)
            load_cell_params(m, pl, True)
        layers.append(m)
    head = torch.nn.Linear(hidden[-1], int(g["num_classes"])).to(dev())
    with torch.no_grad():
        head.weight.copy_(torch.from_numpy(g["head_w"])); head.bias.copy_(torch.from_numpy(g["head_b"]))
    h = x
    for m in layers:
        h = m(h, hiddenState=None)                         # model.py:202 / :217
    logp = torch.nn.functional.log_softmax(head(h[-1, :, :]), dim=1)   # model.py:228-230
    assert state_ratio(logp, g["logp"]) <= 1.0, state_ratio(logp, g["logp"])
    loss = torch.nn.functional.nll_loss(logp, torch.from_numpy(g["labels"]).to(dev()))
    assert abs(loss.item() - float(g["loss"])) <= 1e-5 * abs(float(g["loss"])) + 1e-6
    loss.backward()
    assert grad_ratio(head.weight.grad, g["g_head_w"]) <= 1.0
    for l, m in enumerate(layers):
        owner = m.cell if kind == "FastGRNN" else m
        for k in [k[len("l%d_g_" % l):] for k in g if k.startswith("l%d_g_" % l)]:
            ref = torch.from_numpy(g["l%d_g_%s" % (l, k)])
            if kind == "FastGRNNCUDA" and k in ("W", "U", "W1", "W2", "U1", "U2"):
                ref = ref.t()
            assert grad_ratio(getattr(owner, k).grad, ref) <= 1.0, (l, k)


def test_bidirectional_shared_matches_oracle():
    from kws_b200 import rnn
    torch.manual_seed(3)
    T, B, I, H = 6, 5, 8, 16
    p = O.init_params(I, H)
    x = torch.randn(T, B, I)
    fwd = O.unroll(x, p, None, False)
    rev = O.unroll(x.flip(0), p, None, False)          # rnn.py:661-664: processing order, not re-flipped
    m = rnn.FastGRNN(I, H, bidirectional=True)
    load_cell_params(m.cell, p.tensors(), False)
    with torch.no_grad():
        out = m.to(dev())(x.to(dev()))
    assert out.shape == (T, B, 2 * H)
    assert state_ratio(out, torch.cat([fwd, rev], -1)) <= 1.0


def test_empty_batch_and_sequence():
    from kws_b200 import rnn
    m = rnn.FastGRNNCUDA(8, 16)
    with torch.no_grad():
        assert m(torch.zeros(0, 4, 8, device=dev())).shape == (0, 4, 16)
        assert m(torch.zeros(5, 0, 8, device=dev())).shape == (5, 0, 16)


@pytest.mark.parametrize("path", ["generic", "smem"])
@pytest.mark.parametrize("layout", ["IH", "HI"])
@pytest.mark.parametrize("cfg", ["A", "A7", "B", "C"])
def test_kernel_families_agree_with_oracle(path, layout, cfg, tuning):
    """Every kernel family that covers the flagship shape (I=32, H=128, full rank) is checked on
    its own, in both weight layouts, forward and backward, with a ragged batch."""
    from kws_b200 import _lib, engine
    tuning("FGRNN_SMEM_CFG", cfg)
    if path == "generic" and cfg != "A":
        pytest.skip("tile override only affects the shared-memory family")
    torch.manual_seed(77)
    B, T, I, H = 77, 9, 32, 128
    p = O.init_params(I, H)
    p.bias_gate.add_(0.2 * torch.randn(1, H)); p.zeta.add_(0.3); p.nu.add_(0.5)
    x = torch.randn(B, T, I); h0 = 0.5 * torch.randn(B, H); go = torch.randn(B, T, H)
    ref = O.unroll(x, p, h0.clone().unsqueeze(0), True)
    gref = O.autograd_grads(x, p, h0, go, True)
    tens = p.tensors() if layout == "IH" else O.to_cuda_layout(p)
    params = {k: v.to(dev()).contiguous() for k, v in tens.items()}
    force = {"generic": _lib.PATH_GENERIC, "smem": _lib.PATH_SMEM}[path]
    xg, h0g = x.to(dev()), h0.to(dev())
    assert engine.forward_plan(xg, params, h0g, layout=layout, batch_first=True, force_path=force) == path
    out, z_s, c_s, last = engine.forward(xg, params, h0g, layout=layout, batch_first=True,
                                         save_for_backward=True, want_last=True, force_path=force)
    assert state_ratio(out, ref) <= 1.0
    assert torch.equal(last, out[:, -1])
    g = engine.backward(go.to(dev()), xg, out, z_s, c_s, params, h0g, layout=layout, batch_first=True, force_path=force)
    for k in p.tensors():
        r = gref[k].t() if (layout == "HI" and k in ("W", "U")) else gref[k]
        assert grad_ratio(g[k], r) <= 1.0, k
    assert grad_ratio(g["x"], gref["x"]) <= 1.0 and grad_ratio(g["h0"], gref["h0"]) <= 1.0
    # last-state-only mode (no [T,B,H] write at all)
    _, _, _, last2 = engine.forward(xg, params, h0g, layout=layout, batch_first=True, want_states=False, force_path=force)
    assert torch.equal(last2, last)


def test_auto_plan_selects_persistent_kernel_for_flagship_shape():
    from kws_b200 import engine
    p = {k: v.to(dev()) for k, v in O.init_params(32, 128).tensors().items()}
    x = torch.zeros(64, 99, 32, device=dev())
    assert engine.forward_plan(x, p, None, layout="IH", batch_first=True) in ("smem", "tcgen05")
    p2 = {k: v.to(dev()) for k, v in O.init_params(30, 96).tensors().items()}
    assert engine.forward_plan(torch.zeros(4, 5, 30, device=dev()), p2, None, layout="IH", batch_first=True) == "generic"


@pytest.mark.parametrize("B,T,I,wR,uR,layout,bf,gate,xbf16", [
    (150, 12, 32, 16, 32, "IH", True, "sigmoid", False),    # C4 ranks, three CTAs, ragged last one
    (64, 3, 32, 16, 32, "HI", False, "sigmoid", False),     # FastGRNNCUDA layout, time-major
    (70, 6, 16, 8, 16, "IH", True, "tanh", False),          # other ranks, tanh gate (generic activations)
    (40, 9, 64, 16, 32, "HI", True, "sigmoid", True),       # I = 64 (the largest shape that fits shared memory), bf16 input
])
def test_lowrank_ffma_path_vs_oracle_and_generic(B, T, I, wR, uR, layout, bf, gate, xbf16):
    """H = 256 low-rank forward on the persistent FFMA kernel (fgrnn_lr.cu): against the oracle's factored evaluation
    (rnn.py:280-287) and against the generic kernel, with saved gates and the last state."""
    from kws_b200 import _lib, engine
    torch.manual_seed(77 + B + T)
    p = O.init_params(I, 256, wR, uR)
    p.bias_gate.add_(0.2 * torch.randn(1, 256)); p.bias_update.add_(0.2 * torch.randn(1, 256))
    x = torch.randn(B, T, I) if bf else torch.randn(T, B, I)
    if xbf16:
        x = x.bfloat16().float()
    h0 = 0.5 * torch.randn(B, 256)
    ref = O.unroll(x, p, h0.clone().unsqueeze(0), bf, gate, "tanh")
    tens = p.tensors() if layout == "IH" else O.to_cuda_layout(p)
    params = {k: v.to(dev()).contiguous() for k, v in tens.items()}
    xg = x.to(dev()).bfloat16() if xbf16 else x.to(dev())
    h0g = h0.to(dev())
    kw = dict(layout=layout, batch_first=bf, gate_nl=gate, update_nl="tanh")
    assert engine.forward_plan(xg, params, h0g, **kw) == "lowrank"
    out, z_s, c_s, last = engine.forward(xg, params, h0g, want_last=True, save_for_backward=True, **kw)
    gen, z_g, c_g, _ = engine.forward(xg, params, h0g, save_for_backward=True, force_path=_lib.PATH_GENERIC, **kw)
    torch.cuda.synchronize()
    assert state_ratio(out, ref) <= 1.0, state_ratio(out, ref)
    assert state_ratio(out, gen.cpu()) <= 1.0
    assert state_ratio(z_s, z_g.cpu()) <= 1.0 and state_ratio(c_s, c_g.cpu()) <= 1.0
    assert torch.equal(last, out[:, -1] if bf else out[-1])
    # shapes the kernel does not cover keep the generic path
    big = O.init_params(64, 256, 32, 64)
    assert engine.forward_plan(torch.randn(4, 2, 64, device=dev()), {k: v.to(dev()) for k, v in big.tensors().items()},
                               None, layout="IH", batch_first=True) == "generic"


def test_lowrank_ffma_path_extreme_biases():
    """The low-rank kernel's fused sigmoid/tanh (one EX2 + one RCP per element) is taken per thread only when its four
    units have |b_g - b_u| <= 8; far-apart biases and saturated gates must still meet the tolerance."""
    from kws_b200 import engine
    torch.manual_seed(5)
    p = O.init_params(32, 256, 16, 32)
    p.bias_gate[0, :8] += 12.0           # distance > 8: these unit groups use the two-EX2 form
    p.bias_update[0, 8:16] -= 15.0
    p.bias_gate[0, 16:24] -= 30.0        # gate saturated at 0: e_g clamps at 2^30
    p.bias_gate[0, 24:32] += 30.0        # gate saturated at 1
    p.bias_update[0, 32:40] += 7.5       # inside the fused range, large ratio
    x = 4.0 * torch.randn(37, 11, 32)
    ref = O.unroll(x, p, None, True)
    params = {k: v.to(dev()).contiguous() for k, v in p.tensors().items()}
    assert engine.forward_plan(x.to(dev()), params, None, layout="IH", batch_first=True) == "lowrank"
    out = engine.forward(x.to(dev()), params, None, layout="IH", batch_first=True)[0]
    torch.cuda.synchronize()
    assert torch.isfinite(out).all()
    assert state_ratio(out, ref) <= 1.0, state_ratio(out, ref)
