"""GPU: the low-rank recurrence on the tensor cores (fgrnn_tc_lr.cu: two chained tcgen05 stages, factored order of
rnn.py:280-287 kept) against the CPU oracle with the north-star tolerances, and against the FFMA low-rank kernel it
replaces for inference (FGRNN_TC_LR=0 selects that one)."""
import pytest
import torch

from gpu_helpers import dev, state_ratio
from oracle import fastgrnn_oracle as O

pytestmark = pytest.mark.gpu


def _case(B, T, I, wR, uR, layout, bf, h0_given, xbf16, seed):
    torch.manual_seed(seed)
    p = O.init_params(I, 256, wR, uR)
    p.bias_gate.add_(0.2 * torch.randn(1, 256)); p.bias_update.add_(0.2 * torch.randn(1, 256))
    x = torch.randn(B, T, I) if bf else torch.randn(T, B, I)
    if xbf16:
        x = x.bfloat16().float()
    h0 = 0.5 * torch.randn(B, 256) if h0_given else None
    ref = O.unroll(x, p, None if h0 is None else h0.clone().unsqueeze(0), bf)
    tens = p.tensors() if layout == "IH" else O.to_cuda_layout(p)
    params = {k: v.to(dev()).contiguous() for k, v in tens.items()}
    xg = x.to(dev()).bfloat16() if xbf16 else x.to(dev())
    return p, params, xg, None if h0 is None else h0.to(dev()), ref


@pytest.mark.parametrize("B,T,I,wR,uR,layout,bf,h0,xbf16", [
    (64, 1, 32, 16, 32, "IH", True, False, False),      # one step, one CTA
    (64, 2, 32, 16, 32, "IH", True, True, False),
    (150, 12, 32, 16, 32, "IH", True, True, False),     # C4 ranks, three CTAs, ragged last one
    (64, 7, 32, 16, 32, "HI", False, True, False),      # FastGRNNCUDA layout, time-major
    (70, 6, 16, 8, 16, "IH", True, False, False),       # smaller ranks (zero padded), I = 16: one x k-step
    (40, 9, 24, 12, 20, "HI", True, True, False),       # ranks / features that are not multiples of 16
    (33, 5, 32, 16, 32, "IH", True, True, True),        # bf16 input
    (300, 99, 32, 16, 32, "IH", True, False, False),    # BASELINE config 4 sequence length
])
def test_tc_lowrank_forward_vs_oracle_and_ffma(B, T, I, wR, uR, layout, bf, h0, xbf16, tuning):
    from kws_b200 import engine
    p, params, xg, h0g, ref = _case(B, T, I, wR, uR, layout, bf, h0, xbf16, seed=31 + B + T)
    kw = dict(layout=layout, batch_first=bf)
    assert engine.forward_plan(xg, params, h0g, **kw) == "lowrank"
    out, _, _, last = engine.forward(xg, params, h0g, want_last=True, **kw)
    again = engine.forward(xg, params, h0g, **kw)[0]
    torch.cuda.synchronize()
    tuning("FGRNN_TC_LR", "0")
    ffma = engine.forward(xg, params, h0g, **kw)[0]
    torch.cuda.synchronize()
    r = state_ratio(out, ref)
    assert torch.isfinite(out).all()
    assert r <= 1.0, (r, state_ratio(ffma, ref))
    assert state_ratio(out, ffma.cpu()) <= 1.0
    assert not torch.equal(out, ffma) or T * B == 0            # two different kernels really ran
    assert torch.equal(out, again)                            # deterministic
    assert torch.equal(last, out[:, -1] if bf else out[-1])


@pytest.mark.parametrize("vr", ["14", "12"])
def test_tc_lowrank_valid_rows_per_thread_is_bit_identical(tuning, vr):
    """56- and 48-row CTAs (FGRNN_TC_VR = 14 / 12 valid rows per epilogue thread and sub-tile, the rest of the 2 x 32 MMA
    columns padding) against the 64-row CTAs: same bits, ragged and multi-CTA batches, both layouts, bf16 input, h0."""
    from kws_b200 import engine
    for (B, T, I, wR, uR, layout, bf, h0, xbf16) in [(64, 3, 32, 16, 32, "IH", True, True, False), (150, 12, 32, 16, 32, "IH", True, True, False),
                                                      (57, 7, 24, 12, 20, "HI", False, True, False), (33, 5, 32, 16, 32, "IH", True, False, True),
                                                      (2000, 4, 32, 16, 32, "IH", True, False, False)]:
        p, params, xg, h0g, ref = _case(B, T, I, wR, uR, layout, bf, h0, xbf16, seed=7 + B)
        kw = dict(layout=layout, batch_first=bf, want_last=True)
        tuning("FGRNN_TC_VR", "16")
        a = engine.forward(xg, params, h0g, **kw)
        tuning("FGRNN_TC_VR", vr)
        b = engine.forward(xg, params, h0g, **kw)
        torch.cuda.synchronize()
        assert torch.isfinite(a[0]).all()
        assert torch.equal(a[0], b[0]) and torch.equal(a[3], b[3]), (vr, B)
        assert state_ratio(b[0], ref) <= 1.0


def test_tc_lowrank_last_state_only_and_chunked_carry():
    """want_states=False writes no [B,T,H] at all; two chunks with state carry == one long call, bit for bit."""
    from kws_b200 import engine
    p, params, xg, h0g, ref = _case(96, 20, 32, 16, 32, "IH", True, True, False, seed=5)
    kw = dict(layout="IH", batch_first=True)
    full, _, _, last = engine.forward(xg, params, h0g, want_last=True, **kw)
    _, _, _, only = engine.forward(xg, params, h0g, want_states=False, **kw)
    a, _, _, ha = engine.forward(xg[:, :9].contiguous(), params, h0g, want_last=True, **kw)
    b = engine.forward(xg[:, 9:].contiguous(), params, ha, **kw)[0]
    torch.cuda.synchronize()
    assert torch.equal(only, last)
    assert torch.equal(torch.cat([a, b], 1), full)
    assert state_ratio(full, ref) <= 1.0


def test_tc_lowrank_extreme_biases_and_inputs():
    """Far-apart biases (two-EX2 form for the whole CTA), saturated gates, large inputs."""
    from kws_b200 import engine
    torch.manual_seed(5)
    p = O.init_params(32, 256, 16, 32)
    p.bias_gate[0, :8] += 12.0
    p.bias_update[0, 8:16] -= 15.0
    p.bias_gate[0, 16:24] -= 30.0
    p.bias_gate[0, 24:32] += 30.0
    x = 4.0 * torch.randn(37, 11, 32)
    ref = O.unroll(x, p, None, True)
    params = {k: v.to(dev()).contiguous() for k, v in p.tensors().items()}
    out = engine.forward(x.to(dev()), params, None, layout="IH", batch_first=True)[0]
    torch.cuda.synchronize()
    assert torch.isfinite(out).all()
    assert state_ratio(out, ref) <= 1.0, state_ratio(out, ref)


def test_tc_lowrank_nan_input_propagates():
    """A NaN feature poisons its own row only (as in the reference), not its tile neighbours."""
    from kws_b200 import engine
    p, params, xg, _, ref = _case(64, 4, 32, 16, 32, "IH", True, False, False, seed=9)
    xg = xg.clone(); xg[5, 1, 3] = float("nan")
    out = engine.forward(xg, params, None, layout="IH", batch_first=True)[0]
    torch.cuda.synchronize()
    assert torch.isnan(out[5, 1:]).all() and torch.isfinite(out[5, 0]).all()
    keep = [r for r in range(64) if r != 5]
    assert torch.isfinite(out[keep]).all()
    assert state_ratio(out[keep], ref[keep]) <= 1.0


def test_training_forward_keeps_the_ffma_kernel():
    """save_for_backward needs z_t / c_t, which the tensor-core kernel does not store: same results as before."""
    from kws_b200 import _lib, engine
    p, params, xg, h0g, ref = _case(64, 5, 32, 16, 32, "IH", True, True, False, seed=2)
    out, z_s, c_s, _ = engine.forward(xg, params, h0g, layout="IH", batch_first=True, save_for_backward=True)
    gen, z_g, c_g, _ = engine.forward(xg, params, h0g, layout="IH", batch_first=True, save_for_backward=True, force_path=_lib.PATH_GENERIC)
    torch.cuda.synchronize()
    assert state_ratio(out, ref) <= 1.0 and state_ratio(z_s, z_g.cpu()) <= 1.0 and state_ratio(c_s, c_g.cpu()) <= 1.0
