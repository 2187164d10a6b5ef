"""GPU: the tcgen05/TMEM forward recurrence (fp16 hi+lo split, 3 MMAs per product, fp32
accumulation in TMEM) against the CPU oracle with the north-star tolerances."""
import pytest
import torch

from gpu_helpers import dev, state_ratio
from oracle import fastgrnn_oracle as O

pytestmark = pytest.mark.gpu


def _run(B, T, I, layout, h0_given, seed, x_bf16=False, wscale=1.0):
    from kws_b200 import _lib, engine
    torch.manual_seed(seed)
    p = O.init_params(I, 128)
    p.bias_gate.add_(0.2 * torch.randn(1, 128)); p.bias_update.add_(0.2 * torch.randn(1, 128))
    p.W.mul_(wscale); p.U.mul_(wscale)
    x = torch.randn(B, T, I)
    if x_bf16:
        x = x.bfloat16().float()
    h0 = 0.5 * torch.randn(B, 128) if h0_given else None
    ref = O.unroll(x, p, None if h0 is None else h0.clone().unsqueeze(0), True)
    tens = p.tensors() if layout == "IH" else O.to_cuda_layout(p)
    params = {k: v.to(dev()).contiguous() for k, v in tens.items()}
    xg = x.to(dev())
    if x_bf16:
        xg = xg.bfloat16()
    h0g = None if h0 is None else h0.to(dev())
    assert engine.forward_plan(xg, params, h0g, layout=layout, batch_first=True, force_path=_lib.PATH_TCGEN05) == "tcgen05"
    out, _, _, last = engine.forward(xg, params, h0g, layout=layout, batch_first=True, want_last=True,
                                     force_path=_lib.PATH_TCGEN05)
    torch.cuda.synchronize()
    return out, last, ref


@pytest.mark.parametrize("B,T,I,layout,h0", [
    (128, 1, 32, "IH", False),
    (128, 3, 32, "IH", True),
    (77, 9, 32, "HI", True),          # ragged single tile
    (300, 17, 16, "IH", False),       # three tiles, I = 16
    (64, 99, 32, "IH", False),        # BASELINE config-1 shape
])
def test_tcgen05_forward_vs_oracle(B, T, I, layout, h0):
    out, last, ref = _run(B, T, I, layout, h0, seed=11 + B + T)
    r = state_ratio(out, ref)
    assert r <= 1.0, r
    assert torch.equal(last, out[:, -1])


def test_tcgen05_bf16_input_and_weight_scales():
    out, _, ref = _run(96, 20, 32, "IH", True, seed=5, x_bf16=True)
    assert state_ratio(out, ref) <= 1.0
    for ws in (0.2, 3.0):             # power-of-two operand scaling adapts to the weight magnitude
        out, _, ref = _run(40, 6, 32, "HI", False, seed=6, wscale=ws)
        assert state_ratio(out, ref) <= 1.0, ws
