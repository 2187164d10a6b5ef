"""GPU: the tcgen05/TMEM forward recurrence (weights in tensor memory, fp16 hi+lo split, three
accumulators per sub-tile, x by TMA) against the CPU oracle with the north-star tolerances."""
import pytest
import torch

from gpu_helpers import dev, state_ratio
from oracle import fastgrnn_oracle as O

pytestmark = pytest.mark.gpu


def _run(B, T, I, layout, h0_given, seed, x_bf16=False, wscale=1.0, batch_first=True, save=False):
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
    if not batch_first:
        xg = xg.transpose(0, 1).contiguous()
    if x_bf16:
        xg = xg.bfloat16()
    h0g = None if h0 is None else h0.to(dev())
    assert engine.forward_plan(xg, params, h0g, layout=layout, batch_first=batch_first, force_path=_lib.PATH_TCGEN05) == "tcgen05"
    out, z_s, c_s, last = engine.forward(xg, params, h0g, layout=layout, batch_first=batch_first, want_last=True,
                                         save_for_backward=save, force_path=_lib.PATH_TCGEN05)
    torch.cuda.synchronize()
    if not batch_first:
        out = out.transpose(0, 1)
    return out, last, ref, z_s, c_s


@pytest.mark.parametrize("B,T,I,layout,h0", [
    (64, 1, 32, "IH", False),
    (128, 3, 32, "IH", True),
    (77, 9, 32, "HI", True),          # ragged single CTA, second sub-tile partly empty
    (300, 17, 16, "IH", False),       # five CTAs, I = 16
    (40, 5, 64, "HI", True),          # I = 64 (MFCC + deltas, preprocessing.py)
    (33, 4, 24, "IH", True),          # I = 24: K padded to 32 with zeros
    (64, 99, 32, "IH", False),        # BASELINE config-1 shape
])
def test_tcgen05_forward_vs_oracle(B, T, I, layout, h0):
    out, last, ref, _, _ = _run(B, T, I, layout, h0, seed=11 + B + T)
    r = state_ratio(out, ref)
    assert r <= 1.0, r
    assert torch.equal(last, out[:, -1])


@pytest.mark.parametrize("ns,nt,acc2", [("16", "2", None), ("16", "4", None), ("32", "2", None), ("32", "2", "1"), ("16", "2", "1")])
def test_tcgen05_forward_every_tile_configuration(tuning, ns, nt, acc2):
    """The launcher picks 2x16-row sub-tiles for one-wave batches and 4x16 beyond; FGRNN_TC_NS / FGRNN_TC_NT pin a
    configuration.  All three (incl. the 2x32 variant) must meet the tolerance on ragged, multi-CTA, saved-gate runs."""
    tuning("FGRNN_TC_NS", ns)
    tuning("FGRNN_TC_NT", nt)
    if acc2 is not None:                               # two accumulators per sub-tile, lo products first (opt-in)
        tuning("FGRNN_TC_ACC2", acc2)
    for (B, T, I, layout, h0, save) in [(77, 9, 32, "HI", True, False), (300, 17, 16, "IH", False, True), (33, 4, 24, "IH", True, False)]:
        out, last, ref, z_s, c_s = _run(B, T, I, layout, h0, seed=21 + B, save=save)
        assert state_ratio(out, ref) <= 1.0, (ns, nt, B)
        assert torch.equal(last, out[:, -1])
        if save:
            assert float(z_s.min()) >= 0.0 and float(z_s.max()) <= 1.0 and float(c_s.abs().max()) <= 1.0


def test_tcgen05_alternating_epilogue_is_bit_identical_to_the_split_one(tuning):
    """2 x 32-row sub-tiles: sixteen epilogue warps serving both sub-tiles in turn (ALT, the default) compute the very
    bits of the kernel with eight warps per sub-tile (FGRNN_TC_ALT=0) -- states, saved gates, last state; ragged and
    multi-CTA batches, with and without h0."""
    from kws_b200 import _lib, engine
    tuning("FGRNN_TC_NS", "32")
    tuning("FGRNN_TC_NT", "2")
    for (B, T, save, h0_given) in [(64, 5, False, False), (77, 9, True, True), (300, 17, True, False), (8192, 6, False, True)]:
        torch.manual_seed(B)
        p = O.init_params(32, 128)
        params = {k: v.to(dev()).contiguous() for k, v in p.tensors().items()}
        x = torch.randn(B, T, 32, device=dev())
        h0 = 0.5 * torch.randn(B, 128, device=dev()) if h0_given else None
        kw = dict(layout="IH", batch_first=True, want_last=True, save_for_backward=save, force_path=_lib.PATH_TCGEN05)
        tuning("FGRNN_TC_ALT", "1")
        a = engine.forward(x, params, h0, **kw)
        tuning("FGRNN_TC_ALT", "0")
        b = engine.forward(x, params, h0, **kw)
        torch.cuda.synchronize()
        assert torch.equal(a[0], b[0]) and torch.equal(a[3], b[3]), B
        if save:
            assert torch.equal(a[1], b[1]) and torch.equal(a[2], b[2]), B
        if B <= 300:
            ref = O.unroll(x.cpu(), p, None if h0 is None else h0.cpu().unsqueeze(0), True)
            assert state_ratio(a[0], ref) <= 1.0


@pytest.mark.parametrize("ns,vr", [("32", "14")])
def test_tcgen05_valid_rows_per_thread_is_bit_identical(tuning, ns, vr):
    """2 x 32-row sub-tiles with 14 VALID rows per epilogue thread (56-row CTAs, FGRNN_TC_VR=14, opt-in) compute the very bits of the 64-row CTAs: a batch row is an independent column of the MMA.
    Ragged and multi-CTA batches, saved gates, h0, both layouts, time-major output, bf16 input."""
    from kws_b200 import _lib, engine
    tuning("FGRNN_TC_NS", ns)
    tuning("FGRNN_TC_NT", "2")
    cases = [(64, 5, 32, False, False, True, False), (77, 9, 32, True, True, True, False), (300, 17, 16, True, False, False, False),
             (57, 4, 24, False, True, True, True), (8192, 6, 32, False, True, True, False), (4099, 3, 32, True, False, True, False)]
    for (B, T, I, save, h0_given, bf, xbf16) in cases:
        torch.manual_seed(B + T)
        p = O.init_params(I, 128)
        params = {k: v.to(dev()).contiguous() for k, v in p.tensors().items()}
        x = torch.randn(B, T, I, device=dev()) if bf else torch.randn(T, B, I, device=dev())
        if xbf16:
            x = x.bfloat16()
        h0 = 0.5 * torch.randn(B, 128, device=dev()) if h0_given else None
        kw = dict(layout="IH", batch_first=bf, want_last=True, save_for_backward=save, force_path=_lib.PATH_TCGEN05)
        tuning("FGRNN_TC_VR", "16" if ns == "32" else "8")
        a = engine.forward(x, params, h0, **kw)
        tuning("FGRNN_TC_VR", vr)
        b = engine.forward(x, params, h0, **kw)
        torch.cuda.synchronize()
        assert torch.isfinite(a[0]).all()
        assert torch.equal(a[0], b[0]) and torch.equal(a[3], b[3]), (vr, B)
        if save:
            assert torch.equal(a[1], b[1]) and torch.equal(a[2], b[2]), (vr, B)
        if B <= 300:
            ref = O.unroll(x.float().cpu(), p, None if h0 is None else h0.cpu().unsqueeze(0), bf)
            assert state_ratio(b[0], ref) <= 1.0


def test_tcgen05_time_major_and_saved_gates():
    out, last, ref, z_s, c_s = _run(70, 6, 32, "HI", True, seed=3, batch_first=False, save=True)
    assert state_ratio(out, ref) <= 1.0
    # z_s, c_s [T,B,H] reproduce the state update h_t = z h_{t-1} + (sz (1 - z) + sn) c   (rnn.py:294-295)
    assert z_s.shape == (6, 70, 128) and c_s.shape == (6, 70, 128)
    assert float(z_s.min()) >= 0.0 and float(z_s.max()) <= 1.0 and float(c_s.abs().max()) <= 1.0


def test_tcgen05_bf16_input_and_weight_scales():
    out, _, ref, _, _ = _run(96, 20, 32, "IH", True, seed=5, x_bf16=True)
    assert state_ratio(out, ref) <= 1.0
    for ws in (0.05, 0.2, 2.0):       # power-of-two operand scaling adapts to the weight magnitude
        out, _, ref, _, _ = _run(40, 6, 32, "HI", False, seed=6, wscale=ws)
        assert state_ratio(out, ref) <= 1.0, ws


def test_tcgen05_matches_ffma_path_on_a_full_wave():
    """8192 rows (128 CTAs): the two kernel families agree to well inside the tolerance."""
    from kws_b200 import _lib, engine
    torch.manual_seed(0)
    p = O.init_params(32, 128)
    params = {k: v.to(dev()).contiguous() for k, v in p.tensors().items()}
    x = torch.randn(8192, 12, 32, device=dev())
    a = engine.forward(x, params, None, layout="IH", batch_first=True, force_path=_lib.PATH_TCGEN05)[0]
    b = engine.forward(x, params, None, layout="IH", batch_first=True, force_path=_lib.PATH_SMEM)[0]
    torch.cuda.synchronize()
    assert state_ratio(a, b.cpu()) <= 1.0


@pytest.mark.parametrize("B,T,I,layout,h0_given,bf", [
    (64, 5, 32, "IH", True, True),
    (77, 9, 32, "HI", True, True),        # ragged chunk
    (130, 7, 16, "IH", False, False),     # three row blocks per step, time-major, I = 16
    (33, 3, 64, "HI", True, True),        # I = 64
    (2048, 4, 32, "IH", False, True),     # more chunks than one wave of CTAs would take alone: 128 chunks
])
def test_tcgen05_backward_against_autograd(B, T, I, layout, h0_given, bf):
    """Default (auto) backward: reverse recurrence + dW/dU contractions on the tensor cores (bf16 hi/lo split,
    per-CTA partials, fixed-order reduce) against CPU autograd of the oracle, all 12 gradient slots."""
    from kws_b200 import engine
    from gpu_helpers import grad_ratio
    torch.manual_seed(B + T)
    p = O.init_params(I, 128)
    p.bias_gate.add_(0.2 * torch.randn(1, 128)); p.zeta.add_(0.3); p.nu.add_(0.5)
    x = torch.randn(B, T, I)
    h0 = 0.5 * torch.randn(B, 128) if h0_given else None
    go = torch.randn(B, T, 128) / B                       # mean-loss sized gradients (tiny for large B)
    gref = O.autograd_grads(x, p, h0 if h0 is not None else torch.zeros(B, 128), go, True)
    tens = p.tensors() if layout == "IH" else O.to_cuda_layout(p)
    params = {k: v.to(dev()).contiguous() for k, v in tens.items()}
    xg, gog = x.to(dev()), go.to(dev())
    if not bf:
        xg, gog = xg.transpose(0, 1).contiguous(), gog.transpose(0, 1).contiguous()
    h0g = None if h0 is None else h0.to(dev())
    out, z_s, c_s, _ = engine.forward(xg, params, h0g, layout=layout, batch_first=bf, save_for_backward=True)
    g = engine.backward(gog, xg, out, z_s, c_s, params, h0g, layout=layout, batch_first=bf)
    torch.cuda.synchronize()
    for k in p.tensors():
        r = gref[k].t() if (layout == "HI" and k in ("W", "U")) else gref[k]
        assert grad_ratio(g[k], r) <= 1.0, (k, grad_ratio(g[k], r))
    gx = g["x"] if bf else g["x"].transpose(0, 1)
    assert grad_ratio(gx, gref["x"]) <= 1.0
    if h0_given:
        assert grad_ratio(g["h0"], gref["h0"]) <= 1.0
    # run-to-run identical (no atomics anywhere)
    g2 = engine.backward(gog, xg, out, z_s, c_s, params, h0g, layout=layout, batch_first=bf)
    torch.cuda.synchronize()
    assert all(torch.equal(g[k], g2[k]) for k in ("W", "U", "bias_gate", "zeta"))


@pytest.mark.parametrize("ns", ["16", "32"])
def test_tcgen05_backward_both_sub_tile_widths(tuning, ns):
    """The reverse recurrence takes 16-row sub-tiles while the batch fits one wave of 32-row CTAs and 32-row ones
    beyond; FGRNN_TC_BR_NS pins the width (the per-CTA partial rows follow it)."""
    tuning("FGRNN_TC_BR_NS", ns)
    test_tcgen05_backward_against_autograd(77, 9, 32, "HI", True, True)
    test_tcgen05_backward_against_autograd(130, 7, 16, "IH", False, False)


def test_training_step_replays_from_a_cuda_graph():
    """The C-ABI calls are capture-safe (no allocation, no host sync, tensor maps baked into the kernel
    parameters): a captured forward + BPTT + SGD step updates the weights exactly like the eager step."""
    from kws_b200 import graphs, rnn, sharding

    def build():
        torch.manual_seed(3)
        layer = rnn.FastGRNN(32, 128, batch_first=False).to(dev())
        head = torch.nn.Linear(128, 13).to(dev())
        plist = list(layer.cell.parameters()) + list(head.parameters())
        bucket = sharding.GradBucket(plist)
        opt = torch.optim.SGD(plist, lr=1e-2)
        return layer, head, plist, bucket, opt

    torch.manual_seed(4)
    x = torch.randn(9, 96, 32, device=dev())
    labels = torch.randint(0, 13, (96,), device=dev())

    def make_step(layer, head, bucket, opt):
        def step():
            bucket.zero()
            hs = layer(x)
            loss = torch.nn.functional.nll_loss(torch.nn.functional.log_softmax(head(hs[-1]), dim=1), labels)
            loss.backward()
            opt.step()
        return step

    layer_e, head_e, plist_e, bucket_e, opt_e = build()
    eager = make_step(layer_e, head_e, bucket_e, opt_e)
    for _ in range(5):                       # 3 warm-up calls + capture do not run the graph; see below
        eager()
    layer_g, head_g, plist_g, bucket_g, opt_g = build()
    cap = graphs.CapturedStep(make_step(layer_g, head_g, bucket_g, opt_g), warmup=3)
    assert cap.launches >= 3                 # forward, reverse recurrence + contraction (one fused launch), reduce
    for _ in range(2):
        cap()
    torch.cuda.synchronize()
    for a, b in zip(plist_e, plist_g):       # 3 eager warm-ups + 2 replays == 5 eager steps, bit for bit
        assert torch.equal(a, b)


def test_tcgen05_last_state_only_and_chunked_carry():
    """want_states=False writes no [B,T,H] tensor at all; feeding h_T of one chunk as h0 of the next is
    bit-identical to one long call (the state never leaves fp32)."""
    from kws_b200 import _lib, engine
    torch.manual_seed(21)
    p = O.init_params(32, 128)
    params = {k: v.to(dev()).contiguous() for k, v in p.tensors().items()}
    x = torch.randn(100, 30, 32, device=dev())
    kw = dict(layout="IH", batch_first=True, force_path=_lib.PATH_TCGEN05)
    full, _, _, last = engine.forward(x, params, None, want_last=True, **kw)
    none_out, _, _, last2 = engine.forward(x, params, None, want_states=False, **kw)
    assert none_out is None and torch.equal(last, last2)
    a, _, _, ha = engine.forward(x[:, :11].contiguous(), params, None, want_last=True, **kw)
    b, _, _, hb = engine.forward(x[:, 11:].contiguous(), params, ha, want_last=True, **kw)
    assert torch.equal(torch.cat([a, b], 1), full) and torch.equal(hb, last)
    # a strided (non-contiguous) time slice goes through the TMA map without a copy
    c, _, _, _ = engine.forward(x[:, 11:], params, ha, **kw)
    assert torch.equal(c, b)


def test_tcgen05_extreme_biases_and_saturated_gates():
    """Units whose gate and update biases are far apart leave the shared-exponential form (e_u = e_g^2 * ratio)
    for the two-exponential form; saturated gates (|pre| ~ 40) must stay finite."""
    from kws_b200 import _lib, engine
    torch.manual_seed(8)
    p = O.init_params(32, 128)
    p.bias_gate[0, :8] += 12.0; p.bias_update[0, 8:16] -= 11.0; p.bias_gate[0, 16:24] -= 30.0; p.bias_update[0, 24:32] += 25.0
    x = torch.randn(50, 7, 32)
    x[:5] *= 8.0                                            # |pre| up to ~20 on top of the bias offsets
    ref = O.unroll(x, p, None, True)
    params = {k: v.to(dev()).contiguous() for k, v in p.tensors().items()}
    out = engine.forward(x.to(dev()), params, None, layout="IH", batch_first=True, force_path=_lib.PATH_TCGEN05)[0]
    assert torch.isfinite(out).all()
    assert state_ratio(out, ref) <= 1.0


def test_tcgen05_backward_with_bf16_inputs():
    """bf16 x (config 5 style): the dW contraction reads the bf16 tile directly (no lo part); parity against
    autograd of the oracle fed the same bf16-rounded inputs."""
    from kws_b200 import engine
    from gpu_helpers import grad_ratio
    torch.manual_seed(31)
    B, T, I = 90, 6, 32
    p = O.init_params(I, 128)
    x = torch.randn(B, T, I).bfloat16()
    go = torch.randn(B, T, 128) / B
    gref = O.autograd_grads(x.float(), p, torch.zeros(B, 128), go, True)
    params = {k: v.to(dev()).contiguous() for k, v in p.tensors().items()}
    xg, gog = x.to(dev()), go.to(dev())
    out, z_s, c_s, _ = engine.forward(xg, params, None, layout="IH", batch_first=True, save_for_backward=True)
    ref = O.unroll(x.float(), p, None, True)
    assert state_ratio(out, ref) <= 1.0
    g = engine.backward(gog, xg, out, z_s, c_s, params, None, layout="IH", batch_first=True)
    torch.cuda.synchronize()
    for k in p.tensors():
        assert grad_ratio(g[k], gref[k]) <= 1.0, (k, grad_ratio(g[k], gref[k]))


@pytest.mark.parametrize("I,H", [(32, 128), (64, 256)])
def test_tcgen05_nan_input_poisons_its_own_row_only(I, H):
    """A NaN feature makes that row's states NaN from its step on (as rnn.py does) -- the activation clamps keep NaN
    (max.NaN / min.NaN) -- and leaves the other rows of the tile untouched."""
    from kws_b200 import engine
    torch.manual_seed(4)
    p = O.init_params(I, H)
    x = torch.randn(64, 5, I)
    ref = O.unroll(x, p, None, True)
    xg = x.to(dev()); xg[9, 2, 1] = float("nan")
    params = {k: v.to(dev()).contiguous() for k, v in p.tensors().items()}
    assert engine.forward_plan(xg, params, None, layout="IH", batch_first=True) == "tcgen05"
    out = engine.forward(xg, params, None, layout="IH", batch_first=True)[0]
    torch.cuda.synchronize()
    assert torch.isnan(out[9, 2:]).all() and torch.isfinite(out[9, :2]).all()
    keep = [r for r in range(64) if r != 9]
    assert torch.isfinite(out[keep]).all()
    if I + H < 320:
        assert state_ratio(out[keep], ref[keep]) <= 1.0


@pytest.mark.parametrize("B,T,I", [(96, 12, 32), (2048, 20, 32), (5000, 6, 16)])
def test_fused_reverse_recurrence_and_contraction_equal_the_two_launches(tuning, B, T, I):
    """One launch whose contraction CTAs follow the recurrence CTAs through published progress (the default when the
    recurrence leaves SMs free) gives the very bits of the two separate launches (FGRNN_TC_BWD_FUSED=0): the partial sums
    are per contraction CTA and the CTA count is part of the summation order, so the comparison pins the count too."""
    from kws_b200 import engine
    torch.manual_seed(B + T)
    p = O.init_params(I, 128)
    params = {k: v.to(dev()).contiguous() for k, v in p.tensors().items()}
    x = torch.randn(T, B, I, device=dev())
    go = torch.randn(T, B, 128, device=dev()) / B
    out, z_s, c_s, _ = engine.forward(x, params, None, layout="IH", save_for_backward=True)
    tuning("FGRNN_TC_BWD_FUSED", "1")
    a = engine.backward(go, x, out, z_s, c_s, params, None, layout="IH")
    a2 = engine.backward(go, x, out, z_s, c_s, params, None, layout="IH")
    tuning("FGRNN_TC_BWD_FUSED", "0")
    b = engine.backward(go, x, out, z_s, c_s, params, None, layout="IH")
    torch.cuda.synchronize()
    from gpu_helpers import grad_ratio
    for k in a:
        assert torch.equal(a[k], a2[k]), k                     # run-to-run identical
        if k in ("W", "U"):                                    # different CTA counts: different (fixed) summation orders
            assert grad_ratio(a[k], b[k].cpu()) <= 1.0, k
        else:
            assert torch.equal(a[k], b[k]), k
