"""GPU: the keyword spotter's training step on the last hidden state (model.py:227-231,
trainClassifier.py:225-240): fused head kernel, BPTT from the last state's gradient only (``grad_t0``), flat SGD --
against torch autograd on the CPU oracle."""
import pytest
import torch

from gpu_helpers import dev, grad_ratio
from oracle import fastgrnn_oracle as O

pytestmark = pytest.mark.gpu


def _head_reference(h, W, b, labels):
    h = h.clone().requires_grad_(True); W = W.clone().requires_grad_(True); b = b.clone().requires_grad_(True)
    logp = torch.nn.functional.log_softmax(torch.nn.functional.linear(h, W, b), dim=1)      # model.py:228-230
    loss = torch.nn.functional.nll_loss(logp, labels)                                       # trainClassifier.py:236
    loss.backward()
    return loss.detach(), logp.detach(), W.grad, b.grad, h.grad


@pytest.mark.parametrize("B,H,C", [(64, 128, 13), (2048, 128, 13), (77, 128, 13), (5, 256, 16), (130, 64, 2)])
def test_head_kernel_matches_torch(B, H, C):
    from kws_b200 import train_step
    torch.manual_seed(B + H)
    h = torch.randn(B, H); W = 0.3 * torch.randn(C, H); b = 0.1 * torch.randn(C)
    labels = torch.randint(0, C, (B,))
    loss_r, logp_r, dW_r, db_r, dh_r = _head_reference(h, W, b, labels)
    loss, dW, db, dh, logp = train_step.head_nll(h.to(dev()), W.to(dev()), b.to(dev()), labels.to(dev()), want_logp=True)
    loss2, dW2, db2, dh2, _ = train_step.head_nll(h.to(dev()), W.to(dev()), b.to(dev()), labels.to(dev()))
    torch.cuda.synchronize()
    assert torch.allclose(logp.cpu(), logp_r, rtol=1e-5, atol=1e-6)          # logits tolerance of the north star
    assert torch.allclose(loss.cpu(), loss_r, rtol=1e-5, atol=1e-6)
    for got, ref in ((dW, dW_r), (db, db_r), (dh, dh_r)):
        assert grad_ratio(got, ref) <= 1.0
    # fixed summation order: launches agree bit for bit
    assert torch.equal(loss, loss2) and torch.equal(dW, dW2) and torch.equal(db, db2) and torch.equal(dh, dh2)


def test_head_kernel_on_a_strided_last_state():
    """h_T given as the last time step of a batch-first [B,T,H] tensor (row stride T*H)."""
    from kws_b200 import train_step
    torch.manual_seed(3)
    B, T, H, C = 70, 5, 128, 13
    hs = torch.randn(B, T, H); W = 0.3 * torch.randn(C, H); b = 0.1 * torch.randn(C); labels = torch.randint(0, C, (B,))
    loss_r, _, dW_r, _, dh_r = _head_reference(hs[:, -1], W, b, labels)
    loss, dW, _, dh, _ = train_step.head_nll(hs.to(dev())[:, -1], W.to(dev()), b.to(dev()), labels.to(dev()))
    assert torch.allclose(loss.cpu(), loss_r, rtol=1e-5, atol=1e-6)
    assert grad_ratio(dW, dW_r) <= 1.0 and grad_ratio(dh, dh_r) <= 1.0


@pytest.mark.parametrize("path,B,T,I,H,ranks", [
    ("tcgen05", 96, 12, 32, 128, (None, None)),
    ("tcgen05", 2048, 99, 32, 128, (None, None)),       # BASELINE config 3, one GPU's share
    ("smem", 50, 7, 32, 128, (None, None)),
    ("generic", 21, 6, 20, 48, (None, None)),
    ("generic", 21, 6, 32, 256, (16, 32)),              # low rank
])
def test_last_state_gradient_equals_the_dense_zero_padded_one(path, B, T, I, H, ranks):
    """grad_t0 = T-1 with a [1,B,H] gradient == the full [T,B,H] gradient that is zero before the last step,
    bit for bit (the kernels add the same zeros they no longer read)."""
    from kws_b200 import _lib, engine
    torch.manual_seed(T + B)
    p = O.init_params(I, H, *ranks)
    params = {k: v.to(dev()).contiguous() for k, v in p.tensors().items()}
    x = torch.randn(T, B, I, device=dev())
    force = {"tcgen05": _lib.PATH_TCGEN05, "smem": _lib.PATH_SMEM, "generic": _lib.PATH_GENERIC}[path]
    out, z_s, c_s, _ = engine.forward(x, params, None, layout="IH", save_for_backward=True, force_path=force)
    g_last = torch.randn(1, B, H, device=dev())
    dense = torch.zeros(T, B, H, device=dev()); dense[-1] = g_last[0]
    a = engine.backward(dense, x, out, z_s, c_s, params, None, layout="IH", force_path=force)
    b = engine.backward(g_last, x, out, z_s, c_s, params, None, layout="IH", force_path=force, grad_t0=T - 1)
    torch.cuda.synchronize()
    for k in a:
        assert torch.equal(a[k], b[k]), k
    # a gradient for the last THREE steps
    if T > 3:
        g3 = torch.randn(3, B, H, device=dev())
        dense = torch.zeros(T, B, H, device=dev()); dense[-3:] = g3
        a = engine.backward(dense, x, out, z_s, c_s, params, None, layout="IH", force_path=force)
        b = engine.backward(g3, x, out, z_s, c_s, params, None, layout="IH", force_path=force, grad_t0=T - 3)
        for k in a:
            assert torch.equal(a[k], b[k]), k
    with pytest.raises(RuntimeError):
        engine.backward(g_last, x, out, z_s, c_s, params, None, layout="IH", force_path=force, grad_t0=T)


@pytest.mark.parametrize("B,T,batch_first", [(64, 20, False), (96, 9, True)])
def test_fused_step_matches_autograd_on_the_oracle(B, T, batch_first):
    """Two SGD steps of LastStateTrainStep == two steps of torch autograd + torch.optim.SGD on the CPU oracle
    (FastGRNN -> out[-1] -> Linear -> log_softmax -> NLL), gradients within the north-star 1e-4."""
    from kws_b200 import rnn, train_step
    I, H, C, lr = 32, 128, 13, 0.05
    torch.manual_seed(7)
    p = O.init_params(I, H)
    head_ref = torch.nn.Linear(H, C)
    layer = rnn.FastGRNN(I, H, batch_first=batch_first)
    head = torch.nn.Linear(H, C)
    with torch.no_grad():
        for k, v in p.tensors().items():
            getattr(layer.cell, k).copy_(v)
        head.weight.copy_(head_ref.weight); head.bias.copy_(head_ref.bias)
    layer.to(dev()); head.to(dev())
    step = train_step.LastStateTrainStep(layer, head, lr)
    ref_params = {k: v.clone().requires_grad_(True) for k, v in p.tensors().items()}
    opt = torch.optim.SGD(list(ref_params.values()) + list(head_ref.parameters()), lr=lr)
    for it in range(2):
        x = torch.randn(B, T, I) if batch_first else torch.randn(T, B, I)
        labels = torch.randint(0, C, (B,))
        opt.zero_grad()
        hs = O.unroll_functional(x, O.Params(**ref_params), None, batch_first)
        last = hs[:, -1] if batch_first else hs[-1]
        loss_r = torch.nn.functional.nll_loss(torch.nn.functional.log_softmax(head_ref(last), dim=1), labels)
        loss_r.backward()
        loss = step.compute(x.to(dev()), labels.to(dev()))
        torch.cuda.synchronize()
        assert abs(float(loss) - float(loss_r.detach())) <= 1e-5 * abs(float(loss_r.detach())) + 1e-6
        for k, v in ref_params.items():
            assert grad_ratio(getattr(layer.cell, k).grad, v.grad) <= 1.0, (it, k)
        assert grad_ratio(head.weight.grad, head_ref.weight.grad) <= 1.0
        assert grad_ratio(head.bias.grad, head_ref.bias.grad) <= 1.0
        step.update(); opt.step()
        for k, v in ref_params.items():
            assert torch.allclose(getattr(layer.cell, k).detach().cpu(), v.detach(), rtol=1e-5, atol=1e-6), (it, k)
    # the modules still see the flat buffer: inference after training uses the updated weights
    assert layer.cell.W.data_ptr() == step.flat_params.data_ptr()


@pytest.mark.parametrize("n", [22415, 1001, 6])
def test_peer_step_with_one_rank_is_the_flat_sgd_step(n):
    """fgrnn_sgd_allreduce_peer at world size 1 (nothing to push, the sum is this rank's bucket) == fgrnn_sgd_flat bit for
    bit, over several steps (both parities of the receive area, the kernel's own step counters), odd lengths included;
    `reduced` receives the bucket.  The multi-rank exchange itself is checked by tests/dist_check.py under torchrun."""
    import ctypes as C
    from kws_b200 import _lib, engine, train_step
    lib = _lib.load()
    torch.manual_seed(n)
    p0 = torch.randn(n, device=dev())
    params, ref = p0.clone(), p0.clone()
    reduced = torch.empty(n, device=dev())
    recv = torch.zeros(int(lib.fgrnn_peer_recv_bytes(n, 1)) // 4, device=dev())
    state = torch.zeros(int(lib.fgrnn_peer_state_bytes()) // 4, dtype=torch.int32, device=dev())
    for step in range(3):
        g = torch.randn(n, device=dev())
        d = _lib.FgrnnPeerStep()
        d.abi_version, d.device, d.world, d.rank = _lib.ABI_VERSION, dev().index or 0, 1, 0
        d.params, d.reduced, d.bucket, d.state = params.data_ptr(), reduced.data_ptr(), g.data_ptr(), state.data_ptr()
        d.recv[0] = recv.data_ptr()
        d.n, d.lr, d.grad_scale = n, 0.05, 1.0
        _lib.check(lib.fgrnn_sgd_allreduce_peer(C.byref(d), engine._stream(dev())), "sgd_allreduce_peer")
        train_step.sgd_flat(ref, g, 0.05, 1.0)
        torch.cuda.synchronize()
        assert torch.equal(params, ref), step
        assert torch.equal(reduced, g), step
    grid = min(32, (((n + 1) // 2) + 255) // 256)
    assert state[:grid].tolist() == [3] * grid and int(state[32]) == 0
