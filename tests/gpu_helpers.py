"""Shared helpers for the GPU parity tests (product path = kws_b200 through the C ABI; the
oracle and the golden fixtures are the checker only)."""
import numpy as np
import torch

from conftest import GRAD_RTOL, STATE_ATOL, STATE_RTOL
from oracle import fastgrnn_oracle as O


def dev():
    return torch.device("cuda:0")


def state_ratio(got, ref):
    """max |got-ref| / (atol + rtol|ref|) with the north-star fp32 tolerances; <= 1 passes."""
    return O.tolerance_ratio(got.detach().cpu(), torch.as_tensor(ref), STATE_RTOL, STATE_ATOL)


def grad_ratio(got, ref):
    return O.grad_tolerance_ratio(got.detach().cpu(), torch.as_tensor(ref), GRAD_RTOL)


def golden_dims(g):
    x = g["x"]
    bf = bool(g["batch_first"])
    B, T = (x.shape[0], x.shape[1]) if bf else (x.shape[1], x.shape[0])
    H = g["p_bias_gate"].shape[1]
    wR = g["p_W1"].shape[1] if "p_W1" in g else None
    uR = g["p_U1"].shape[1] if "p_U1" in g else None
    return dict(B=B, T=T, I=x.shape[2], H=H, wRank=wR, uRank=uR, batch_first=bf,
                gate=str(g["gate"]), update=str(g["update"]))


def load_cell_params(cell, params, transpose):
    """Copy oracle-layout arrays into a module (IH layout as is, HI layout transposed)."""
    with torch.no_grad():
        for k, v in params.items():
            t = torch.as_tensor(v)
            if transpose and k in ("W", "U", "W1", "W2", "U1", "U2"):
                t = t.t()
            getattr(cell, k).copy_(t.contiguous())


def fastgrnn_from_golden(g, prefix="p_"):
    from kws_b200 import rnn
    d = golden_dims(g)
    m = rnn.FastGRNN(d["I"], d["H"], gate_nonlinearity=d["gate"], update_nonlinearity=d["update"],
                     wRank=d["wRank"], uRank=d["uRank"], batch_first=d["batch_first"])
    load_cell_params(m.cell, {k[len(prefix):]: v for k, v in g.items() if k.startswith(prefix)}, False)
    return m.to(dev()), d


def fastgrnn_cuda_from_golden(g, prefix="p_"):
    from kws_b200 import rnn
    d = golden_dims(g)
    m = rnn.FastGRNNCUDA(d["I"], d["H"], gate_nonlinearity=d["gate"], update_nonlinearity=d["update"],
                         wRank=d["wRank"], uRank=d["uRank"], batch_first=d["batch_first"])
    load_cell_params(m, {k[len(prefix):]: v for k, v in g.items() if k.startswith(prefix)}, True)
    return m, d


def check_out_against_golden(out, g, d):
    tdim = 1 if d["batch_first"] else 0
    out = out.detach().cpu()
    worst = state_ratio(out.select(tdim, d["T"] - 1), g["out_last"])
    if "out" in g:
        worst = max(worst, state_ratio(out, g["out"]))
    else:
        keep = torch.from_numpy(g["keep_t"])
        worst = max(worst, state_ratio(out.index_select(tdim, keep), g["out_keep"]))
    return worst


def golden_grad_out(g, shape):
    if "grad_out" in g:
        return torch.from_numpy(g["grad_out"].copy())
    return torch.randn(shape, generator=torch.Generator().manual_seed(int(g["grad_out_seed"])))


def assert_state_parity(out, x, p, h0, batch_first, gate="sigmoid", update="tanh"):
    """Hidden-state parity with the north-star tolerance (rtol 1e-5, atol 1e-6).

    Up to I + H = 288 the bar is the plain one: within the tolerance of the fp32 oracle.  The tolerance is only about
    twice the fp32 rounding noise of the reference at I + H = 160; at I + H >= 320 the fp32 oracle itself sits 0.5-1.1x
    the tolerance away from an fp64 evaluation of the same math (rnn.py:273-297), and an implementation whose
    pre-activations are exactly rounded still scores 0.96-1.04 against it -- no summation order but the oracle's own can
    match it to 1.0.  The bar for those shapes: within the tolerance of the fp64 evaluation, and no further from the fp32
    oracle than the tolerance plus the oracle's own deviation (triangle inequality).  With a tanh GATE (z in (-1, 1)
    multiplies the state; the map is not contractive, rounding differences are amplified from step to step) the oracle is
    1.3-2.1x the tolerance away from the fp64 evaluation on the wide shapes; there the bar is to stay within twice the
    reference's own fp32 deviation.  The same triangle bar applies when the fp32 oracle computed on this host is outside
    its normal noise band (> 0.7; 0.33-0.47 is normal on the flagship shape): then the host's CPU arithmetic, not the
    kernel, is what deviates, and the fp64 evaluation is the arbiter."""
    ref = O.unroll(x.float(), p, None if h0 is None else h0.clone().unsqueeze(0), batch_first, gate, update)
    r_oracle = state_ratio(out, ref)
    p64 = O.Params(**{k: v.double() for k, v in p.tensors().items()})
    tr = O.unroll_functional(x.double(), p64, None if h0 is None else h0.double(), batch_first, gate, update)
    r_truth = state_ratio(out.double(), tr)
    r_ot = state_ratio(ref.double(), tr)
    narrow = p.input_size + p.hidden_size < 320 and gate == "sigmoid"
    if not narrow or r_ot > 0.7:
        print("state parity (I=%d H=%d gate=%s): vs fp64 %.3f, vs oracle %.3f (oracle vs fp64 %.3f) on %s"
              % (p.input_size, p.hidden_size, gate, r_truth, r_oracle, r_ot, _cpu_model()))
    if narrow and r_ot <= 0.7:          # the usual case: the fp32 oracle is within its normal noise (0.33-0.47)
        assert r_oracle <= 1.0, (r_truth, r_oracle, r_ot)
        return r_oracle
    assert r_truth <= (1.0 if gate == "sigmoid" else max(1.0, 2.0 * r_ot)), (r_truth, r_oracle, r_ot)
    assert r_oracle <= max(1.0, r_truth + r_ot), (r_truth, r_oracle, r_ot)
    return r_oracle


def _cpu_model():
    try:
        for line in open("/proc/cpuinfo"):
            if line.startswith("model name"):
                return line.split(":", 1)[1].strip()
    except OSError:
        pass
    return "unknown cpu"
