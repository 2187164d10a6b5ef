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
