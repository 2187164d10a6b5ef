"""Host-side helpers the ``rnn`` operator surface needs from the reference's
``utils.py`` (findCUDA :12-36, hardThreshold :53-64, supportBasedThreshold /
copySupport :66-81, countNNZ :104-114).  These are parameter-sized CPU/GPU
utilities outside the hot path; they are restated here so ``kws_b200.rnn`` does
not import the reference's top-level ``utils`` module."""
from __future__ import annotations

import os
import shutil

import numpy as np
import torch


def findCUDA():
    """CUDA toolkit root or None (utils.py:12-36): $CUDA_HOME / $CUDA_PATH, the
    directory two levels above ``nvcc``, or a conventional install path."""
    home = os.environ.get("CUDA_HOME") or os.environ.get("CUDA_PATH")
    if home is None:
        nvcc = shutil.which("nvcc")
        if nvcc:
            home = os.path.dirname(os.path.dirname(nvcc))
    if home is None:
        for cand in ("/usr/local/cuda", "/usr/local/cuda-11", "/usr/local/cuda-12"):
            if os.path.exists(cand):
                home = cand
                break
    return home


def countNNZ(A: torch.Tensor, isSparse) -> int:
    """utils.py:104-114: non-zero count if ``isSparse`` (truthy) else the element count."""
    if isSparse:
        return int(np.count_nonzero(A.detach().cpu().numpy()))
    n = 1
    for s in A.shape:
        n *= int(s)
    return n


def hard_threshold_(A: torch.Tensor, s: float) -> torch.Tensor:
    """In-place iterative-hard-thresholding step with the semantics of utils.py:53-64: keep the
    fraction ``s`` of entries with the largest magnitude (threshold = the (1-s) percentile of |A|,
    'higher' interpolation), zero the rest.  Runs on the tensor's own device."""
    with torch.no_grad():
        flat = A.detach().abs().reshape(-1)
        if flat.numel() == 0 or s >= 1.0:
            return A
        th = torch.quantile(flat.float(), max(0.0, min(1.0, 1.0 - float(s))), interpolation="higher")
        A.masked_fill_(A.abs() < th, 0.0)
    return A


def copy_support_(src: torch.Tensor, dst: torch.Tensor) -> torch.Tensor:
    """utils.py:72-81: zero the entries of ``dst`` where ``src`` is zero (in place)."""
    with torch.no_grad():
        dst.masked_fill_(src.to(dst.device) == 0.0, 0.0)
    return dst
