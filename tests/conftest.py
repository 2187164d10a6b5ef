import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")

# north-star tolerances (BASELINE.json): fp32 hidden states / logits rtol 1e-5 atol 1e-6,
# gradients rtol 1e-4 with atol = 1e-4*max|g_ref| per tensor
STATE_RTOL, STATE_ATOL = 1e-5, 1e-6
GRAD_RTOL = 1e-4


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box via gpurun)")


def pytest_collection_modifyitems(config, items):
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


def load_golden(name):
    path = os.path.join(GOLDEN_DIR, name + ".npz")
    with np.load(path, allow_pickle=False) as f:
        return {k: f[k] for k in f.files}


def golden_case_names():
    return sorted(f[:-4] for f in os.listdir(GOLDEN_DIR)
                  if f.endswith(".npz") and not f.startswith("model_") and not f.startswith("bn_"))


def golden_model_names():
    return sorted(f[:-4] for f in os.listdir(GOLDEN_DIR) if f.endswith(".npz") and f.startswith("model_"))


def params_from_golden(g, prefix="p_"):
    from oracle.fastgrnn_oracle import Params
    kw = {k[len(prefix):]: torch.from_numpy(v.copy()) for k, v in g.items() if k.startswith(prefix)}
    return Params(**kw)


@pytest.fixture(scope="session")
def lib():
    from kws_b200 import _lib
    return _lib.load()


@pytest.fixture
def tuning():
    """Set launcher tuning keys for one test (kws_b200._lib.set_tuning) and clear them afterwards."""
    from kws_b200 import _lib
    used = []

    def set_(name, value):
        _lib.set_tuning(name, value)
        used.append(name)
    yield set_
    for name in used:
        _lib.set_tuning(name, None)
