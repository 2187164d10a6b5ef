"""kws_b200 -- B200-native (sm_100a) FastGRNN recurrence engine behind the operator surface of
adithom/KWS (`rnn.py` modules + the `fastgrnn_cuda` extension signatures).

    from kws_b200 import rnn                      # FastGRNN, FastGRNNCUDA, cells, Functions
    from kws_b200 import fastgrnn_cuda            # forward / backward / forward_unroll / backward_unroll
    from kws_b200 import engine                   # tensor-level API over the C ABI
    from kws_b200 import sharding                 # batch-sharded inference / data-parallel gradients

To let the reference's unmodified `model.py` / `trainClassifier.py` pick this implementation up, put
`kws_b200/compat` first on `sys.path` (it provides top-level `rnn` and `fastgrnn_cuda` modules); see
INTEGRATION.md.  The CUDA library is loaded on first use and there is no fallback if it is missing.
"""
__version__ = "0.1.0"

from . import _lib  # noqa: F401  (ctypes binding; does not load the library until first use)

__all__ = ["rnn", "fastgrnn_cuda", "engine", "sharding", "build_library", "library_path"]


def build_library(verbose: bool = False) -> str:
    """Compile libfastgrnn_b200.so in-tree for sm_100a (nvcc)."""
    return _lib.build(verbose=verbose)


def library_path() -> str:
    return _lib.LIB_PATH


def __getattr__(name):
    if name in ("rnn", "fastgrnn_cuda", "engine", "sharding", "streaming"):
        import importlib
        return importlib.import_module("." + name, __name__)
    raise AttributeError(name)
