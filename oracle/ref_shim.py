"""TEST INFRASTRUCTURE ONLY -- loader for the unmodified reference ``rnn.py``.

Only usable where ``/root/reference`` exists (the build container); it does not
travel to the GPU box.  Nothing in ``-m gpu`` tests, ``smoke()`` or
``bench.py`` may call this at run time -- they use the committed fixtures in
``tests/golden/`` minted by ``oracle/make_golden.py`` through this loader.

The reference files are untouched; two *test-side* shims make the committed
code runnable:

* D1  ``BaseRNN.forward`` passes ``training=`` to every cell (rnn.py:621,659)
      but ``FastGRNNCell.forward(self, input, state)`` (rnn.py:273) takes no
      such kwarg -> wrap it to swallow the kwarg.
* D2  ``FastGRNNCell.forward`` reads ``self.W.device`` (rnn.py:274) which does
      not exist when ``wRank`` is set (rnn.py:246-250) -> per-instance subclass
      exposing ``W`` as an alias of ``W1``.
"""
from __future__ import annotations

import importlib
import os
import sys

REF_ROOT = os.environ.get("KWS_REFERENCE_ROOT", "/root/reference")


def available() -> bool:
    return os.path.isfile(os.path.join(REF_ROOT, "rnn.py"))


_cache = {}


def load():
    """Import the reference's ``rnn`` and ``model`` modules with the shims."""
    if "rnn" in _cache:
        return _cache["rnn"], _cache["model"]
    if not available():
        raise RuntimeError("reference tree not present at %s" % REF_ROOT)
    saved = {k: sys.modules.pop(k, None) for k in ("rnn", "utils", "model")}
    sys.path.insert(0, REF_ROOT)
    try:
        rnn = importlib.import_module("rnn")
        model = importlib.import_module("model")
    finally:
        sys.path.remove(REF_ROOT)
        for k in ("rnn", "utils", "model"):
            mod = sys.modules.pop(k, None)
            _cache.setdefault(k, mod)
            if saved[k] is not None:
                sys.modules[k] = saved[k]
    _orig = rnn.FastGRNNCell.forward

    def forward(self, input, state, training=True):                     # D1
        return _orig(self, input, state)
    rnn.FastGRNNCell.forward = forward
    return rnn, model


def make_fastgrnn(*args, **kwargs):
    """``rnn.FastGRNN(...)`` made runnable for low-rank W (D2)."""
    rnn, _ = load()
    m = rnn.FastGRNN(*args, **kwargs)
    if m.cell._wRank is not None:
        class _LRCell(rnn.FastGRNNCell):
            W = property(lambda s: s._parameters["W1"])
        m.cell.__class__ = _LRCell
    return m


def params_of(m):
    """Reference module -> oracle ``Params`` (shares storage)."""
    from .fastgrnn_oracle import Params
    c = m.cell
    kw = {k: getattr(c, k) for k in ("bias_gate", "bias_update", "zeta", "nu")}
    for k in ("W", "U", "W1", "W2", "U1", "U2"):
        if k in c._parameters:
            kw[k] = c._parameters[k]
    return Params(**kw)
