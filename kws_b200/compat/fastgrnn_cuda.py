"""Top-level ``fastgrnn_cuda`` module: what ``import fastgrnn_cuda`` (rnn.py:9, commented out in the
reference) binds when this directory is on ``sys.path``."""
from kws_b200.fastgrnn_cuda import forward, backward, forward_unroll, backward_unroll  # noqa: F401
