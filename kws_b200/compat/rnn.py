"""Top-level ``rnn`` module for the reference's unmodified callers (``model.py:6`` does
``from rnn import FastGRNN, FastGRNNCUDA, FastGRNNBatchNorm, onnx_exportable_rnn``).
Put this directory first on ``sys.path`` -- see INTEGRATION.md."""
from kws_b200.rnn import *          # noqa: F401,F403
from kws_b200.rnn import fastgrnn_cuda, utils, NON_LINEARITY  # noqa: F401
