"""ctypes binding of ``libfastgrnn_b200.so`` (the C ABI in ``include/fastgrnn_b200.h``).

The structures below mirror the header field for field.  There is no fallback:
if the library has not been built (``python -c "import __graft_entry__ as g;
g.build()"`` or ``make -C kws_b200/csrc``) every entry point raises.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
import threading

ABI_VERSION = 3

# enums (include/fastgrnn_b200.h)
OK, ERR_NULL, ERR_SHAPE, ERR_ENUM, ERR_ALIGN, ERR_WORKSPACE, ERR_CUDA, ERR_DEVICE, ERR_VERSION = range(9)
NL = {"sigmoid": 0, "relu": 1, "tanh": 2, "quantTanh": 3, "quantSigm": 4, "quantSigm4": 5}
NL_NAMES = {v: k for k, v in NL.items()}
LAYOUT_IH, LAYOUT_HI = 0, 1
F32, BF16 = 0, 1
PATH_AUTO, PATH_GENERIC, PATH_SMEM, PATH_TCGEN05, PATH_LOWRANK = -1, 0, 1, 2, 3
PATH_NAMES = {0: "generic", 1: "smem", 2: "tcgen05", 3: "lowrank"}

_HERE = os.path.dirname(os.path.abspath(__file__))
# KWS_B200_LIB points the binding at another build of the same library (developer builds: `make fuzz`)
LIB_PATH = os.environ.get("KWS_B200_LIB") or os.path.join(_HERE, "lib", "libfastgrnn_b200.so")
CSRC_DIR = os.path.join(_HERE, "csrc")

_fp = C.c_void_p   # device pointers travel as integers


class FgrnnProblem(C.Structure):
    _fields_ = [
        ("abi_version", C.c_int32), ("device", C.c_int32),
        ("B", C.c_int32), ("T", C.c_int32), ("I", C.c_int32), ("H", C.c_int32),
        ("rW", C.c_int32), ("rU", C.c_int32),
        ("gate_nl", C.c_int32), ("update_nl", C.c_int32),
        ("weight_layout", C.c_int32), ("x_dtype", C.c_int32),
        ("force_path", C.c_int32), ("reserved0", C.c_int32),
        ("W", _fp), ("U", _fp), ("W1", _fp), ("W2", _fp), ("U1", _fp), ("U2", _fp),
        ("bias_gate", _fp), ("bias_update", _fp), ("zeta", _fp), ("nu", _fp),
        ("x", _fp), ("x_stride_b", C.c_int64), ("x_stride_t", C.c_int64),
        ("h0", _fp),
        ("gate_scale", _fp), ("update_scale", _fp),
    ]


class FgrnnForward(C.Structure):
    _fields_ = [
        ("p", FgrnnProblem),
        ("out", _fp), ("out_stride_b", C.c_int64), ("out_stride_t", C.c_int64),
        ("h_last", _fp), ("save_z", _fp), ("save_c", _fp),
        ("workspace", _fp), ("workspace_bytes", C.c_size_t),
    ]


class FgrnnBackward(C.Structure):
    _fields_ = [
        ("p", FgrnnProblem),
        ("grad_h", _fp), ("grad_stride_b", C.c_int64), ("grad_stride_t", C.c_int64),
        ("hs", _fp), ("hs_stride_b", C.c_int64), ("hs_stride_t", C.c_int64),
        ("z_s", _fp), ("c_s", _fp),
        ("d_x", _fp), ("dx_stride_b", C.c_int64), ("dx_stride_t", C.c_int64),
        ("d_W", _fp), ("d_U", _fp), ("d_W1", _fp), ("d_W2", _fp), ("d_U1", _fp), ("d_U2", _fp),
        ("d_bias_gate", _fp), ("d_bias_update", _fp), ("d_zeta", _fp), ("d_nu", _fp),
        ("d_h0", _fp),
        ("workspace", _fp), ("workspace_bytes", C.c_size_t),
        ("grad_t0", C.c_int32), ("reserved1", C.c_int32),
    ]


PEER_MAX_RANKS, PEER_HANDLE_BYTES = 8, 64


class FgrnnPeerStep(C.Structure):
    _fields_ = [
        ("abi_version", C.c_int32), ("device", C.c_int32), ("world", C.c_int32), ("rank", C.c_int32),
        ("params", _fp), ("reduced", _fp),
        ("bucket", _fp), ("recv", _fp * PEER_MAX_RANKS),
        ("state", _fp), ("n", C.c_int64), ("lr", C.c_float), ("grad_scale", C.c_float),
    ]


# every symbol include/fastgrnn_b200.h declares: name -> (restype, argtypes)
SYMBOLS = {
    "fgrnn_forward_workspace_bytes": (C.c_size_t, [C.POINTER(FgrnnForward)]),
    "fgrnn_backward_workspace_bytes": (C.c_size_t, [C.POINTER(FgrnnBackward)]),
    "fgrnn_forward_plan": (C.c_int, [C.POINTER(FgrnnForward)]),
    "fgrnn_backward_plan": (C.c_int, [C.POINTER(FgrnnBackward)]),
    "fgrnn_forward": (C.c_int, [C.POINTER(FgrnnForward), C.c_void_p]),
    "fgrnn_backward": (C.c_int, [C.POINTER(FgrnnBackward), C.c_void_p]),
    "fgrnn_grad_bucket_layout": (C.c_int64, [C.POINTER(FgrnnProblem), C.POINTER(C.c_int64)]),
    "fgrnn_strerror": (C.c_char_p, [C.c_int]),
    "fgrnn_last_error_detail": (C.c_char_p, []),
    "fgrnn_abi_version": (C.c_int, []),
    "fgrnn_launch_count": (C.c_uint64, []),
    "fgrnn_head_workspace_bytes": (C.c_size_t, [C.c_int32, C.c_int32, C.c_int32]),
    "fgrnn_head_nll": (C.c_int, [C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                 C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_size_t, C.c_int32, C.c_int32, C.c_int32,
                                 C.c_int32, C.c_void_p]),
    "fgrnn_sgd_flat": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_float, C.c_float, C.c_int32, C.c_void_p]),
    "fgrnn_peer_recv_bytes": (C.c_size_t, [C.c_int64, C.c_int32]),
    "fgrnn_peer_state_bytes": (C.c_size_t, []),
    "fgrnn_peer_alloc": (C.c_int, [C.c_size_t, C.c_int32, C.POINTER(C.c_void_p), C.c_char_p]),
    "fgrnn_peer_open": (C.c_int, [C.c_char_p, C.c_int32, C.POINTER(C.c_void_p)]),
    "fgrnn_peer_close": (C.c_int, [C.c_void_p, C.c_int32]),
    "fgrnn_peer_free": (C.c_int, [C.c_void_p, C.c_int32]),
    "fgrnn_sgd_allreduce_peer": (C.c_int, [C.POINTER(FgrnnPeerStep), C.c_void_p]),
    "fgrnn_debug_poison_onchip": (C.c_int, [C.c_int, C.c_void_p]),
    "fgrnn_debug_set_tuning": (C.c_int, [C.c_char_p, C.c_char_p]),
    "fgrnn_ingest_bft": (C.c_int, [C.c_void_p, C.c_int64, C.c_int64, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p,
                                   C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_void_p]),
}

_lib = None
_lock = threading.Lock()


class FastGRNNLibraryError(RuntimeError):
    pass


def build(verbose: bool = False) -> str:
    """Compile the library in-tree for sm_100a with nvcc (``make -C kws_b200/csrc``)."""
    res = subprocess.run(["make", "-C", CSRC_DIR], capture_output=True, text=True)
    if verbose or res.returncode != 0:
        print(res.stdout)
        print(res.stderr)
    if res.returncode != 0:
        raise FastGRNNLibraryError("building libfastgrnn_b200.so failed:\n" + res.stderr[-4000:])
    return LIB_PATH


def load():
    """Load the library (once) and bind every symbol. Raises if it is missing -- no fallback."""
    global _lib
    if _lib is not None:
        return _lib
    with _lock:
        if _lib is not None:
            return _lib
        if not os.path.isfile(LIB_PATH):
            raise FastGRNNLibraryError(
                "kws_b200: %s is missing. Build it with `make -C %s` (needs nvcc, sm_100a). "
                "There is no CPU or PyTorch fallback for the FastGRNN path." % (LIB_PATH, CSRC_DIR))
        import torch  # noqa: F401  -- makes the process-wide libcudart.so.12 resolvable by soname
        lib = C.CDLL(LIB_PATH)
        for name, (restype, argtypes) in SYMBOLS.items():
            fn = getattr(lib, name)
            fn.restype = restype
            fn.argtypes = argtypes
        got = lib.fgrnn_abi_version()
        if got != ABI_VERSION:
            raise FastGRNNLibraryError("ABI mismatch: library %d, binding %d" % (got, ABI_VERSION))
        _lib = lib
    return _lib


def check(rc: int, what: str) -> None:
    """Map a non-zero return code to RuntimeError (the reference raises RuntimeError through
    AT_ASSERTM -> c10::Error for the same conditions, cuda/fastgrnn_cuda.cpp:69-71)."""
    if rc == OK:
        return
    lib = load()
    msg = lib.fgrnn_strerror(rc).decode()
    detail = lib.fgrnn_last_error_detail().decode()
    raise RuntimeError("%s: %s%s" % (what, msg, (": " + detail) if detail else ""))


def set_tuning(name: str, value=None) -> None:
    """Override a launcher tuning key (``FGRNN_TC_NS`` ...; the environment only sets the process-start default)."""
    check(load().fgrnn_debug_set_tuning(name.encode(), None if value is None else str(value).encode()), "set_tuning")


def launch_count() -> int:
    return int(load().fgrnn_launch_count())
