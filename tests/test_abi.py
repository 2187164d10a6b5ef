"""CPU: the C-ABI library loads, exports every symbol include/fastgrnn_b200.h declares, its
structs match the ctypes mirror byte for byte, and descriptor validation returns the documented
error codes.  No compute call is made (there is no GPU here)."""
import ctypes as C
import os
import re
import subprocess
import sys

import pytest

from conftest import ROOT
from kws_b200 import _lib

HEADER = os.path.join(ROOT, "include", "fastgrnn_b200.h")


def declared_symbols():
    src = open(HEADER).read()
    return sorted(set(re.findall(r"FGRNN_API\s+[\w\s\*]+?\b(fgrnn_\w+)\s*\(", src)))


def test_library_is_built_in_tree():
    assert os.path.isfile(_lib.LIB_PATH), "run __graft_entry__.build() / make -C kws_b200/csrc"
    assert os.path.commonpath([ROOT, _lib.LIB_PATH]) == ROOT


def test_every_declared_symbol_is_exported_and_bound(lib):
    names = declared_symbols()
    assert len(names) >= 11
    assert set(names) == set(_lib.SYMBOLS), "ctypes binding and header disagree"
    out = subprocess.run(["nm", "-D", "--defined-only", _lib.LIB_PATH], capture_output=True, text=True, check=True).stdout
    exported = set(re.findall(r" T (fgrnn_\w+)", out))
    assert set(names) <= exported
    for n in names:
        assert getattr(lib, n) is not None
    # nothing but the C ABI leaks out of the library
    leaked = [l for l in out.splitlines() if " T " in l and "fgrnn_" not in l]
    assert not leaked, leaked


def test_library_is_sm100a_only():
    out = subprocess.run(["cuobjdump", "-lelf", _lib.LIB_PATH], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_(\d+a?)", out))
    assert archs == {"100a"}, archs


def test_struct_layout_matches_header(tmp_path):
    """Compile a C program against the real header and compare sizeof/offsetof with ctypes."""
    fields = {"FgrnnProblem": [f for f, _ in _lib.FgrnnProblem._fields_],
              "FgrnnForward": [f for f, _ in _lib.FgrnnForward._fields_],
              "FgrnnBackward": [f for f, _ in _lib.FgrnnBackward._fields_],
              "FgrnnPeerStep": [f for f, _ in _lib.FgrnnPeerStep._fields_]}
    lines = ['#include <stdio.h>', '#include <stddef.h>', '#include "fastgrnn_b200.h"', 'int main(void){']
    for s, fs in fields.items():
        lines.append('printf("%s %%zu\\n", sizeof(%s));' % (s, s))
        for f in fs:
            lines.append('printf("%s.%s %%zu\\n", offsetof(%s, %s));' % (s, f, s, f))
    lines.append('printf("ABI %d\\n", FGRNN_ABI_VERSION); return 0;}')
    src = tmp_path / "layout.c"
    src.write_text("\n".join(lines))
    exe = tmp_path / "layout"
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)], check=True)
    got = dict(l.split() for l in subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.splitlines())
    assert int(got["ABI"]) == _lib.ABI_VERSION
    for s in fields:
        cls = getattr(_lib, s)
        assert int(got[s]) == C.sizeof(cls), s
        for f in fields[s]:
            assert int(got["%s.%s" % (s, f)]) == getattr(cls, f).offset, (s, f)


def test_peer_step_descriptor_validation(lib):
    """fgrnn_sgd_allreduce_peer rejects a bad descriptor before it touches a device; the sizes of the flag / state areas are fixed."""
    assert lib.fgrnn_peer_recv_bytes(21, 4) == 2 * 4 * 11 * 16 and lib.fgrnn_peer_recv_bytes(0, 4) == 0 and lib.fgrnn_peer_state_bytes() == 33 * 4
    assert lib.fgrnn_sgd_allreduce_peer(None, None) == _lib.ERR_NULL
    d = _lib.FgrnnPeerStep()
    d.abi_version, d.world, d.rank, d.n = _lib.ABI_VERSION + 1, 2, 0, 16
    assert lib.fgrnn_sgd_allreduce_peer(C.byref(d), None) == _lib.ERR_VERSION
    d.abi_version = _lib.ABI_VERSION
    for world, rank in [(0, 0), (9, 0), (2, 2), (2, -1)]:
        d.world, d.rank = world, rank
        assert lib.fgrnn_sgd_allreduce_peer(C.byref(d), None) == _lib.ERR_SHAPE
    d.world, d.rank = 2, 1
    assert lib.fgrnn_sgd_allreduce_peer(C.byref(d), None) == _lib.ERR_NULL           # params / state missing
    d.params, d.state = 0x1000, 0x2000
    assert lib.fgrnn_sgd_allreduce_peer(C.byref(d), None) == _lib.ERR_NULL           # bucket / receive areas missing
    d.bucket = 0x3000
    d.recv[0], d.recv[1] = 0x4000, 0x5008
    assert lib.fgrnn_sgd_allreduce_peer(C.byref(d), None) == _lib.ERR_ALIGN
    assert lib.fgrnn_peer_alloc(0, 0, None, None) == _lib.ERR_NULL


def _fwd(**kw):
    d = _lib.FgrnnForward()
    p = d.p
    p.abi_version = _lib.ABI_VERSION
    p.device = 0
    p.B, p.T, p.I, p.H = 4, 3, 32, 128
    p.gate_nl, p.update_nl = 0, 2
    p.force_path = -1
    for k in ("W", "U", "bias_gate", "bias_update", "zeta", "nu", "x"):
        setattr(p, k, 0x1000)
    p.x_stride_b, p.x_stride_t = 96, 32
    d.out = 0x1000
    d.out_stride_b, d.out_stride_t = 384, 128
    for k, v in kw.items():
        if hasattr(p, k):
            setattr(p, k, v)
        else:
            setattr(d, k, v)
    return d


def test_abi_version_and_strerror(lib):
    assert lib.fgrnn_abi_version() == _lib.ABI_VERSION
    assert lib.fgrnn_strerror(0) == b"ok"
    for code in range(1, 9):
        assert lib.fgrnn_strerror(code) not in (b"ok", b"unknown error")
    assert lib.fgrnn_strerror(99) == b"unknown error"
    assert lib.fgrnn_launch_count() >= 0


@pytest.mark.parametrize("kw,code", [
    (dict(abi_version=7), _lib.ERR_VERSION),
    (dict(H=0), _lib.ERR_SHAPE), (dict(H=4096), _lib.ERR_SHAPE), (dict(I=0), _lib.ERR_SHAPE),
    (dict(B=-1), _lib.ERR_SHAPE), (dict(rW=-2), _lib.ERR_SHAPE),
    (dict(gate_nl=9), _lib.ERR_ENUM), (dict(update_nl=-1), _lib.ERR_ENUM),
    (dict(weight_layout=3), _lib.ERR_ENUM), (dict(x_dtype=5), _lib.ERR_ENUM), (dict(force_path=9), _lib.ERR_ENUM),
    (dict(W=None), _lib.ERR_NULL), (dict(U=None), _lib.ERR_NULL), (dict(rW=8), _lib.ERR_NULL),
    (dict(rU=8), _lib.ERR_NULL), (dict(zeta=None), _lib.ERR_NULL), (dict(x=None), _lib.ERR_NULL),
    (dict(out=None), _lib.ERR_NULL), (dict(x=0x1001), _lib.ERR_ALIGN),
])
def test_forward_validation_codes(lib, kw, code):
    d = _fwd(**kw)
    assert lib.fgrnn_forward(C.byref(d), None) == code
    assert lib.fgrnn_last_error_detail() != b""
    assert lib.fgrnn_forward_plan(C.byref(d)) == -1
    assert lib.fgrnn_forward_workspace_bytes(C.byref(d)) == 0


def test_null_descriptor(lib):
    assert lib.fgrnn_forward(None, None) == _lib.ERR_NULL
    assert lib.fgrnn_backward(None, None) == _lib.ERR_NULL


def test_empty_problem_is_a_noop(lib):
    # B*T == 0 returns OK before touching the device
    assert lib.fgrnn_forward(C.byref(_fwd(B=0)), None) == _lib.OK
    assert lib.fgrnn_forward(C.byref(_fwd(T=0)), None) == _lib.OK


def test_forward_workspace_and_plan(lib):
    d = _fwd()
    assert lib.fgrnn_forward_plan(C.byref(d)) in (_lib.PATH_GENERIC, _lib.PATH_SMEM, _lib.PATH_TCGEN05)
    ih = lib.fgrnn_forward_workspace_bytes(C.byref(d))
    d.p.weight_layout = _lib.LAYOUT_HI
    hi = lib.fgrnn_forward_workspace_bytes(C.byref(d))
    assert ih % 256 == 0 and hi % 256 == 0


def test_backward_validation_and_workspace(lib):
    g = _lib.FgrnnBackward()
    C.memmove(C.byref(g.p), C.byref(_fwd().p), C.sizeof(_lib.FgrnnProblem))
    assert lib.fgrnn_backward(C.byref(g), None) == _lib.ERR_NULL           # grad_h missing
    g.grad_h = g.hs = g.z_s = g.c_s = 0x1000
    g.d_W = g.d_U = g.d_bias_gate = 0x1000
    need = lib.fgrnn_backward_workspace_bytes(C.byref(g))
    B, T, I, H = 4, 3, 32, 128
    assert need >= B * T * H * 4                                            # dPre is materialised
    assert lib.fgrnn_backward_plan(C.byref(g)) >= 0


def test_grad_bucket_layout(lib):
    p = _fwd().p
    offs = (C.c_int64 * 8)()
    n = lib.fgrnn_grad_bucket_layout(C.byref(p), offs)
    assert n == 32 * 128 + 128 * 128 + 128 + 128 + 2 == 20738
    assert list(offs) == [0, -1, 4096, -1, 20480, 20608, 20736, 20737]
    p.rW, p.rU, p.H = 16, 32, 256
    n = lib.fgrnn_grad_bucket_layout(C.byref(p), offs)
    assert n == 512 + 4096 + 8192 + 8192 + 512 + 2 == 21506
    assert lib.fgrnn_grad_bucket_layout(None, offs) == -1


def test_missing_library_fails_loudly(monkeypatch, tmp_path):
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", str(tmp_path / "nope.so"))
    with pytest.raises(_lib.FastGRNNLibraryError, match="no CPU or PyTorch fallback"):
        _lib.load()


def test_plan_selection_between_kernel_families(lib):
    """fgrnn_*_plan is pure host logic: the tcgen05 family takes the flagship shape, everything it does not
    cover falls through to the FFMA / generic families; forcing an unsupported family is an error, not a fallback."""
    plan = lambda **kw: lib.fgrnn_forward_plan(C.byref(_fwd(**kw)))
    assert plan() == _lib.PATH_TCGEN05                                       # full rank, H=128, I=32, aligned
    assert plan(I=64, x_stride_b=192, x_stride_t=64) == _lib.PATH_TCGEN05
    assert plan(x_dtype=_lib.BF16, x=0x1000) == _lib.PATH_TCGEN05
    assert plan(I=28, x_stride_b=84, x_stride_t=28) == _lib.PATH_SMEM        # I % 8 != 0: FFMA family
    assert plan(x=0x1004) == _lib.PATH_GENERIC                               # x not 16-byte aligned: no vector / TMA access
    assert plan(gate_nl=2) == _lib.PATH_TCGEN05                              # tanh gate: hoisted-projection kernels (fgrnn_tc_wx.cu)
    assert plan(H=256, out_stride_b=768, out_stride_t=256) == _lib.PATH_TCGEN05   # H = 256 full rank: CTA pair
    assert plan(I=256, x_stride_b=768, x_stride_t=256) == _lib.PATH_TCGEN05  # the default model's second layer
    assert plan(I=264, x_stride_b=792, x_stride_t=264) == _lib.PATH_GENERIC  # I > 256
    assert plan(H=64, out_stride_b=192, out_stride_t=64) == _lib.PATH_GENERIC
    assert plan(gate_nl=1) == _lib.PATH_SMEM                                 # relu gate: FFMA family
    assert plan(gate_scale=0x1000) == _lib.PATH_TCGEN05                      # folded BatchNorm factors: wide tcgen05 or generic only
    assert plan(gate_scale=0x1000, I=28, x_stride_b=84, x_stride_t=28) == _lib.PATH_GENERIC
    assert plan(gate_scale=0x1000, force_path=_lib.PATH_SMEM) == -1
    assert plan(rU=32, U1=0x1000, U2=0x1000) == _lib.PATH_GENERIC            # low rank
    lr = dict(H=256, rW=16, rU=32, W1=0x1000, W2=0x1000, U1=0x1000, U2=0x1000, out_stride_b=768, out_stride_t=256)
    assert plan(**lr) == _lib.PATH_LOWRANK                                   # C4 shape: persistent low-rank FFMA kernel
    assert plan(**dict(lr, gate_nl=2)) == _lib.PATH_LOWRANK                  # every nonlinearity
    assert plan(**dict(lr, rU=30)) == _lib.PATH_GENERIC                      # rank not a multiple of 4
    assert plan(**dict(lr, I=64, rW=32, rU=64, x_stride_b=192, x_stride_t=64)) == _lib.PATH_GENERIC   # does not fit shared memory
    assert plan(**dict(lr, U1=0x1004)) == _lib.PATH_GENERIC                  # weights not 16-byte aligned
    assert plan(**dict(lr, force_path=_lib.PATH_GENERIC)) == _lib.PATH_GENERIC
    assert plan(force_path=_lib.PATH_LOWRANK) == -1                          # full-rank problem forced onto the low-rank family
    assert plan(force_path=_lib.PATH_SMEM) == _lib.PATH_SMEM
    d = _fwd(I=28, x_stride_b=84, x_stride_t=28, force_path=_lib.PATH_TCGEN05)
    assert lib.fgrnn_forward(C.byref(d), None) == _lib.ERR_SHAPE and lib.fgrnn_forward_plan(C.byref(d)) == -1

    g = _lib.FgrnnBackward()
    C.memmove(C.byref(g.p), C.byref(_fwd().p), C.sizeof(_lib.FgrnnProblem))
    g.grad_h = g.hs = g.z_s = g.c_s = 0x1000
    g.grad_stride_b, g.grad_stride_t, g.hs_stride_b, g.hs_stride_t = 384, 128, 384, 128
    g.d_W = g.d_U = 0x1000
    assert lib.fgrnn_backward_plan(C.byref(g)) == _lib.PATH_TCGEN05
    g.grad_stride_b = 386                                                    # rows no longer 16-byte aligned
    assert lib.fgrnn_backward_plan(C.byref(g)) == _lib.PATH_GENERIC
    g.grad_stride_b = 384
    g.p.gate_nl = 2
    assert lib.fgrnn_backward_plan(C.byref(g)) == _lib.PATH_SMEM
    g.p.gate_nl = 0
    g.p.force_path = _lib.PATH_LOWRANK                                       # forward-only family: backward plans generic
    assert lib.fgrnn_backward_plan(C.byref(g)) == _lib.PATH_GENERIC
