"""Developer tool: per-step timeline of the low-rank tensor-core kernel (CTA 0, steps 16..23) from the clock64 stamps of
the `make trace` build.  Run with KWS_B200_LIB=kws_b200/lib/libfastgrnn_b200_trace.so.  argv: B"""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from kws_b200 import _lib, engine
from oracle import fastgrnn_oracle as O
B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
dev = torch.device("cuda:0")
torch.manual_seed(0)
p = O.init_params(32, 256, 16, 32)
params = {k: v.to(dev).contiguous() for k, v in p.tensors().items()}
x = torch.randn(B, 99, 32, device=dev)
for _ in range(3):
    engine.forward(x, params, None, layout="IH", batch_first=True)
torch.cuda.synchronize()
buf = (C.c_longlong * 256)()
lib = _lib.load()
assert lib.fgrnn_debug_tl_trace(buf) == 0
names = ["mma:HREADY", "mma:S1 issued", "mma:SREADY", "mma:S2 issued", "hop:D1FULL", "hop:SREADY arr", "epi:DFULL", "epi:HREADY arr",
         "epi:stores", "epi0:DFULL", "epi0:HREADY arr"]
t0 = buf[0]
for t in range(8):
    for s in range(2):
        st = [buf[(t * 2 + s) * 16 + k] - t0 for k in range(11)]
        print("t=%d s=%d " % (16 + t, s) + "  ".join("%s %d" % (n.split(":")[1] if False else n, v) for n, v in zip(names, st)))
