"""Developer tool: a few forward launches of the BASELINE config-4 shape (for ncu / timing).  argv: B [T] [reps]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from kws_b200 import engine
from oracle import fastgrnn_oracle as O
B = int(sys.argv[1]); T = int(sys.argv[2]) if len(sys.argv) > 2 else 99; reps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
dev = torch.device("cuda:0")
torch.manual_seed(0)
p = O.init_params(32, 256, 16, 32)
params = {k: v.to(dev).contiguous() for k, v in p.tensors().items()}
x = torch.randn(B, T, 32, device=dev)
out = torch.empty(B, T, 256, device=dev)
for _ in range(3):
    engine.forward(x, params, None, layout="IH", batch_first=True, out=out)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(reps):
    engine.forward(x, params, None, layout="IH", batch_first=True, out=out)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / reps
print("B=%d T=%d: %.4f ms per forward, %.1f GB/s algorithmic" % (B, T, ms, B * T * (32 + 256) * 4 / ms / 1e6))
