"""Developer tool: CUDA-event timing of the forward for the default model's layer shapes (and the flagship shape)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from kws_b200 import engine, _lib
from oracle import fastgrnn_oracle as O

dev = torch.device("cuda:0")
B = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
T = 99
for (I, H) in ((64, 256), (256, 128), (32, 128)):
    torch.manual_seed(0)
    p = O.init_params(I, H)
    params = {k: v.to(dev).contiguous() for k, v in p.tensors().items()}
    x = torch.randn(B, T, I, device=dev)
    out = torch.empty(B, T, H, device=dev)
    for path, name in ((-1, "auto"), (_lib.PATH_GENERIC, "generic")):
        n = 2 if path == _lib.PATH_GENERIC else 20
        for _ in range(2):
            engine.forward(x, params, None, layout="IH", batch_first=True, out=out, force_path=path)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n):
            engine.forward(x, params, None, layout="IH", batch_first=True, out=out, force_path=path)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / n
        byts = B * T * (I + H) * 4
        print("I=%d H=%d B=%d %s (%s): %.3f ms  %.1f M seq/s  %.0f GB/s algorithmic (%.2f of 6454.6)" % (
            I, H, B, name, engine.forward_plan(x, params, None, layout="IH", batch_first=True, force_path=path), ms, B / ms / 1e3, byts / ms / 1e6, byts / ms / 1e6 / 6454.6), flush=True)
