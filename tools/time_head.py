"""Developer tool: warm timing of the head kernel (CUDA events; eager launches and inside a CUDA graph).  argv: B"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from kws_b200 import train_step
dev = torch.device("cuda:0")
B, H, C = int(sys.argv[1]) if len(sys.argv) > 1 else 2048, 128, 13
h = torch.randn(B, H, device=dev); W = torch.randn(C, H, device=dev); b = torch.randn(C, device=dev); y = torch.randint(0, C, (B,), device=dev)
ws = torch.zeros(1 << 20, dtype=torch.uint8, device=dev); dW = torch.empty(C, H, device=dev); db = torch.empty(C, device=dev)
for _ in range(10):
    train_step.head_nll(h, W, b, y, dW=dW, db=db, workspace=ws)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(200):
    train_step.head_nll(h, W, b, y, dW=dW, db=db, workspace=ws)
e1.record(); torch.cuda.synchronize()
print("head_nll B=%d: %.2f us per call (incl. host launch + 2 torch.empty)" % (B, 1e3 * e0.elapsed_time(e1) / 200))
g = torch.cuda.CUDAGraph()
with torch.cuda.graph(g):
    for _ in range(20):
        train_step.head_nll(h, W, b, y, dW=dW, db=db, workspace=ws)
g.replay(); torch.cuda.synchronize()
e0.record(); g.replay(); e1.record(); torch.cuda.synchronize()
print("head_nll B=%d in a graph: %.2f us per kernel" % (B, 1e3 * e0.elapsed_time(e1) / 20))
