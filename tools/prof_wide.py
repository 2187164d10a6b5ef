"""Developer tool: a few forward launches of one default-model layer shape (for ncu)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from kws_b200 import engine
from oracle import fastgrnn_oracle as O
I, H, B = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
dev = torch.device("cuda:0")
torch.manual_seed(0)
p = O.init_params(I, H)
params = {k: v.to(dev).contiguous() for k, v in p.tensors().items()}
x = torch.randn(B, 99, I, device=dev)
out = torch.empty(B, 99, H, device=dev)
for _ in range(3):
    engine.forward(x, params, None, layout="IH", batch_first=True, out=out)
torch.cuda.synchronize()
print("ok")
