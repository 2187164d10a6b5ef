"""Developer tool: run the forward while a second stream hammers HBM / steals SMs, and compare every result bitwise with a
quiet run (looking for a timing window in the tcgen05 kernels)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from kws_b200 import engine
from oracle import fastgrnn_oracle as O

dev = torch.device("cuda:0")
torch.manual_seed(0)
p = O.init_params(32, 128)
params = {k: v.to(dev).contiguous() for k, v in p.tensors().items()}
noise_a = torch.randn(64 << 20, device=dev)
noise_b = torch.empty_like(noise_a)
side = torch.cuda.Stream()
bad = 0
total = 0
for B, save in ((64, True), (2048, True), (2048, False), (8192, False), (5000, True)):
    x = torch.randn(B, 99, 32, device=dev)
    x2 = torch.randn(1000 + 37 * (B % 7), 40, 32, device=dev)
    quiet = engine.forward(x, params, None, layout="IH", batch_first=True, save_for_backward=save)
    torch.cuda.synchronize()
    ref = [t.clone() if t is not None else None for t in quiet[:3]]
    for it in range(int(os.environ.get('STRESS_ITERS', '40'))):
        with torch.cuda.stream(side):
            for _ in range(1 + it % 4):
                noise_b.copy_(noise_a); noise_b.mul_(1.0001)
            if it % 2:                                   # a second recurrence kernel competing for SMs / TMEM
                engine.forward(x2, params, None, layout="IH", batch_first=True)
        if it % 3 == 0:
            torch.cuda._sleep(20000 * (it % 5))
        got = engine.forward(x, params, None, layout="IH", batch_first=True, save_for_backward=save)
        torch.cuda.synchronize()
        total += 1
        for a, b in zip(got[:3], ref):
            if a is not None and not torch.equal(a, b):
                bad += 1
                d = (a - b).abs()
                print("MISMATCH B=%d save=%s it=%d max %.3e count %d" % (B, save, it, float(d.max()), int((d > 0).sum())))
                break
print("timing stress: %d runs, %d mismatches" % (total, bad))
