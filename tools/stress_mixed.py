"""Developer tool: repeat the full-size property test sequence (deterministic, chunked carry, batch slices) many
times in one process, and again across fresh processes, to flush out launch-order dependent races."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import torch
from kws_b200 import rnn
from oracle import fastgrnn_oracle as O
from gpu_helpers import load_cell_params

n = int(sys.argv[1]) if len(sys.argv) > 1 else 30
torch.manual_seed(0)
p = O.init_params(32, 128)
m = rnn.FastGRNN(32, 128, batch_first=True)
load_cell_params(m.cell, p.tensors(), False)
m = m.cuda()
x = torch.randn(8192, 99, 32).cuda()
bad = 0
with torch.no_grad():
    for i in range(n):
        full = m(x)
        again = m(x)
        h = torch.zeros(1, 8192, 128, device="cuda")
        a = m(x[:, :50].contiguous(), h)
        b = m(x[:, 50:].contiguous(), h)
        lo = m(x[:4099]); hi = m(x[4099:])
        ok = torch.equal(full, again) and torch.equal(torch.cat([a, b], 1), full) and torch.equal(torch.cat([lo, hi], 0), full)
        if not ok:
            bad += 1
            print("iteration", i, "mismatch:", torch.equal(full, again), torch.equal(torch.cat([a, b], 1), full), torch.equal(torch.cat([lo, hi], 0), full), flush=True)
print("mixed sequence: %d iterations, %d mismatches" % (n, bad))
