"""Developer tool: run the default (tcgen05) forward and backward many times on the same inputs and check
that every run is bit-identical to the first (no races, no atomics)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from kws_b200 import engine
from oracle import fastgrnn_oracle as O

n = int(sys.argv[1]) if len(sys.argv) > 1 else 200
torch.manual_seed(0)
for (B, T) in ((8192, 99), (2048, 99), (333, 40)):
    p = O.init_params(32, 128)
    params = {k: v.cuda().contiguous() for k, v in p.tensors().items()}
    x = torch.randn(B, T, 32, device="cuda")
    go = torch.randn(B, T, 128, device="cuda") / B
    h0 = 0.3 * torch.randn(B, 128, device="cuda")
    out0, z0, c0, _ = engine.forward(x, params, h0, layout="IH", batch_first=True, save_for_backward=True)
    g0 = engine.backward(go, x, out0, z0, c0, params, h0, layout="IH", batch_first=True)
    torch.cuda.synchronize()
    bad_f = bad_b = 0
    for i in range(n):
        out, z, c, _ = engine.forward(x, params, h0, layout="IH", batch_first=True, save_for_backward=True)
        g = engine.backward(go, x, out, z, c, params, h0, layout="IH", batch_first=True)
        if not (torch.equal(out, out0) and torch.equal(z, z0) and torch.equal(c, c0)):
            bad_f += 1
        if not all(torch.equal(g[k], g0[k]) for k in g0):
            bad_b += 1
    torch.cuda.synchronize()
    print("B=%d T=%d: %d runs, forward mismatches %d, backward mismatches %d" % (B, T, n, bad_f, bad_b), flush=True)
