"""Developer tool: in a fresh process, run the forward twice on the same inputs and report where the FIRST launch
differs from the second (a rare first-launch deviation was seen twice in round 1)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from kws_b200 import engine
from oracle import fastgrnn_oracle as O

B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
dev = torch.device("cuda:0")
torch.manual_seed(0)
p = O.init_params(32, 128)
params = {k: v.to(dev).contiguous() for k, v in p.tensors().items()}
x = torch.randn(B, 99, 32).to(dev)
save = len(sys.argv) > 2 and sys.argv[2] == "save"
a = engine.forward(x, params, None, layout="IH", batch_first=True, save_for_backward=save)[0].clone()
torch.cuda.synchronize()
outs = [engine.forward(x, params, None, layout="IH", batch_first=True, save_for_backward=save)[0].clone() for _ in range(3)]
torch.cuda.synchronize()
same_later = all(torch.equal(outs[0], o) for o in outs[1:])
d = (a - outs[0]).abs()
if float(d.max()) == 0.0 and same_later:
    print("OK")
else:
    nz = d.nonzero()
    ts = nz[:, 1]
    print("DIFF max %.3e count %d first_t %d last_t %d rows %d units %d later_equal %s" % (
        float(d.max()), nz.shape[0], int(ts.min()), int(ts.max()), nz[:, 0].unique().numel(), nz[:, 2].unique().numel(), same_later))
    t0 = int(ts.min())
    first = nz[ts == t0]
    print("  at first differing step t=%d: rows %s units(min,max,count) %d %d %d maxdiff %.3e" % (
        t0, sorted(set(first[:, 0].tolist()))[:20], int(first[:, 2].min()), int(first[:, 2].max()), first.shape[0], float(d[:, t0].max())))
