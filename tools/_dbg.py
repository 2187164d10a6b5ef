import sys, os
sys.path.insert(0, '/root/repo')
import torch
from kws_b200 import _lib, engine
from oracle import fastgrnn_oracle as O
sys.path.insert(0, '/root/repo/tools')
from precision_check import truth64
torch.manual_seed(0)
p = O.init_params(32, 128)
x = torch.randn(64, 99, 32)
ref = O.unroll(x, p, None, True)
tr = truth64(x, p, None)
params = {k: v.cuda().contiguous() for k, v in p.tensors().items()}
def rr(o, b):
    r = ((o.double() - b.double()).abs() / (1e-6 + 1e-5 * b.double().abs()))
    bad = torch.nonzero(r > 1)
    return float(r.max()), len(bad), (sorted(set(bad[:, 0].tolist())), sorted(set(bad[:, 1].tolist()))[:12], sorted(set(bad[:, 2].tolist()))[:12]) if len(bad) else None
for name, path in (("smem", _lib.PATH_SMEM), ("tc", _lib.PATH_TCGEN05), ("tc", _lib.PATH_TCGEN05)):
    out = engine.forward(x.cuda(), params, None, layout="IH", batch_first=True, force_path=path)[0]
    torch.cuda.synchronize()
    print(name, 'gpu-vs-truth', rr(out, tr))
    oc = out.cpu()
    print(name, 'cpu-vs-ref', rr(oc, ref), 'cpu-vs-truth', rr(oc, tr.cpu()), 'gpu again', rr(out, tr))
