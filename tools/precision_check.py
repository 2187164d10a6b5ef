"""Developer tool: error of each forward kernel family against an fp64 evaluation of the same
recurrence (torch, on the GPU) and against the fp32 CPU oracle, in units of the north-star
tolerance (rtol 1e-5, atol 1e-6).  Usage: python tools/precision_check.py [B T seeds]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from kws_b200 import _lib, engine
from oracle import fastgrnn_oracle as O


def truth64(x, p, h0):
    W, U = p.W.double().cuda(), p.U.double().cuda()
    bg, bu = p.bias_gate.double().cuda(), p.bias_update.double().cuda()
    sz, sn = torch.sigmoid(p.zeta.double().cuda()), torch.sigmoid(p.nu.double().cuda())
    x = x.double().cuda()
    h = torch.zeros(x.shape[0], U.shape[0], dtype=torch.float64, device="cuda") if h0 is None else h0.double().cuda()
    outs = []
    for t in range(x.shape[1]):
        pre = x[:, t] @ W + h @ U
        z = torch.sigmoid(pre + bg); c = torch.tanh(pre + bu)
        h = z * h + (sz * (1 - z) + sn) * c
        outs.append(h)
    return torch.stack(outs, 1)


def ratio(a, b):
    a = a.double(); b = b.double()
    return float(((a - b).abs() / (1e-6 + 1e-5 * b.abs())).max())


def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
    T = int(sys.argv[2]) if len(sys.argv) > 2 else 99
    seeds = int(sys.argv[3]) if len(sys.argv) > 3 else 3
    paths = [("smem", _lib.PATH_SMEM), ("tcgen05", _lib.PATH_TCGEN05)]
    for seed in range(seeds):
        for h0_given in (False, True):
            torch.manual_seed(seed)
            p = O.init_params(32, 128)
            x = torch.randn(B, T, 32)
            h0 = 0.5 * torch.randn(B, 128) if h0_given else None
            ref = O.unroll(x, p, None if h0 is None else h0.clone().unsqueeze(0), True)
            tr = truth64(x, p, h0)
            params = {k: v.cuda().contiguous() for k, v in p.tensors().items()}
            line = "seed %d h0 %d | oracle-vs-truth %.3f" % (seed, h0_given, ratio(ref.cuda(), tr))
            for name, path in paths:
                out, _, _, _ = engine.forward(x.cuda(), params, None if h0 is None else h0.cuda(), layout="IH",
                                              batch_first=True, force_path=path)
                torch.cuda.synchronize()
                line += " | %s: vs-truth %.3f vs-oracle %.3f" % (name, ratio(out, tr), ratio(out.cpu(), ref))
                # where does the error sit: first step vs last step
                line += " (t0 %.3f, tlast %.3f)" % (ratio(out[:, 0], tr[:, 0]), ratio(out[:, -1], tr[:, -1]))
            print(line, flush=True)


if __name__ == "__main__":
    main()
