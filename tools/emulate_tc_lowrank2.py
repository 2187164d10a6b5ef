"""Developer tool: variants of tools/emulate_tc_lowrank.py (same bit-exact tcgen05 accumulate model) used to pick the
accumulator scheme of the two-stage low-rank kernel (fgrnn_tc_lr.cu):
   order    'inter' : lo.hi, hi.lo, hi.hi per k-step in one chain     'lofirst': all lo products first, then the hi.hi chain
   s1main   number of accumulators the hi.hi chain of stage 1 (16 k-steps of h.U1) is split over
   s2acc    1: one accumulator per stage-2 tile; 2: lo products in their own accumulator
Prints max |h - ref| / (1e-6 + 1e-5 |ref|) against the fp32 oracle / an fp64 evaluation."""
import sys
import numpy as np
import torch
sys.path.insert(0, '/root/repo')
from oracle import fastgrnn_oracle as O
from tools.emulate_tc_schemes import mma, split_fp16, sigmoid32
from tools.emulate_tc_lowrank import scale_of
f32 = np.float32


def chain(B, N, prods_lo, prods_hi, order, nmain, sep_lo):
    """prods_*: lists of (a[B,16], b[16,N]).  Returns fp32 total."""
    accs = []
    if order == 'inter':
        acc = np.zeros((B, N))
        nk = len(prods_hi)
        for i in range(nk):
            acc = mma(acc, *prods_lo[2 * i]); acc = mma(acc, *prods_lo[2 * i + 1]); acc = mma(acc, *prods_hi[i])
        return acc.astype(f32)
    lo = np.zeros((B, N))
    for p in prods_lo: lo = mma(lo, *p)
    mains = [np.zeros((B, N)) for _ in range(nmain)]
    if not sep_lo: mains[0] = lo
    for i, p in enumerate(prods_hi):
        g = i * nmain // len(prods_hi)
        mains[g] = mma(mains[g], *p)
    tot = lo.astype(f32) if sep_lo else np.zeros((B, N), f32)
    for g in range(nmain - 1, -1, -1): tot = (tot + mains[g].astype(f32)).astype(f32)
    return tot


def run(seed, order, s1main, s1sep, s2main, s2sep, T=99, B=8, I=32, H=256, rW=16, rU=32):
    torch.manual_seed(seed)
    p = O.init_params(I, H, rW, rU)
    x = torch.randn(B, T, I)
    ref = O.unroll(x, p, None, True).numpy()
    W1, W2, U1, U2 = (t.numpy() for t in (p.W1, p.W2, p.U1, p.U2))
    bg = p.bias_gate.numpy()[0]; bu = p.bias_update.numpy()[0]
    sz = sigmoid32(np.array(p.zeta.item(), f32)); sn = sigmoid32(np.array(p.nu.item(), f32))
    hd = np.zeros((B, H)); tr = np.zeros((B, T, H))
    for t in range(T):
        pre = (x[:, t].numpy().astype(np.float64) @ W1.astype(np.float64)) @ W2.astype(np.float64) + (hd @ U1.astype(np.float64)) @ U2.astype(np.float64)
        z = 1 / (1 + np.exp(-(pre + bg))); c = np.tanh(pre + bu)
        hd = z * hd + (float(sz) * (1 - z) + float(sn)) * c
        tr[:, t] = hd
    S1 = scale_of(U1, W1); S2 = scale_of(U2, W2)
    U1h, U1l = split_fp16(U1, S1); W1h, W1l = split_fp16(W1, S1)
    A2 = np.concatenate([U2, W2], 0)
    A2h, A2l = split_fp16(A2, S2)
    h = np.zeros((B, H), f32); out = np.zeros((B, T, H), f32); xs = x.numpy()
    for t in range(T):
        hh, hl = split_fp16(h, 0); xh, xl = split_fp16(xs[:, t], 0)
        def stage1(ah, al, bh, bl, K, nmain):
            lo, hi = [], []
            for k in range(0, K, 16):
                lo += [(al[:, k:k+16], bh[k:k+16]), (ah[:, k:k+16], bl[k:k+16])]; hi += [(ah[:, k:k+16], bh[k:k+16])]
            return chain(B, bh.shape[1], lo, hi, order, nmain, s1sep)
        s = (stage1(hh, hl, U1h, U1l, H, s1main) * f32(2.0 ** -S1)).astype(f32)
        sx = (stage1(xh, xl, W1h, W1l, I, 1) * f32(2.0 ** -S1)).astype(f32)
        v = np.concatenate([s, sx], 1)
        vh, vl = split_fp16(v, 0)
        lo, hi = [], []
        for k in range(0, rU + rW, 16):
            lo += [(vl[:, k:k+16], A2h[k:k+16]), (vh[:, k:k+16], A2l[k:k+16])]; hi += [(vh[:, k:k+16], A2h[k:k+16])]
        D = chain(B, H, lo, hi, order, s2main, s2sep)
        pre = (D * f32(2.0 ** -S2)).astype(f32)
        z = sigmoid32((pre + bg).astype(f32)); c = np.tanh((pre + bu).astype(np.float64)).astype(f32)
        g_ = (sz * (f32(1) - z) + sn).astype(f32)
        h = ((z * h).astype(f32) + (g_ * c).astype(f32)).astype(f32)
        out[:, t] = h
    r_or = (np.abs(out.astype(np.float64) - ref) / (1e-6 + 1e-5 * np.abs(ref))).max()
    r_tr = (np.abs(out.astype(np.float64) - tr) / (1e-6 + 1e-5 * np.abs(tr))).max()
    r_ot = (np.abs(ref.astype(np.float64) - tr) / (1e-6 + 1e-5 * np.abs(tr))).max()
    return r_or, r_tr, r_ot


if __name__ == '__main__':
    cfgs = [('inter', 1, 0, 1, 0), ('lofirst', 1, 0, 1, 0), ('lofirst', 2, 0, 1, 0), ('lofirst', 2, 1, 1, 1), ('lofirst', 2, 1, 3, 1)]
    if len(sys.argv) > 1:
        cfgs = [tuple(int(v) if v.isdigit() else v for v in a.split(',')) for a in sys.argv[1:]]
    for c in cfgs:
        for seed in (0, 1, 2):
            print(c, 'seed', seed, 'vs-oracle %.3f  vs-fp64 %.3f  (oracle vs fp64 %.3f)' % run(seed, *c), flush=True)
