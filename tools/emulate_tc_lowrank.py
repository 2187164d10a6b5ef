"""Developer tool: bit-exact emulation (tcgen05.mma kind::f16 accumulate model of tools/emulate_tc_schemes.py) of the two-stage
low-rank recurrence planned for C4 (H = 256, wRank 16, uRank 32):
   stage 1: [s | sx] = [h.U1 | x.W1]   lo products -> C1, hi.hi -> M1 (one chain each: TMEM has no room for more)
   stage 2: pre = [s | sx].[U2 ; W2]    one accumulator per 128-unit tile, 3 k-steps x 3 products
Prints max |h - ref| / (1e-6 + 1e-5 |ref|) against the fp32 oracle and against an fp64 evaluation."""
import sys
import numpy as np
import torch
sys.path.insert(0, '/root/repo')
from oracle import fastgrnn_oracle as O
from tools.emulate_tc_schemes import mma, split_fp16, sigmoid32
f32 = np.float32


def scale_of(*mats):
    mx = max(float(np.abs(m).max()) for m in mats)
    return int(np.floor(np.log2(30000.0 / mx)))


def run(seed, T=99, B=8, I=32, H=256, rW=16, rU=32, split_stage1=1):
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
    A2 = np.concatenate([U2, W2], 0)                       # [rU + rW, H]
    A2h, A2l = split_fp16(A2, S2)
    h = np.zeros((B, H), f32); out = np.zeros((B, T, H), f32); xs = x.numpy()
    for t in range(T):
        hh, hl = split_fp16(h, 0); xh, xl = split_fp16(xs[:, t], 0)
        # stage 1 (s: [B, rU], sx: [B, rW]); chains split in `split_stage1` pieces
        def stage1(ah, al, bh, bl, K):
            nk = K // 16
            Cs = [np.zeros((B, bh.shape[1])) for _ in range(split_stage1)]; Ms = [np.zeros((B, bh.shape[1])) for _ in range(split_stage1)]
            for i, k in enumerate(range(0, K, 16)):
                g = i * split_stage1 // nk
                Cs[g] = mma(Cs[g], al[:, k:k+16], bh[k:k+16]); Cs[g] = mma(Cs[g], ah[:, k:k+16], bl[k:k+16])
                Ms[g] = mma(Ms[g], ah[:, k:k+16], bh[k:k+16])
            tot = np.zeros((B, bh.shape[1]), f32)
            for g in range(split_stage1): tot = (tot + (Cs[g].astype(f32) + Ms[g].astype(f32)).astype(f32)).astype(f32)
            return tot
        # the kernel accumulates the h part and the x part into the same accumulators (different lanes): separate sums here
        s = (stage1(hh, hl, U1h, U1l, H) * f32(2.0 ** -S1)).astype(f32)
        sx = (stage1(xh, xl, W1h, W1l, I) * f32(2.0 ** -S1)).astype(f32)
        v = np.concatenate([s, sx], 1)                     # [B, rU + rW]
        vh, vl = split_fp16(v, 0)
        D = np.zeros((B, H))
        for k in range(0, rU + rW, 16):
            D = mma(D, vl[:, k:k+16], A2h[k:k+16]); D = mma(D, vh[:, k:k+16], A2l[k:k+16]); D = mma(D, vh[:, k:k+16], A2h[k:k+16])
        pre = (D.astype(f32) * f32(2.0 ** -S2)).astype(f32)
        z = sigmoid32((pre + bg).astype(f32)); c = np.tanh((pre + bu).astype(np.float64)).astype(f32)
        g_ = (sz * (f32(1) - z) + sn).astype(f32)
        h = ((z * h).astype(f32) + (g_ * c).astype(f32)).astype(f32)
        out[:, t] = h
    r_or = (np.abs(out.astype(np.float64) - ref) / (1e-6 + 1e-5 * np.abs(ref))).max()
    r_tr = (np.abs(out.astype(np.float64) - tr) / (1e-6 + 1e-5 * np.abs(tr))).max()
    r_ot = (np.abs(ref.astype(np.float64) - tr) / (1e-6 + 1e-5 * np.abs(tr))).max()
    return r_or, r_tr, r_ot


if __name__ == '__main__':
    for sp in (1, 2):
        for seed in (0, 1):
            print('stage-1 chains split in %d: seed %d  vs-oracle %.3f  vs-fp64 %.3f  (oracle vs fp64 %.3f)' % ((sp, seed) + run(seed, split_stage1=sp)), flush=True)
