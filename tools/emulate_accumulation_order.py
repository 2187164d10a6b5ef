import numpy as np, torch, sys
sys.path.insert(0,'/root/repo')
from oracle import fastgrnn_oracle as O
torch.manual_seed(0)
p = O.init_params(32,128)
x = torch.randn(64,99,32)
ref = O.unroll(x,p,None,True).numpy()
W=p.W.numpy().astype(np.float64); U=p.U.numpy().astype(np.float64)
bg=p.bias_gate.numpy()[0]; bu=p.bias_update.numpy()[0]
sz=np.float32(1/(1+np.exp(-np.float64(p.zeta.item())))); sn=np.float32(1/(1+np.exp(-np.float64(p.nu.item()))))
xs=x.numpy()
f32=np.float32
def seq(acc, A, Wm):   # sequential fma chain over k
    for k in range(A.shape[1]):
        acc = (acc.astype(np.float64) + A[:,k:k+1].astype(np.float64)*Wm[k:k+1,:]).astype(f32)
    return acc
def run(mode):
    h=np.zeros((64,128),f32); out=np.zeros((64,99,128),f32)
    for t in range(99):
        xt=xs[:,t,:]
        z0=np.zeros((64,128),f32)
        if mode=='x_first': pre=seq(seq(z0,xt,W),h,U)
        elif mode=='h_first': pre=seq(seq(z0,h,U),xt,W)
        elif mode=='separate': pre=(seq(z0,xt,W)+seq(z0,h,U)).astype(f32)
        elif mode=='h_2chains':   # two chains over even/odd 4-blocks of k for the h part, x added last
            a=z0.copy(); b=z0.copy()
            for kb in range(0,128,8):
                a=seq(a,h[:,kb:kb+4],U[kb:kb+4]); b=seq(b,h[:,kb+4:kb+8],U[kb+4:kb+8])
            pre=seq((a+b).astype(f32),xt,W)
        elif mode=='exact': pre=(xt.astype(np.float64)@W+h.astype(np.float64)@U).astype(f32)
        elif mode=='mma8':  # tensor-core like: exact sum of 8 products, one rounding per 8 (fp32 accumulate)
            acc=z0.astype(np.float64)
            A=np.concatenate([xt,h],1).astype(np.float64); M=np.concatenate([W,U],0)
            for kb in range(0,160,8):
                acc=(acc+A[:,kb:kb+8]@M[kb:kb+8]).astype(f32).astype(np.float64)
            pre=acc.astype(f32)
        a1=(pre+bg).astype(f32); a2=(pre+bu).astype(f32)
        z=(1/(1+np.exp(-a1.astype(np.float64)))).astype(f32); c=np.tanh(a2.astype(np.float64)).astype(f32)
        g=(sz*(f32(1)-z)+sn).astype(f32)
        h=((z*h).astype(f32)+(g*c).astype(f32)).astype(f32)
        out[:,t]=h
    r=np.abs(out.astype(np.float64)-ref)/(1e-6+1e-5*np.abs(ref))
    return r.max()
for m in ['exact','separate','x_first','h_first','h_2chains','mma8']:
    print(m, round(run(m),3))
