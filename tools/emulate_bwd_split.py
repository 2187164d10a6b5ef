import torch, sys
sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests')
from oracle import fastgrnn_oracle as O
def bf(v): return v.bfloat16().float()
def tf32(v):
    i = v.view(torch.int32); i = (i + 0x1000) & ~0x1fff; return i.view(torch.float32)   # RN-ish to 10-bit mantissa
def split(v, kind):
    if kind=='bf16': hi=bf(v); lo=bf(v-hi); return hi,lo
    if kind=='tf32': hi=tf32(v.clone()); lo=tf32((v-hi).clone()); return hi,lo
    return v, torch.zeros_like(v)
def run(B,T,I,seed,kind):
    torch.manual_seed(seed)
    p = O.init_params(I,128)
    p.bias_gate.add_(0.2*torch.randn(1,128)); p.zeta.add_(0.3); p.nu.add_(0.5)
    x=torch.randn(B,T,I); h0=0.5*torch.randn(B,128); go=torch.randn(B,T,128)/B
    p64=p.map(lambda v:v.double())
    g64=O.autograd_grads(x.double(),p64,h0.double(),go.double(),True)
    # forward in fp64 for z,c,h (isolate backward error)
    W=p.W.double(); U=p.U.double(); sz=torch.sigmoid(p.zeta.double()); sn=torch.sigmoid(p.nu.double())
    h=h0.double(); hs=[h]; zs=[]; cs=[]
    for t in range(T):
        pre=x[:,t].double()@W + h@U
        z=torch.sigmoid(pre+p.bias_gate.double()); c=torch.tanh(pre+p.bias_update.double())
        h=z*h+(sz*(1-z)+sn)*c; hs.append(h); zs.append(z); cs.append(c)
    Uf=p.U.float(); Uhi,Ulo=split(Uf,kind)
    delta=torch.zeros(B,128); dnu=0.0; dze=0.0
    szf=float(sz); snf=float(sn)
    for t in reversed(range(T)):
        z=zs[t].float(); c=cs[t].float(); hp=hs[t].float()
        G=go[:,t]+delta
        dc=(szf*(1-z)+snf)*(1-c*c)*G; dz=(hp-szf*c)*z*(1-z)*G; dpre=dc+dz
        dhi,dlo=split(dpre,kind)
        if kind=='bf16x3':
            dh=bf(dpre); dm=bf(dpre-dh); dl=bf(dpre-dh-dm)
            uh=bf(Uf); um=bf(Uf-uh); ul=bf(Uf-uh-um)
            D=lambda a,b:(a.double()@b.double().t())
            mm=(D(dh,uh)+D(dm,um)+D(dh,um)+D(dm,uh)+D(dh,ul)+D(dl,uh)).float()
        elif kind=='exact': mm=dpre@Uf.t()
        else: mm=(dhi.double()@Uhi.double().t()+dlo.double()@Uhi.double().t()+dhi.double()@Ulo.double().t()).float()
        delta=z*G+mm
        dnu+=float((c*G).double().sum()); dze+=float(((1-z)*c*G).double().sum())
    dnu*=float(sn*(1-sn)); dze*=float(sz*(1-sz))
    r=lambda got,ref: abs(got-float(ref))/(2e-4*abs(float(ref)))
    return r(dnu,g64['nu']), r(dze,g64['zeta'])
for kind in ('bf16','bf16x3'):
    print(kind, [tuple(round(v,3) for v in run(77,9,32,s,kind)) for s in (86,1,2,3,4)])
