import sys, os, math, torch
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests")
import t5_resnet_vqa_b200 as pkg
from util import Caller
C = Caller(pkg)
BF=torch.bfloat16
def rnd(*shape, seed=0, scale=1.0, dtype=torch.float32):
    g = torch.Generator().manual_seed(seed)
    return (torch.randn(*shape, generator=g) * scale).cuda().to(dtype)
for (B,H,Lq,Lk,hd) in [(64,12,32,32,64),(64,8,32,32,96)]:
    D=H*hd
    q,k,v,dO = [rnd(B*L, D, seed=i, scale=0.5, dtype=BF) for i,L in ((1,Lq),(2,Lk),(3,Lk),(4,Lq))]
    bias = rnd(H,Lq,Lk,seed=5)
    ref=None
    for it in range(20):
        out = torch.zeros(B*Lq, D, dtype=BF, device="cuda")
        stats = torch.zeros(B*H*Lq,2,device="cuda")
        C.attn_fwd(B,H,Lq,Lk,hd,q,D,k,D,v,D,out,D,None,bias,None,0.125,0.0,0,None,stats=stats)
        dq,dk,dv=[torch.zeros_like(t) for t in (q,k,v)]
        C.attn_bwd(B,H,Lq,Lk,hd,q,D,k,D,v,D,None,dO,D,dq,D,dk,D,dv,D,None,0.125,0.0,0,None,stats=stats,bias=bias)
        torch.cuda.synchronize()
        cur=(out,stats,dq,dk,dv)
        if ref is None: ref=cur
        else:
            for n,a,b in zip("out stats dq dk dv".split(),ref,cur):
                if not torch.equal(a,b): print("NONDET", hd, it, n, float((a.float()-b.float()).abs().max()))
print("det done")
