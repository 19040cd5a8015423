import sys, torch
sys.path.insert(0,'/root/repo/multi-degradation-image-enhancement_b200'); sys.path.insert(0,'/root/repo')
import cdan_b200_native as native
import torch.nn.functional as F
dev=torch.device('cuda',0)
which = sys.argv[1] if len(sys.argv) > 1 else 'all'
cases = {'dense': (1,64,16,24,40,3,True,False,False), 'trans': (1,128,64,8,16,1,True,False,False), 'dec4': (2,64,3,8,8,3,False,True,False),
         'c1': (1,3,64,38,52,3,False,True,False), 'c1p': (2,3,64,16,24,3,False,True,True)}
for name,(n,ci,co,h,w,ks,pre,relu,pool) in cases.items():
    if which not in ('all', name): continue
    g=torch.Generator().manual_seed(1)
    x=torch.randn((n,ci,h,w),generator=g); wt=torch.randn((co,ci,ks,ks),generator=g)*0.1; b=torch.randn((co,),generator=g)*0.1
    p=(torch.rand((ci,),generator=g)+0.5, torch.randn((ci,),generator=g)*0.3) if pre else None
    a = x if p is None else F.relu(x*p[0].view(1,-1,1,1)+p[1].view(1,-1,1,1))
    ref=F.conv2d(a,wt,b,padding=ks//2)
    if relu: ref=F.relu(ref)
    if pool: ref=F.max_pool2d(ref,2,2)
    try:
        got=native.op_conv2d(x.to(dev),wt,b,None if p is None else p[0],None if p is None else p[1],relu=relu,pool=pool,dtype='bf16',impl=0).cpu()
        print(name,'maxerr',float((got-ref).abs().max()), flush=True)
    except Exception as e:
        print(name,'FAILED',e, flush=True); break
