import sys, torch
sys.path.insert(0,'/root/repo/multi-degradation-image-enhancement_b200'); sys.path.insert(0,'/root/repo')
import cdan_b200_native as native
import torch.nn.functional as F
dev=torch.device('cuda',0)
n,ci,co,h,w = [int(v) for v in sys.argv[1:6]]
g=torch.Generator().manual_seed(1)
x=torch.randn((n,ci,h,w),generator=g); wt=torch.randn((co,ci,3,3),generator=g)*0.05; b=torch.randn((co,),generator=g)*0.1
p=(torch.rand((ci,),generator=g)+0.5, torch.randn((ci,),generator=g)*0.3)
a = F.relu(x*p[0].view(1,-1,1,1)+p[1].view(1,-1,1,1))
ref=F.conv2d(a,wt,b,padding=1)
got=native.op_conv2d(x.to(dev),wt,b,p[0],p[1],relu=False,pool=False,dtype='bf16',impl=0).cpu()
print(sys.argv[1:6],'maxerr',float((got-ref).abs().max()), flush=True)
