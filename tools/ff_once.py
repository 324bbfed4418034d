import sys, torch
sys.path.insert(0,'/root/repo')
import vdpp_b200
from vdpp_b200 import native
from vdpp_b200.models.native_unet import interleave_geglu
M,C=230400,320; inner=4*C
h=lambda *s,scale=1.0:(torch.randn(*s,device='cuda')*scale).half()
x=h(M,C,scale=0.5); w1=h(2*inner,C,scale=C**-0.5); b1=h(2*inner,scale=0.1); w2=h(C,inner,scale=inner**-0.5); b2=h(C,scale=0.1); r1=h(M,C)
o=torch.empty(M,C,device='cuda',dtype=torch.float16)
w1i,b1i,_=interleave_geglu(w1,b1,half=64)
for _ in range(3): native.ff_geglu(o,x,w1i,b1i,w2,b2,r1=r1)
torch.cuda.synchronize()
print('done')
