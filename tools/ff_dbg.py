"""Where does the fused feed-forward kernel (svdpp_ff_geglu_f16, csrc/ff_fused.cu) spend its time?  Level-0 shape
(230400 x 320), boost clocks, with parts of the kernel switched off through the "ff_dbg" tuning bits (results are wrong
then): 1 = no GELU arithmetic, 2 = no GEMM2, 4 = no final epilogue (residual loads, stores), 7 = all three.
   python tools/ff_dbg.py [path/to/libsvdpp_variant.so]"""
import sys, torch
sys.path.insert(0,'/root/repo')
import vdpp_b200
from vdpp_b200 import native
from vdpp_b200.models.native_unet import interleave_geglu
if len(sys.argv) > 1:
    native.use_library(sys.argv[1])
    print('library', sys.argv[1])
M,C=230400,320; inner=4*C
h=lambda *s,scale=1.0:(torch.randn(*s,device='cuda')*scale).half()
x=h(M,C,scale=0.5); w1=h(2*inner,C,scale=C**-0.5); b1=h(2*inner,scale=0.1); w2=h(C,inner,scale=inner**-0.5); b2=h(C,scale=0.1); r1=h(M,C)
o=torch.empty(M,C,device='cuda',dtype=torch.float16)
w1i,b1i,_=interleave_geglu(w1,b1,half=64)
def t(n=10):
    for _ in range(3): native.ff_geglu(o,x,w1i,b1i,w2,b2,r1=r1)
    torch.cuda.synchronize()
    e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): native.ff_geglu(o,x,w1i,b1i,w2,b2,r1=r1)
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1)/n
for pair in (1,2):
    native.set_tuning('ff_pair',pair)
    for dbg in (0,1,2,4,7):
        native.set_tuning('ff_dbg',dbg)
        print('pair',pair,'dbg',dbg,'ms',round(t(),4),flush=True)
native.set_tuning('ff_dbg',0)
