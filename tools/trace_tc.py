import sys, ctypes, torch
sys.path.insert(0, "/root/repo")
from mineral_b200 import ops
from mineral_b200._native import lib
DEV="cuda:0"
M,N,K=32768,128,256
g=torch.Generator(device=DEV).manual_seed(0)
groups=[]
for t in range(2):
    x=torch.randn(M,K,device=DEV,generator=g); w=torch.randn(N,K,device=DEV,generator=g)/16; b=torch.randn(N,device=DEV,generator=g)
    hi=torch.empty(N,K,device=DEV); lo=torch.empty(N,K,device=DEV)
    ops.split_tf32([(w,hi,lo,N,K,K,K)])
    groups.append(dict(a=x,b=hi,b_lo=lo,c=torch.empty(M,N,device=DEV),bias=b,N=N,K=K))
for _ in range(3): ops.tc_linear_fwd(groups,M,1)
torch.cuda.synchronize()
lib.mnr_tc_trace(1,None)
ops.tc_linear_fwd(groups,M,1); torch.cuda.synchronize()
buf=(ctypes.c_longlong*(4*256*2))()
lib.mnr_tc_trace(0,buf)
import numpy as np
tr=np.array(buf[:]).reshape(4,256,2)
t0=tr[0,0,0]
def rel(x): return (x-t0)/1.9  # ~ns at 1.9 GHz
print("kb : prod_issue | conv_start conv_done | mma_start mma_issued   (ns from first TMA issue)")
for kb in range(34):
    print(f"{kb:3d}: {rel(tr[0,kb,0]):8.0f} | {rel(tr[1,kb,0]):8.0f} {rel(tr[1,kb,1]):8.0f} | {rel(tr[2,kb,0]):8.0f} {rel(tr[2,kb,1]):8.0f}")
print("tile: epi_start epi_done")
for it in range(5): print(it, f"{rel(tr[3,it,0]):8.0f} {rel(tr[3,it,1]):8.0f}")
