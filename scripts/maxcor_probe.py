import time, numpy as np, sys
sys.path.insert(0,'/root/repo')
import saver_b200
from saver_b200 import ops
import torch
h=saver_b200.default_handle(0)
for (n,p,B) in ((2000,5000,1024),(10000,20000,1024)):
    x=torch.randn(p,n,dtype=torch.float64,device='cuda')
    ops.set_design(x,handle=h)
    Y=torch.rand(B,n,dtype=torch.float64,device='cuda')
    for it in range(3):
        torch.cuda.synchronize(); h.synchronize(); t=time.time()
        o=ops.maxcor(Y,handle=h); h.synchronize(); dt=time.time()-t
    print(n,p,B,'maxcor ms',dt*1e3,'TFLOP/s',2*n*p*B/dt/1e12)
    ref=(torch.abs(((x-x.mean(1,keepdim=True))/x.std(1,unbiased=False,keepdim=True))@((Y-Y.mean(1,keepdim=True))/Y.std(1,unbiased=False,keepdim=True)).T)/n).max(0).values
    print('maxdiff',(o-ref).abs().max().item())
    # cublas dgemm for comparison
    a=torch.randn(p,n,dtype=torch.float64,device='cuda'); b=torch.randn(n,B,dtype=torch.float64,device='cuda')
    for it in range(3):
        torch.cuda.synchronize(); t=time.time(); c=a@b; torch.cuda.synchronize(); dt=time.time()-t
    print('cublas dgemm TFLOP/s',2*n*p*B/dt/1e12)
