import sys, os, ctypes
sys.path.insert(0, "/root/repo")
import numpy as np, torch
from pde_opt_b200 import Domain
from pde_opt_b200.adjoint import _fwd, _bwd
from pde_opt_b200.equations import AdvectionDiffusion2D
from pde_opt_b200.functions import GaussianVelocity
N,H,B,K=128,0.02,512,500
dom=Domain((N,N),((-N*H/2,N*H/2),)*2,"d")
eq=AdvectionDiffusion2D(dom,GaussianVelocity(0.1,0.01),0.1)
y0=torch.from_numpy((0.5+0.01*np.random.default_rng(0).normal(size=(B,N,N))).astype(np.float32)).cuda()
ctrl=torch.tensor([0.1,-0.1,0.1,0.01],device="cuda").expand(B,10,4).contiguous()
dts=np.full(K,1e-4,np.float32)
tables=eq.tables_on(y0.device,1.0); desc=eq.ad_desc()
traj=torch.empty((K,B,N,N),device="cuda"); y1=torch.empty_like(y0); lam=torch.randn_like(y0); g=torch.zeros_like(ctrl)
def t(fn,n=3):
    fn(); torch.cuda.synchronize()
    e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize(); return e0.elapsed_time(e1)/n*1e-3
a=t(lambda:_fwd(desc,y0,y1,dts,tables,ctrl,50,0,None))
b=t(lambda:_fwd(desc,y0,y1,dts,tables,ctrl,50,0,traj))
c=t(lambda:_bwd(desc,traj,lam,dts,tables,ctrl,50,0,g))
print(f"fwd {B*K/a/1e6:.2f} M  fwd+traj {B*K/b/1e6:.2f} M  bwd {B*K/c/1e6:.2f} M env-steps/s")
