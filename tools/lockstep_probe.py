"""All robots of the batch identical (every warp runs the same instruction stream, in phase) vs the mixed batch:
separates instruction-fetch contention between warps in different phases from the work itself."""
import sys
sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests')
import numpy as np, torch
from helpers import make_batch
from pympc_quadruped_b200 import *
from pympc_quadruped_b200.engine import MpcqEngine
B=4096
bt=make_batch(A1Config,10,B,'mixed',(Gait.TROTTING10,),5,solve=False)
eng=MpcqEngine(bt['cfg'],A1Config)
t=lambda a,dt=torch.float32: torch.as_tensor(a).to(device='cuda:0',dtype=dt)
X=[t(bt['x0']),t(bt['feet']),t(bt['gait']),t(bt['xref']),t(bt['yaw'])]
r=eng.solve(X[0],X[1],X[2],X[3],yaw=X[4]); torch.cuda.synchronize()
it=r.iters[:,0].cpu().numpy()
def timeit(a,label):
    for _ in range(3): eng.solve(a[0],a[1],a[2],a[3],yaw=a[4],want=())
    torch.cuda.synchronize()
    e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): eng.solve(a[0],a[1],a[2],a[3],yaw=a[4],want=())
    e1.record(); torch.cuda.synchronize()
    ms=e0.elapsed_time(e1)/10
    print(f'{label}: {ms*1e3:8.1f} us/launch  {ms*1e3/B*148:7.2f} us*SM/env')
timeit(X,'mixed batch (rounds mean %.2f)'%it.mean())
for R in (1,2,4,6,8,10):
    idx=np.flatnonzero(it==R)
    if len(idx)==0: continue
    k=int(idx[0])
    a=[x[k:k+1].expand(B,*x.shape[1:]).contiguous() for x in X]
    timeit(a,f'4096 copies of env {k} ({R} rounds)')
# same multiset of rounds as the mixed batch but grouped: sort envs by rounds
order=np.argsort(it,kind='stable')
a=[x[torch.as_tensor(order,device='cuda:0')].contiguous() for x in X]
timeit(a,'mixed batch sorted by rounds (schedule on)')
