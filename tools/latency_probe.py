import sys, time
sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests')
import numpy as np, torch
from helpers import make_batch
from pympc_quadruped_b200 import *
from pympc_quadruped_b200.engine import MpcqEngine
bt=make_batch(A1Config,10,8192,'mixed',(Gait.TROTTING10,),9,solve=False)
eng=MpcqEngine(bt['cfg'],A1Config)
t=lambda a,dt=torch.float32: torch.as_tensor(a).to(device='cuda:0',dtype=dt)
X=[t(bt['x0']),t(bt['feet']),t(bt['gait']),t(bt['xref']),t(bt['yaw'])]
res=eng.solve(X[0],X[1],X[2],X[3],yaw=X[4]); torch.cuda.synchronize()
it=res.iters[:,0].cpu().numpy()
def timeit(idx,reps=20):
    idx=torch.as_tensor(idx,device='cuda:0')
    a=[x[idx].contiguous() for x in X]
    for _ in range(3): eng.solve(a[0],a[1],a[2],a[3],yaw=a[4])
    e0=torch.cuda.Event(enable_timing=True); e1=torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    for _ in range(reps): eng.solve(a[0],a[1],a[2],a[3],yaw=a[4])
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1)/reps
out=[]
for r in (1,4,8,12):
    sel=np.flatnonzero(it==r)
    if len(sel)==0: continue
    ms=timeit(sel[:1]); out.append(f'r{r}: {ms*1e3:6.1f} us ({ms*1e3/r:5.1f}/round)')
print(' | '.join(out))
hard=np.argsort(-it)[:148]
print(f'148 hardest envs (1 per SM): {timeit(hard):.3f} ms, max rounds {it[hard].max()};  B=4096: {timeit(np.arange(4096)):.3f} ms')
