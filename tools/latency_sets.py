"""Per-input-set latency of one mpcq_solve launch (4 096 robots, the bench's 64 seeded sets): p50 / p90 / max and the rounds of the
slowest sets, for solver knobs given as key=value (e.g. max_pdas_rounds=10)."""
import sys
sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests')
import numpy as np, torch
from helpers import make_batch
from pympc_quadruped_b200 import *
from pympc_quadruped_b200.engine import MpcqEngine
kn={}
for a in sys.argv[1:]:
    k,v=a.split('='); kn[k]=float(v) if ('.' in v or 'e' in v) else int(v)
S,B=32,4096
t=lambda a,dt=torch.float32: torch.as_tensor(a).to(device='cuda:0',dtype=dt)
sets=[]
for s in range(S):
    bt=make_batch(A1Config,10,B,'mixed',(Gait.TROTTING10,),500+s,solve=False)
    sets.append([t(bt['x0']),t(bt['feet']),t(bt['gait']),t(bt['xref']),t(bt['yaw'])])
eng=MpcqEngine(bt['cfg'],A1Config,**kn)
res=eng.solve(*sets[0][:4],yaw=sets[0][4],want=('iters','status'))
lat=[]; info=[]
for rep in range(3):
    for s in range(S):
        X=sets[s]
        e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(); e0.record(); eng.solve(X[0],X[1],X[2],X[3],yaw=X[4],out=res); e1.record(); torch.cuda.synchronize()
        if rep==2:
            lat.append(e0.elapsed_time(e1)); it=res.iters.cpu().numpy(); st=res.status.cpu().numpy()
            info.append((it[:,0].max(), it[:,1].max(), int((st&2).astype(bool).sum()), float(it[:,0].mean())))
lat=np.array(lat); o=np.argsort(lat)
print(kn, f'p50 {np.median(lat):.3f} p90 {np.percentile(lat,90):.3f} max {lat.max():.3f} mean {lat.mean():.3f}  p90/p50 {np.percentile(lat,90)/np.median(lat):.3f}')
for s in o[-4:]: print('   set',s,f'{lat[s]:.3f} ms  rounds max {info[s][0]} AS max {info[s][1]} fallback envs {info[s][2]} rounds mean {info[s][3]:.2f}')
