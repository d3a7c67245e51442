"""Device-resident solves/s of the larger size classes (standing n = 120, H = 30 trot n = 180, H = 30 standing n = 360)."""
import sys
sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests')
import numpy as np, torch
from helpers import make_batch
from pympc_quadruped_b200 import *
from pympc_quadruped_b200.engine import MpcqEngine
def bench(robot,H,B,regime,gaits,dtype,seed=1,reps=5,**knobs):
    bt=make_batch(robot,H,B,regime,gaits,seed,solve=False)
    eng=MpcqEngine(bt['cfg'],robot,dtype=dtype,**knobs)
    t=lambda a,dt: torch.as_tensor(a).to(device='cuda:0',dtype=dt)
    x0,feet,gait,xref,yaw=t(bt['x0'],dtype),t(bt['feet'],dtype),t(bt['gait'],torch.float32),t(bt['xref'],dtype),t(bt['yaw'],dtype)
    res=eng.solve(x0,feet,gait,xref,yaw=yaw); torch.cuda.synchronize()
    it=res.iters.cpu().numpy(); st=res.status.cpu().numpy()
    e0=torch.cuda.Event(enable_timing=True); e1=torch.cuda.Event(enable_timing=True)
    for _ in range(2): eng.solve(x0,feet,gait,xref,yaw=yaw,out=res)
    torch.cuda.synchronize(); e0.record()
    for _ in range(reps): eng.solve(x0,feet,gait,xref,yaw=yaw,out=res)
    e1.record(); torch.cuda.synchronize()
    ms=e0.elapsed_time(e1)/reps
    print(f'{robot.__name__} H={H} B={B} {regime} {gaits[0].name if hasattr(gaits[0],"name") else ""} {dtype}: {ms:.3f} ms/step -> {B/ms*1e3:.0f} solves/s | rounds mean {it[:,0].mean():.2f} max {it[:,0].max()} fallback {(st&2).astype(bool).sum()} unverified {(~(st&1).astype(bool)).sum()}',flush=True)
which=sys.argv[1:] or ['stand10','trot30','stand30','trot30f64']
if 'stand10' in which: bench(A1Config,10,4096,'mixed',(Gait.STANDING,),torch.float32)
if 'trot30' in which: bench(A1Config,30,4096,'mixed',(Gait.TROTTING10,),torch.float32)
if 'stand30' in which: bench(A1Config,30,1024,'mixed',(Gait.STANDING,),torch.float32,reps=2)
if 'trot30f64' in which: bench(A1Config,30,1024,'mixed',(Gait.TROTTING10,),torch.float64,reps=2)
if 'trot16' in which: bench(A1Config,16,4096,'mixed',(Gait.TROTTING16,),torch.float32)
if 'pace10' in which: bench(A1Config,10,4096,'mixed',(Gait.PACING10,),torch.float32)
