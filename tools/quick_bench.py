import sys, time
sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests')
import numpy as np, torch
from helpers import make_batch
from pympc_quadruped_b200 import *
from pympc_quadruped_b200.engine import MpcqEngine
from pympc_quadruped_b200.synth import GAIT_MIX
def bench(robot,H,B,regime,gaits,dtype,seed=1,reps=20,**knobs):
    bt=make_batch(robot,H,B,regime,gaits,seed,solve=False)
    eng=MpcqEngine(bt['cfg'],robot,dtype=dtype,**knobs)
    t=lambda a,dt: torch.as_tensor(a).to(device='cuda:0',dtype=dt)
    x0,feet,gait,xref,yaw=t(bt['x0'],dtype),t(bt['feet'],dtype),t(bt['gait'],torch.float32),t(bt['xref'],dtype),t(bt['yaw'],dtype)
    res=eng.solve(x0,feet,gait,xref,yaw=yaw); torch.cuda.synchronize()
    it=res.iters.cpu().numpy(); st=res.status.cpu().numpy()
    e0=torch.cuda.Event(enable_timing=True); e1=torch.cuda.Event(enable_timing=True)
    for _ in range(3): eng.solve(x0,feet,gait,xref,yaw=yaw,out=res)
    torch.cuda.synchronize(); e0.record()
    for _ in range(reps): eng.solve(x0,feet,gait,xref,yaw=yaw,out=res)
    e1.record(); torch.cuda.synchronize()
    ms=e0.elapsed_time(e1)/reps
    print(knobs, f'{robot.__name__} H={H} B={B} {regime} {dtype}: {ms:.3f} ms/step -> {B/ms*1e3:.0f} solves/s | facts mean {it[:,0].mean():.2f} max {it[:,0].max()} fallback {(st&2).astype(bool).sum()} unverified {(~(st&1).astype(bool)).sum()}',flush=True)
bench(A1Config,10,4096,'mixed',(Gait.TROTTING10,),torch.float32)
bench(A1Config,10,4096,'nominal',(Gait.TROTTING10,),torch.float32)
bench(A1Config,10,4096,'aggressive',(Gait.TROTTING10,),torch.float32)
bench(A1Config,10,4096,'mixed',(Gait.TROTTING10,),torch.float64)
bench(AliengoConfig,10,16384,'mixed',GAIT_MIX,torch.float64)
bench(A1Config,10,4096,'mixed',(Gait.STANDING,),torch.float32)
bench(A1Config,30,4096,'mixed',(Gait.TROTTING10,),torch.float32,reps=5)
bench(A1Config,10,65536,'mixed',(Gait.TROTTING10,),torch.float32,reps=5)
