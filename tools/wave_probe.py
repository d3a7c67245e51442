"""Time of one launch for batch sizes around one resident wave (is the kernel bound by per-warp latency or by a per-SM resource?)."""
import sys
sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests')
import numpy as np, torch
from helpers import make_batch
from pympc_quadruped_b200 import *
from pympc_quadruped_b200.engine import MpcqEngine
bt=make_batch(A1Config,10,8192,'mixed',(Gait.TROTTING10,),5,solve=False)
eng=MpcqEngine(bt['cfg'],A1Config)
t=lambda a,dt=torch.float32: torch.as_tensor(a).to(device='cuda:0',dtype=dt)
X=[t(bt['x0']),t(bt['feet']),t(bt['gait']),t(bt['xref']),t(bt['yaw'])]
for B in [int(a) for a in sys.argv[1:]] or (148,296,592,1036,1184,2072,4096,8192):
    a=[x[:B].contiguous() for x in X]
    for _ in range(3): r=eng.solve(a[0],a[1],a[2],a[3],yaw=a[4],want=())
    torch.cuda.synchronize()
    e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20): r=eng.solve(a[0],a[1],a[2],a[3],yaw=a[4],want=())
    e1.record(); torch.cuda.synchronize()
    ms=e0.elapsed_time(e1)/20
    print(f'B={B:5d}: {ms*1e3:8.1f} us/launch  {ms*1e3/B*148:7.2f} us*SM/env  {B/ms/1e3:6.2f} M solves/s')
