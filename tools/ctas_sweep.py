import sys, os, json, subprocess
# bench-like sets: run quick timing of mpcq_solve on the bench workload for several CTAs/SM
sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests')
import numpy as np, torch
k=sys.argv[1]
os.environ['MPCQ_CTAS_PER_SM']=k
from helpers import make_batch
from pympc_quadruped_b200 import *
from pympc_quadruped_b200.engine import MpcqEngine
for B,nset in ((4096,16),(65536,2)):
    sets=[make_batch(A1Config,10,B,'mixed',(Gait.TROTTING10,),100+i,solve=False) for i in range(nset)]
    eng=MpcqEngine(sets[0]['cfg'],A1Config)
    t=lambda a,dt=torch.float32: torch.as_tensor(a).to(device='cuda:0',dtype=dt)
    D=[(t(b['x0']),t(b['feet']),t(b['gait']),t(b['xref']),t(b['yaw'])) for b in sets]
    out=None
    for s in range(5):
        x=D[s%nset]; r=eng.solve(x[0],x[1],x[2],x[3],yaw=x[4],want=())
    torch.cuda.synchronize()
    e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
    n=100 if B==4096 else 10
    e0.record()
    for s in range(n):
        x=D[s%nset]; r=eng.solve(x[0],x[1],x[2],x[3],yaw=x[4],want=())
    e1.record(); torch.cuda.synchronize()
    ms=e0.elapsed_time(e1)/n
    print(f"ctas/SM {k} B={B}: {ms:.3f} ms/step -> {B/ms/1e3:.2f} M solves/s")
