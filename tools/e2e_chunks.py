"""mpcq_solve_host at B = 4096 with 1-4 chunks (MPCQ_HOST_CHUNKS, experiments only): wall time per call, pinned buffers."""
import sys, time, os
sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests')
import numpy as np, torch
from helpers import make_batch
from pympc_quadruped_b200 import *
from pympc_quadruped_b200.engine import MpcqEngine
B=4096
sets=[make_batch(A1Config,10,B,'mixed',(Gait.TROTTING10,),100+i,solve=False) for i in range(8)]
pin=lambda a,dt: torch.empty(a.shape,dtype=dt,pin_memory=True).copy_(torch.as_tensor(a).to(dt)).numpy()
P=[(pin(b['x0'],torch.float32),pin(b['feet'],torch.float32),pin(b['gait'],torch.float32),pin(b['xref'],torch.float32),pin(b['yaw'],torch.float32)) for b in sets]
out={"forces":torch.empty((B,12),dtype=torch.float32,pin_memory=True).numpy(),"status":torch.empty((B,),dtype=torch.int32,pin_memory=True).numpy()}
for nch, direct in ((1,1),(2,1),(4,1),(1,0),(2,0)):
    os.environ['MPCQ_HOST_CHUNKS']=str(nch); os.environ['MPCQ_HOST_DIRECT']=str(direct)
    eng=MpcqEngine(sets[0]['cfg'],A1Config)
    for s in range(5):
        x=P[s%8]; eng.solve_host(x[0],x[1],x[2],x[3],yaw=x[4],out=out)
    torch.cuda.synchronize(); t0=time.perf_counter()
    for s in range(100):
        x=P[s%8]; eng.solve_host(x[0],x[1],x[2],x[3],yaw=x[4],out=out)
    dt=(time.perf_counter()-t0)/100
    print(f'chunks {nch} direct results {direct}: {dt*1e3:.3f} ms/step -> {B/dt/1e6:.2f} M solves/s')
