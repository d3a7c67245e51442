"""Fixed cost of the host entry points: wall time per call of mpcq_tick_host / mpcq_solve_host / mpcq_solve (+ sync) for small and full batches."""
import sys, time
sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests')
import numpy as np, torch
from helpers import make_batch
from pympc_quadruped_b200 import *
from pympc_quadruped_b200.engine import MpcqEngine
from pympc_quadruped_b200.synth import synth_states, synth_gait_params
for B in (64, 1024, 4096):
    bt=make_batch(A1Config,10,B,'mixed',(Gait.TROTTING10,),5,solve=False)
    eng=MpcqEngine(bt['cfg'],A1Config)
    st=synth_states(B,A1Config,'mixed',seed=5); offs,durs,segs,it0=synth_gait_params(B,(Gait.TROTTING10,),seed=5)
    sc=np.zeros((B,29)); sc[:,0:4],sc[:,4:7],sc[:,7:10],sc[:,10:13]=st['quat_base'],st['pos_base'],st['ang_vel_base'],st['lin_vel_base']
    sc[:,13:25]=st['pos_base_feet'].reshape(B,12); sc[:,25:28],sc[:,28]=st['vel_cmd_body'],st['yaw_rate_cmd']
    gp=np.concatenate([offs,durs,segs[:,None],(it0*20)[:,None]],axis=1).astype(np.int32)
    pin=lambda a: torch.empty(a.shape,dtype=torch.as_tensor(a).dtype,pin_memory=True).copy_(torch.as_tensor(a)).numpy()
    hsc,hgp=pin(sc),pin(gp)
    out=dict(forces=torch.empty((B,12),dtype=torch.float32,pin_memory=True).numpy(),status=torch.empty((B,),dtype=torch.int32,pin_memory=True).numpy())
    eng.tick_host(hsc,hgp,20,first_run=True,out=out)
    def timeit(f,n=200):
        for _ in range(10): f()
        torch.cuda.synchronize(); t0=time.perf_counter()
        for _ in range(n): f()
        torch.cuda.synchronize(); return (time.perf_counter()-t0)/n*1e6
    t=lambda a,dt=torch.float32: torch.as_tensor(a).to(device='cuda:0',dtype=dt)
    X=[t(bt['x0']),t(bt['feet']),t(bt['gait']),t(bt['xref']),t(bt['yaw'])]
    res=eng.solve(X[0],X[1],X[2],X[3],yaw=X[4],want=())
    dev=timeit(lambda:(eng.solve(X[0],X[1],X[2],X[3],yaw=X[4],out=res),torch.cuda.synchronize()))
    tick=timeit(lambda:eng.tick_host(hsc,hgp,20,out=out))
    tick1=timeit(lambda:(eng.tick_reset(),eng.tick_host(hsc,hgp,20,first_run=True,out=out)))
    hx=[pin(bt['x0']),pin(bt['feet']),pin(bt['gait']),pin(bt['xref']),pin(bt['yaw'].astype(np.float32))]
    sh=timeit(lambda:eng.solve_host(hx[0],hx[1],hx[2],hx[3],yaw=hx[4],out=out))
    print(f'B={B}: device solve+sync {dev:.1f} us | tick_host {tick:.1f} us (reset + first_run each call: {tick1:.1f}) | solve_host {sh:.1f} us')
