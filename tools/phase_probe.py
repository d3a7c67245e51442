import sys, ctypes as C
sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests')
import numpy as np, torch
from helpers import make_batch
from pympc_quadruped_b200 import *
from pympc_quadruped_b200 import _capi
from pympc_quadruped_b200.engine import MpcqEngine
names=['setup_model','hess_apply','build_slots','chol_factor','invert_factor','reduced_gradient','apply_step','pdas_update','refine (incl.)','solve_env(total)','  schur_factor','dual_round (incl.)','dual_apply (incl.)','  schur_solve','  dual_combine']
import os
HH=int(os.environ.get('HH','10')); BB=int(os.environ.get('BB','4096'))
bt=make_batch(A1Config,HH,BB,'mixed',(Gait.STANDING if os.environ.get('GAIT')=='stand' else Gait.TROTTING10,),9,solve=False)
eng=MpcqEngine(bt['cfg'],A1Config)
lib=_capi.load_library()
t=lambda a,dt=torch.float32: torch.as_tensor(a).to(device='cuda:0',dtype=dt)
X=[t(bt['x0']),t(bt['feet']),t(bt['gait']),t(bt['xref']),t(bt['yaw'])]
res=eng.solve(X[0],X[1],X[2],X[3],yaw=X[4]); torch.cuda.synchronize()
it=res.iters[:,0].cpu().numpy()
buf=(C.c_ulonglong*16)()
def probe(idx,label):
    idx=torch.as_tensor(idx,device='cuda:0'); a=[x[idx].contiguous() for x in X]
    eng.solve(a[0],a[1],a[2],a[3],yaw=a[4]); lib.mpcq_debug_phase_cycles(buf)
    eng.solve(a[0],a[1],a[2],a[3],yaw=a[4]); lib.mpcq_debug_phase_cycles(buf)
    v=np.array(list(buf)[:15],dtype=float); rounds=it[idx.cpu().numpy()].sum()
    tot=v[9]
    ne=len(idx)
    print(f'--- {label}: envs {ne} rounds {rounds}  total {tot/ne:.0f} cycles/env')
    for n,c in zip(names,v): print(f'   {n:20s} {c/ne:9.0f} cycles/env  {100*c/tot:5.1f}%')
sel=np.flatnonzero(it==int(np.median(it)))[:1]
probe(sel,'one env alone (median rounds)')
probe(np.argsort(-it)[:1],'the env with the most rounds, alone')
probe(np.arange(BB),'whole batch under load')
