import sys, time, ctypes as C
sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests')
import numpy as np
from helpers import make_batch
from pympc_quadruped_b200.configs import *
from pympc_quadruped_b200._capi import *
from pympc_quadruped_b200.gait import Gait
lib=C.CDLL('/root/repo/tests/emu/_build/libmpcq_emu.so')
def run(robot,H,B,regime,gaits,seed,**knobs):
    bt=make_batch(robot,H,B,regime,gaits,seed,solve=False)
    cfg=make_config(extract_mpc_constants(bt['cfg'],robot), MPCQ_F32, **knobs)
    rt=np.float32
    x0=bt['x0'].astype(rt); yaw=bt['yaw'].astype(rt); feet=bt['feet'].astype(rt); xref=bt['xref'].astype(rt); gait=bt['gait']
    f=np.zeros((B,12),rt); u=np.zeros((B,12*H),rt); iters=np.zeros((B,2),np.int32); resid=np.zeros((B,2)); status=np.zeros(B,np.int32); active=np.zeros((B,4*H),np.uint8)
    p=lambda a:a.ctypes.data_as(C.c_void_p)
    t=time.time()
    rc=lib.mpcq_emu_solve_f32(C.byref(cfg),B,p(x0),p(yaw),p(feet),p(gait),p(xref),p(f),p(u),p(iters),p(resid),p(status),p(active))
    dt=time.time()-t
    print(f'rc={rc} {dt/B*1e3:.1f} ms/env status {np.bincount(status,minlength=4)}')
    pc=(C.c_long*16)(); lib.mpcq_emu_phase_calls(pc); print('per env: chol %.2f tri %.2f hess %.2f redgrad %.2f'%(pc[3]/B,pc[4]/B,pc[1]/B,pc[5]/B))
    print('facts hist', np.bincount(iters[:,0]))
    print('AS hist', np.bincount(iters[:,1]))
    print('facts mean', iters[:,0].mean(), 'sum over top-1%', np.sort(iters[:,0])[-B//100:].sum(), 'total', iters[:,0].sum())
    return bt,u,status,iters
if __name__=='__main__':
    B=int(sys.argv[1]); seed=int(sys.argv[2]) if len(sys.argv)>2 else 3
    kn={}
    for a in sys.argv[3:]:
        k,v=a.split("="); kn[k]=float(v) if "." in v or "e" in v else int(v)
    import os; run(A1Config,int(os.environ.get('HH','10')),B,os.environ.get('REG','mixed'),(Gait.TROTTING10,),seed,**kn)
