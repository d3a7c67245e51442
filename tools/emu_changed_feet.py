import sys, ctypes as C, os, re, subprocess
sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests')
import numpy as np
if len(sys.argv)>1 and sys.argv[1]=='child':
    from helpers import make_batch
    from pympc_quadruped_b200.configs import *
    from pympc_quadruped_b200._capi import *
    from pympc_quadruped_b200.gait import Gait
    libt=C.CDLL('/root/repo/tests/emu/_build/libmpcq_emu_trace.so')
    B=256; H=10
    bt=make_batch(A1Config,H,B,'mixed',(Gait.TROTTING10,),3,solve=False)
    cfg=make_config(extract_mpc_constants(bt['cfg'],A1Config), MPCQ_F32)
    rt=np.float32
    x0=bt['x0'].astype(rt); yaw=bt['yaw'].astype(rt); feet=bt['feet'].astype(rt); xref=bt['xref'].astype(rt); gait=bt['gait']
    n=B
    f=np.zeros((n,12),rt); u=np.zeros((n,12*H),rt); iters=np.zeros((n,2),np.int32); resid=np.zeros((n,2)); status=np.zeros(n,np.int32); active=np.zeros((n,4*H),np.uint8)
    p=lambda a:a.ctypes.data_as(C.c_void_p)
    libt.mpcq_emu_solve_f32(C.byref(cfg),n,p(x0),p(yaw),p(feet),p(gait),p(xref),p(f),p(u),p(iters),p(resid),p(status),p(active))
    sys.exit(0)
out=subprocess.run([sys.executable,__file__,'child'],capture_output=True,text=True).stdout
nch=[]; ks=[]
for line in out.splitlines():
    if 'changed steps:' in line:
        nch.append(len(line.split(':')[1].split()))
    m=re.search(r'k_start (\d+) of n (\d+)',line)
    if m: ks.append(int(m.group(1)))
nch=np.array(nch); ks=np.array(ks)
print('factorisations',len(ks))
print('changed feet per factorisation hist', np.bincount(nch))
print('k_start hist (per 4)', np.bincount(ks//4))
later=nch[nch<20]
print('excluding first rounds (20 changed): n', len(later), 'mean changed', later.mean(), 'frac <=2 feet', (later<=2).mean(), ' <=4', (later<=4).mean())
