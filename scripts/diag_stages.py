import os, sys, json, ctypes as C, numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bayes_od_rc_b200 import synthetic, _cabi
from bayes_od_rc_b200.engine import BayesODConfig, BayesODEngine
K=int(os.environ.get('DIAG_K','11')); B=32
spec=synthetic.SceneSpec(N=10,K=K,config_id=3)
batch=synthetic.make_batch(spec,B,device='cuda',with_counts=False)
A=batch['anchors'].shape[0]
cfg=BayesODConfig(use_full_covar=True,max_survivors=32768)
eng=BayesODEngine(B,10,A,K,cfg)
for i in range(5): eng.run(batch['cls'],batch['box'],batch['cov'],batch['anchors'],None)
eng.stage_ms_accum()
for i in range(20): eng.run(batch['cls'],batch['box'],batch['cov'],batch['anchors'],None)
s,n=eng.stage_ms_accum(); print('K',K,'k1dbg',os.environ.get('BOD_K1_DEBUG'),{k:round(v/n,4) for k,v in s.items()})
if os.environ.get('BOD_K3_DEBUG'):
    out=(C.c_longlong*(B*8))()
    lib=_cabi.load(); lib.bod_debug_k3_counters.argtypes=[C.c_void_p,C.c_void_p]
    print('rc',lib.bod_debug_k3_counters(eng._ctx,out))
    a=np.array(out[:]).reshape(B,8)
    print('rounds',a[:8,3],'dets',a[:8,6],'S',a[:8,4])
    print('per round cycles: C',(a[:,5]/a[:,3]).round(0)[:8],'A',(a[:,0]/a[:,3]).round(0)[:8],'B',(a[:,1]/a[:,3]).round(0)[:8],'list/round',(a[:,2]/a[:,3]).round(1)[:8])
    print('total cycles per image (max)', (a[:,0]+a[:,1]+a[:,5]).max(), 'mean', (a[:,0]+a[:,1]+a[:,5]).mean())
