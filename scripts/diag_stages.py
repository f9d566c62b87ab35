import os, sys, json, ctypes as C, numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bayes_od_rc_b200 import synthetic, _cabi
from bayes_od_rc_b200.engine import BayesODConfig, BayesODEngine
K=int(os.environ.get('DIAG_K','11')); B=32
spec=synthetic.SceneSpec(N=10,K=K,config_id=3)
batch=synthetic.make_batch(spec,B,device='cuda',with_counts=False)
A=batch['anchors'].shape[0]
cfg=BayesODConfig(use_full_covar=True,max_survivors=32768,ranking_method=os.environ.get('DIAG_RANK','score'))
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
    r=a[:,3]
    names=['pass A','pass B','rank','rounds','pairwise','walk+writes','warp merge','wait at barrier']
    print('rounds',a[:8,3])
    tot=sum(a[:,i] for i in (0,1,2,4,5,6,7))
    for i in (6,7,2,4,5,0,1):
        print('%-16s per round %s share %.2f' % (names[i], (a[:,i]/r).round(0)[:8], a[:,i].sum()/tot.sum()))
    print('total cycles per image (max)', tot.max(), 'mean', tot.mean())
