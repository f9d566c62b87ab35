"""Per-stage CUDA-event times of one serial context on the bench scene (K1 / scan / K2 / soft-NMS / K4).
Env: DIAG_K (classes), DIAG_B (images), DIAG_N, DIAG_H / DIAG_W (image size) or DIAG_WORKLOAD (a bench.py workload name), DIAG_RANK,
BOD_K3_THREADS (256 / 512 / 1024)."""
import os, sys, json, ctypes as C, numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bayes_od_rc_b200 import synthetic, _cabi
from bayes_od_rc_b200.engine import BayesODConfig, BayesODEngine
K=int(os.environ.get('DIAG_K','11')); B=int(os.environ.get('DIAG_B','32'))
N=int(os.environ.get('DIAG_N','10'))
kw={}
if os.environ.get('DIAG_H'): kw=dict(im_h=int(os.environ['DIAG_H']),im_w=int(os.environ['DIAG_W']))
if os.environ.get('DIAG_WORKLOAD'):
    import bench
    wl=bench.WORKLOADS[os.environ['DIAG_WORKLOAD']]; K=wl['K']; N=wl['N']; B=int(os.environ.get('DIAG_B',wl['B']))
    kw=dict(im_h=wl['im_h'],im_w=wl['im_w'],config_id=wl['config_id'],**wl.get('spec',{}))
spec=synthetic.SceneSpec(N=N,K=K,**({'config_id':3,**kw}))
batch=synthetic.make_batch(spec,B,device='cuda',with_counts=False)
A=batch['anchors'].shape[0]
cfg=BayesODConfig(use_full_covar=True,max_survivors=32768,ranking_method=os.environ.get('DIAG_RANK','score'))
eng=BayesODEngine(B,N,A,K,cfg)
for i in range(5): eng.run(batch['cls'],batch['box'],batch['cov'],batch['anchors'],None)
eng.stage_ms_accum()
for i in range(20): eng.run(batch['cls'],batch['box'],batch['cov'],batch['anchors'],None)
s,n=eng.stage_ms_accum(); print('K',K,'B',B,'k3 threads',os.environ.get('BOD_K3_THREADS','512'),{k:round(v/n,4) for k,v in s.items()})
if os.environ.get('BOD_K3_DEBUG'):     # needs a library built with -DBOD_DIAGNOSTICS (k3_softnms.cu, bod_api.cu)
    out=(C.c_longlong*(B*384+8))()
    lib=_cabi.load(); lib.bod_debug_k3_counters.argtypes=[C.c_void_p,C.c_void_p]
    print('rc',lib.bod_debug_k3_counters(eng._ctx,out))
    cnt=np.array(out[B*384:]); a=np.array(out[:B*384]).reshape(B,32,12)
    print('counters, cumulative over all launches (-DBOD_DIAGNOSTICS=2: walks, bounded, untouched, products, product entries, folds, woken; '
          '=1: batches ended by the bound of the candidates outside the examined 32 / by a pending list too long for the loop / full, rounds):', cnt)
    nw=int(os.environ.get('BOD_K3_THREADS','512'))//32
    a=a[:,:nw]
    names=['merge','wait barrier 1','rank','pairwise','walk','wait barrier 2','pass A','pass B','listed','spilled weights','rounds','psm']
    rounds=a[:,0,10]
    print('rounds',rounds[:8],'psm',a[:8,0,11])
    print('listed candidates per image',a[:,:,8].sum(1)[:8],'spilled weights read',a[:,:,9].sum(1)[:8])
    for i in range(8):
        print('%-16s warp 0 per round %s | mean over warps %s | max over warps %s' % (names[i], (a[:8,0,i]/rounds[:8]).round(0), (a[:8,:,i].mean(1)/rounds[:8]).round(0), (a[:8,:,i].max(1)/rounds[:8]).round(0)))
    print('total cycles warp 0 (mean over images)', a[:,0,:8].sum(1).mean())
