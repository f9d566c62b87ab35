"""Device timeline of a pipelined context (diagnostics build: every stage kernel stamps %globaltimer at the start of
its first CTA, at the start of the CTA placed last and at its end; bod_debug_timeline returns the stamps of every lane's
last run).  Prints the last `lanes` consecutive runs, times in microseconds relative to the first moments kernel.
Env: TL_WORKLOAD (bench.py workload name), TL_B (images), TL_LANES, TL_STEPS."""
import os, sys, ctypes as C, numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ.setdefault('BOD_TIMELINE', '1')
import bench
from bayes_od_rc_b200 import synthetic, _cabi
from bayes_od_rc_b200.engine import BayesODConfig, BayesODEngine
wl = dict(bench.WORKLOADS[os.environ.get('TL_WORKLOAD', 'bdd_covar_b32_k11')])
B = int(os.environ.get('TL_B', wl['B'])); lanes = int(os.environ.get('TL_LANES', '4')); steps = int(os.environ.get('TL_STEPS', '41'))
spec = synthetic.SceneSpec(im_h=wl['im_h'], im_w=wl['im_w'], N=wl['N'], K=wl['K'], config_id=wl['config_id'], **wl.get('spec', {}))
batch = synthetic.make_batch(spec, B, device='cuda', with_counts=False)
A = batch['anchors'].shape[0]
cfg = BayesODConfig(use_full_covar=wl['use_full_covar'], cov_layout=_cabi.COV_FULL16, max_survivors=min(A, 32768), pipeline_depth=lanes,
                    scale_v=wl.get('scale_v', 1.0), scale_u=wl.get('scale_u', 1.0))
eng = BayesODEngine(B, wl['N'], A, wl['K'], cfg)
st = torch.cuda.Stream()
for i in range(steps):
    eng.run(batch['cls'], batch['box'], batch['cov'], batch['anchors'], None, stream=st.cuda_stream)
    if not os.environ.get('TL_NOFETCH'):
        eng.fetch_async()
eng.wait_results(st.cuda_stream); torch.cuda.synchronize()
lib = _cabi.load(); lib.bod_debug_timeline.argtypes = [C.c_void_p, C.c_void_p]
out = (C.c_ulonglong * (20 * lanes))()
print('rc', lib.bod_debug_timeline(eng._ctx, out))
t = np.array(out[:], dtype=np.float64).reshape(lanes, 5, 4)
order = np.argsort(t[:, 0, 0])
t0 = t[order[0], 0, 0]
names = ['moments', 'scan', 'posterior', 'soft-NMS', 'fusion']
print(f'workload {os.environ.get("TL_WORKLOAD", "bdd_covar_b32_k11")} B={B} lanes={lanes}: start of first CTA / start of last CTA / end, us')
for l in order:
    print(f'lane {l}: ' + ' | '.join(f'{names[k]} {(t[l,k,0]-t0)/1e3:8.1f} {(t[l,k,1]-t0)/1e3:8.1f} {(t[l,k,2]-t0)/1e3:8.1f}' for k in range(5)))
