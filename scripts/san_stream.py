"""Small pipelined streaming run for compute-sanitizer (scripts/sanitize_gpu.sh): B = 2, four lanes, held inputs, eight runs,
every run's result block fetched by ticket and compared with a serial context."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
from bayes_od_rc_b200 import synthetic
from bayes_od_rc_b200.engine import BayesODConfig, BayesODEngine
spec = synthetic.SceneSpec(im_h=96, im_w=160, N=6, K=8, g_min=4, g_max=6, box_hi=90., config_id=9)
B, runs = 2, 8
batches = [synthetic.make_batch(spec, B, device='cuda', with_counts=False, first_image_id=10 * i) for i in range(runs)]
A = batches[0]['anchors'].shape[0]
ser = BayesODEngine(B, 6, A, 8, BayesODConfig(use_full_covar=True))
ref = []
for b in batches:
    ser.run(b['cls'], b['box'], b['cov'], b['anchors'], None); ref.append(ser.fetch())
for graphs in ('0', '2'):
    os.environ['BOD_GRAPHS'] = graphs
    eng = BayesODEngine(B, 6, A, 8, BayesODConfig(use_full_covar=True, pipeline_depth=4))
    eng.set_input_hold(True)
    st = torch.cuda.Stream()
    tickets = []
    for b in batches:
        eng.run(b['cls'], b['box'], b['cov'], b['anchors'], None, stream=st.cuda_stream)
        tickets.append(eng.fetch_async())
        if len(tickets) >= 4:
            i = len(tickets) - 4
            r = eng.collect(tickets[i])
            assert np.array_equal(r.nms_indices, ref[i].nms_indices) and np.array_equal(r.means.view(np.uint32), ref[i].means.view(np.uint32)), i
    for i in range(runs - 3, runs):
        r = eng.collect(tickets[i])
        assert np.array_equal(r.nms_indices, ref[i].nms_indices), i
    del eng
print('stream ok')
