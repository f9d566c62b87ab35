"""PDQ evaluator end to end on the GPU box: PDQ.score over 64 synthetic 720x1280 images (60 detections, 20 objects each),
wall clock and a cProfile of the host half.  Usage: python scripts/pdq_e2e_probe.py"""
import sys, time, numpy as np
sys.path.insert(0, '.'); sys.path.insert(0, 'scripts')
from bayes_od_rc_b200 import pdq as ppdq
import pdq_bench
rng = np.random.default_rng(0)
H, W = 720, 1280
matches = []
for i in range(64):
    b, c, g = pdq_bench.scene(rng, 60, 20)
    cat = rng.dirichlet(np.ones(8) * 0.3, 60); cat[np.arange(60), rng.integers(0, 7, 60)] += 1.0; cat /= cat.sum(1, keepdims=True)
    gts = [ppdq.GroundTruthBox(x, int(rng.integers(0, 7)), (H, W)) for x in g]
    dets = [ppdq.PBoxDet(cat[j], b[j], [c[j, 0], c[j, 1]]) for j in range(60)]
    matches.append((gts, dets))
ev = ppdq.PDQ((H, W), images_per_call=64)
ev.score(matches)
import cProfile, pstats
t = time.perf_counter(); s = ev.score(matches); dt = time.perf_counter() - t
print('score', s, 'ms per 64 images', dt * 1e3, ev.get_assignment_counts(), ev._engine.last_ms())
pr = cProfile.Profile(); pr.enable(); ev.score(matches); pr.disable()
pstats.Stats(pr).sort_stats('cumulative').print_stats(12)
