"""ROI parity stress on the GPU box: dense PDQ maps of random corners (strong correlation, sub-pixel to image-sized sigmas,
clipped windows) against oracle/pdq_oracle.c.  Usage: python scripts/pdq_roi_stress.py"""
import sys, numpy as np
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
from bayes_od_rc_b200 import pdq
from oracle import pdq as opdq
rng = np.random.default_rng(123)
H, W = 97, 143
eng = pdq.PdqEngine((H, W))
bad = 0; n = 0
for it in range(60):
    D = 8
    boxes, covs = [], []
    for _ in range(D):
        x1, y1 = rng.integers(0, W - 12), rng.integers(0, H - 12)
        x2, y2 = rng.integers(x1 + 4, W), rng.integers(y1 + 4, H)
        cs = []
        for _ in range(2):
            s1, s2 = np.exp(rng.uniform(np.log(0.3), np.log(40), 2))
            r = rng.uniform(-0.97, 0.97)
            cs.append(np.array([[s1 * s1, r * s1 * s2], [r * s1 * s2, s2 * s2]]))
        boxes.append([x1, y1, x2, y2]); covs.append(cs)
    boxes = np.array(boxes, np.int32); covs = np.array(covs)
    try:
        hm = eng.heatmaps(boxes, covs)
    except Exception as e:
        # the reference raises for some corners; check the oracle agrees that at least one raises
        try:
            opdq.heatmaps((H, W), boxes, covs); print('GPU raised, oracle did not', e); bad += 1
        except ValueError:
            pass
        continue
    ohm = opdq.heatmaps((H, W), boxes, covs)
    n += D
    if not (np.array_equal(hm > 0, ohm > 0) and np.abs(hm - ohm).max() <= 2e-7):
        bad += 1; print('mismatch', it, np.abs(hm - ohm).max(), np.count_nonzero((hm > 0) != (ohm > 0)))
print('checked', n, 'detections; mismatching batches:', bad)
