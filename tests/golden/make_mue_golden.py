"""Mints tests/golden/mue_*.json by executing the reference's own uncertainty scoring
(/root/reference/src/core/evaluation_utils_2d.py: compute_gaussian_entropy_np :280-285,
compute_categorical_entropy_np :288-290, compute_mu_error :129-212, evaluate_u_error :236-250)
on seeded synthetic detections / ground truth shaped like what
offline_eval/bdd/compute_uncertainty_error.py:91-132 builds.  Run here (the reference is not on
the GPU box); the fixtures are committed."""
import copy
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, "/root/reference")
from src.core import evaluation_utils_2d as ev      # noqa: E402

CATS = ['car', 'truck', 'bus', 'person', 'rider', 'bike', 'motor', 'bkgrnd']


def scene(rng, n_images, dets_per_image, gts_per_image, K=8, ties=False):
    gt, pred, covs, params = [], [], [], []
    for im in range(n_images):
        name = f"frame_{im:04d}.jpg"
        G = int(rng.integers(0, gts_per_image + 1))
        boxes = []
        for _ in range(G):
            x1, y1 = rng.uniform(0, 1100), rng.uniform(0, 600)
            w, h = rng.uniform(20, 300), rng.uniform(20, 200)
            b = [float(np.float32(x1)), float(np.float32(y1)), float(np.float32(x1 + w)), float(np.float32(y1 + h))]
            boxes.append(b)
            gt.append({'name': name, 'category': CATS[int(rng.integers(0, 4))], 'bbox': b})
        D = int(rng.integers(1, dets_per_image + 1))
        for _ in range(D):
            if boxes and rng.uniform() < 0.7:                       # a detection near a ground-truth box
                b = np.asarray(boxes[int(rng.integers(0, len(boxes)))]) + rng.normal(0, 6, 4)
            else:
                x1, y1 = rng.uniform(0, 1100), rng.uniform(0, 600)
                b = np.asarray([x1, y1, x1 + rng.uniform(20, 300), y1 + rng.uniform(20, 200)])
            b = [float(np.float32(v)) for v in b]
            L = np.tril(rng.normal(0, 0.4, (4, 4)), -1) + np.diag(rng.uniform(1.0, 3.0, 4))
            cov = (L @ L.T * 70.0 * rng.uniform(0.05, 1.0)).astype(np.float32)
            p = rng.dirichlet(np.full(K, 0.3)).astype(np.float32) + np.float32(1e-6)
            p = (p / p.sum()).astype(np.float32)
            if ties:
                p = np.round(p, 1).astype(np.float32) + np.float32(0.0125)
                p = (p / p.sum()).astype(np.float32)
            covs.append(cov); params.append(p)
            pred.append({'name': name, 'category': CATS[int(np.argmax(p)) % 4], 'bbox': b})
    return gt, pred, np.stack(covs), np.stack(params)


def main():
    cases = {"mue_small": dict(seed=1, n_images=12, dets=8, gts=5), "mue_ties": dict(seed=2, n_images=20, dets=10, gts=4, ties=True),
             "mue_large": dict(seed=3, n_images=150, dets=24, gts=10)}
    for name, kw in cases.items():
        rng = np.random.default_rng(kw["seed"])
        gt, pred, covs, params = scene(rng, kw["n_images"], kw["dets"], kw["gts"], ties=kw.get("ties", False))
        g_ent = [float(ev.compute_gaussian_entropy_np(c)) for c in covs]
        c_ent = [float(ev.compute_categorical_entropy_np(p)) for p in params]
        out = {}
        for method, ent in (("gaussian", g_ent), ("categorical", [np.float32(x) for x in c_ent])):
            pr = copy.deepcopy(pred)
            for p, e in zip(pr, ent):
                p['entropy_score'] = e
            for thr in ([0.5], [0.7]):
                res = ev.evaluate_u_error(copy.deepcopy(gt), copy.deepcopy(pr), iou_thresholds=thr)
                out[f"{method}@{thr[0]}"] = dict(min_u_errors=res[0], mean=float(res[1]), cats=res[2],
                                                 scores_at_min=[float(x) for x in res[3]])
        names = sorted({x['name'] for x in gt + pred})
        nid = {n: i for i, n in enumerate(names)}
        np.savez_compressed(
            os.path.join(HERE, name + ".npz"),
            meta=json.dumps(dict(case=name, generator="tests/golden/make_mue_golden.py; reference functions executed verbatim",
                                 numpy=np.__version__, categories=CATS, **kw)),
            gt_image=np.asarray([nid[x['name']] for x in gt], np.int32), gt_cat=np.asarray([CATS.index(x['category']) for x in gt], np.int32),
            gt_box=np.asarray([x['bbox'] for x in gt], np.float64).reshape(-1, 4),
            pred_image=np.asarray([nid[x['name']] for x in pred], np.int32),
            pred_cat=np.asarray([CATS.index(x['category']) for x in pred], np.int32),
            pred_box=np.asarray([x['bbox'] for x in pred], np.float64).reshape(-1, 4),
            covs=covs, params=params, gaussian_entropy=np.asarray(g_ent, np.float64), categorical_entropy=np.asarray(c_ent, np.float32),
            results=json.dumps(out))
        print(name, len(gt), "gt", len(pred), "pred", {k: round(v["mean"], 6) for k, v in out.items()})


if __name__ == "__main__":
    main()
