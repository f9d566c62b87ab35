"""CPU restatement of the reference's uncertainty scoring (TEST INFRASTRUCTURE ONLY; the product
package never imports it): entropies and the minimum-uncertainty-error curve of
``/root/reference/src/core/evaluation_utils_2d.py``.  Pinned by ``tests/golden/mue_*.json``, which
``tests/golden/make_mue_golden.py`` mints by executing the reference's own functions.

Restated from scratch in array form (the reference walks Python lists of dicts):
    gaussian_entropy      evaluation_utils_2d.py:280-285
    categorical_entropy   evaluation_utils_2d.py:288-290
    mu_error              evaluation_utils_2d.py:129-212  (compute_mu_error)
    evaluate_u_error      evaluation_utils_2d.py:236-250
"""
from __future__ import annotations

import numpy as np


def gaussian_entropy(covs):
    """[n,4,4] -> [n] float64: d/2 + d/2 log(2 pi) + log(round(det, 5) + 1e-12) / 2 with d = 4 (:281-284)."""
    covs = np.asarray(covs)
    half = covs.shape[-1] / 2.0
    det = np.round(np.linalg.det(covs.astype(covs.dtype)), 5) + 1e-12
    return half + half * np.log(2 * np.pi) + 0.5 * np.log(det)


def categorical_entropy(params):
    """[n,K] float32 -> [n] float32: -sum p log p, every operation in the array's own precision (:289)."""
    params = np.asarray(params)
    return np.stack([-np.sum(p * np.log(p)) for p in params]) if len(params) else np.zeros(0, params.dtype)


def mu_error(pred_boxes, scores, pred_image, gt_boxes, gt_image, thresholds):
    """One category.  pred_boxes [n,4] / gt_boxes [G,4] float64 [x1,y1,x2,y2]; *_image = image ids.
    Returns (min_u_error, flat arg-min into the [n,T] matrix, ranking, u_error matrix)."""
    thresholds = np.asarray(thresholds, np.float64)
    n, T = len(pred_boxes), len(thresholds)
    ranking = np.argsort(np.asarray(scores, np.float64), kind="stable")          # ascending entropy, ties keep input order (:138-141)
    per_image = {}
    for j, im in enumerate(gt_image):
        per_image.setdefault(int(im), []).append(j)
    taken = np.zeros((len(gt_boxes), T), bool)
    tp = np.zeros((n, T))
    for r, i in enumerate(ranking):
        rows = per_image.get(int(pred_image[i]), [])
        if not rows:
            continue                                                             # no ground truth in the image: FP everywhere
        g = gt_boxes[rows]
        b = pred_boxes[i]
        iw = np.maximum(np.minimum(g[:, 2], b[2]) - np.maximum(g[:, 0], b[0]) + 1.0, 0.0)   # :161-167
        ih = np.maximum(np.minimum(g[:, 3], b[3]) - np.maximum(g[:, 1], b[1]) + 1.0, 0.0)
        inter = iw * ih
        union = ((b[2] - b[0] + 1.0) * (b[3] - b[1] + 1.0) + (g[:, 2] - g[:, 0] + 1.0) * (g[:, 3] - g[:, 1] + 1.0) - inter)
        ov = inter / union
        k = int(np.argmax(ov))                                                   # first maximum (:175-176)
        for t in range(T):                                                       # :181-192
            if ov[k] > thresholds[t] and not taken[rows[k], t]:
                taken[rows[k], t] = True
                tp[r, t] = 1.0
    fp = 1.0 - tp
    total_tp, total_fp = tp.sum(0), fp.sum(0)
    u = 0.5 * (total_tp - np.cumsum(tp, 0)) / np.maximum(total_tp, 1.0) + 0.5 * np.cumsum(fp, 0) / np.maximum(total_fp, 1.0)
    return float(u.min()), int(np.argmin(u)), ranking, u


def evaluate_u_error(gt, pred, iou_thresholds=(0.5,)):
    """Dict-list interface of :236-250 over mu_error (same return values)."""
    cats = sorted({g["category"] for g in gt})
    mins = np.zeros((len(iou_thresholds), len(cats)))
    at = np.zeros((len(iou_thresholds), len(cats)))
    for c, cat in enumerate(cats):
        g = [x for x in gt if x["category"] == cat]
        p = [x for x in pred if x["category"] == cat]
        if not p:
            continue
        names = {}
        for x in g + p:
            names.setdefault(x["name"], len(names))
        m, flat, ranking, _ = mu_error(np.asarray([x["bbox"] for x in p], np.float64), [x["entropy_score"] for x in p],
                                       [names[x["name"]] for x in p], np.asarray([x["bbox"] for x in g], np.float64).reshape(-1, 4),
                                       [names[x["name"]] for x in g], iou_thresholds)
        mins[:, c] = m
        at[:, c] = np.asarray([x["entropy_score"] for x in p])[ranking][flat]
    return mins.flatten().tolist(), float(np.mean(mins)), cats, at.flatten().tolist()
