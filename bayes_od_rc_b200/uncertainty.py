"""Drop-in for the uncertainty scoring of ``src/core/evaluation_utils_2d.py`` that the
offline scripts ``src/retina_net/offline_eval/{bdd,kitti}/compute_uncertainty_error.py``
drive (:91-132 there): same function names, arguments and return values, computed on the
GPU through ``bod_entropies`` / ``bod_mu_error`` (csrc/ku_uncertainty.cu).

    compute_gaussian_entropy_np(cov)            evaluation_utils_2d.py:280-285
    compute_categorical_entropy_np(cat_params)  evaluation_utils_2d.py:288-290
    compute_mu_error(gt, predictions, thresholds)   :129-212
    evaluate_u_error(gt, pred, iou_thresholds)      :236-250

``gaussian_entropies`` / ``categorical_entropies`` are the batched forms a script should
call once per frame (or once per validation set) instead of once per detection.  Install
over the reference with::

    import bayes_od_rc_b200.uncertainty as fast
    from src.core import evaluation_utils_2d as ev
    for name in ("compute_gaussian_entropy_np", "compute_categorical_entropy_np",
                 "compute_mu_error", "evaluate_u_error"):
        setattr(ev, name, getattr(fast, name))

There is no CPU implementation here; without the CUDA library the calls fail.
"""
from __future__ import annotations

import ctypes as C
from collections import defaultdict

import numpy as np

from . import _cabi


def _check(rc):
    if rc != _cabi.BOD_OK:
        lib = _cabi.load()
        raise _cabi.BodError(rc, (lib.bod_uncertainty_last_error() or b"").decode())


def gaussian_entropies(covs, device=0) -> np.ndarray:
    """[n,4,4] covariances -> [n] binary64 entropies (compute_gaussian_entropy_np of each)."""
    covs = np.ascontiguousarray(covs, np.float32).reshape(-1, 4, 4)
    out = np.empty(len(covs), np.float64)
    _check(_cabi.load().bod_entropies(int(device), len(covs), 0, covs.ctypes.data, None, out.ctypes.data, None))
    return out


def categorical_entropies(cat_params, device=0) -> np.ndarray:
    """[n,K] categorical parameter vectors -> [n] binary32 entropies (compute_categorical_entropy_np of each)."""
    p = np.ascontiguousarray(cat_params, np.float32)
    p = p.reshape(-1, p.shape[-1])
    out = np.empty(len(p), np.float32)
    _check(_cabi.load().bod_entropies(int(device), len(p), p.shape[1], None, p.ctypes.data, None, out.ctypes.data))
    return out


def compute_gaussian_entropy_np(cov):
    cov = np.asarray(cov)
    if cov.shape != (4, 4):
        raise ValueError("the GPU path covers the 4x4 box covariances of this repo (evaluation_utils_2d.py:280-285)")
    return gaussian_entropies(cov[None])[0]


def compute_categorical_entropy_np(cat_params):
    return categorical_entropies(np.asarray(cat_params)[None])[0]


def group_by_key(detections, key):                  # evaluation_utils_2d.py:272-276
    groups = defaultdict(list)
    for d in detections:
        groups[d[key]].append(d)
    return groups


def compute_mu_error(gt, predictions, thresholds, device=0):
    """Minimum uncertainty error of one category (:129-212): returns (min_u_error, score_at_min_u_error).
    Unlike the reference it does not annotate the prediction dicts with 'iou' / 'is_tp'."""
    names = {}
    for g in gt:
        names.setdefault(g['name'], len(names))
    for p in predictions:
        names.setdefault(p['name'], len(names))
    n_images = len(names)
    n = len(predictions)
    if n == 0:
        raise ValueError("compute_mu_error: no predictions (the reference raises on np.min of an empty array)")
    # ground truth as CSR rows per image, in the order the boxes appear (np.argmax ties -> first box)
    gt_img = np.fromiter((names[g['name']] for g in gt), np.int64, len(gt))
    gorder = np.argsort(gt_img, kind='stable')
    gt_boxes = np.asarray([[float(z) for z in g['bbox']] for g in gt], np.float64).reshape(-1, 4)[gorder]
    gt_off = np.zeros(n_images + 1, np.int32)
    np.cumsum(np.bincount(gt_img, minlength=n_images), out=gt_off[1:])
    scores = np.asarray([p['entropy_score'] for p in predictions], np.float64)
    order = np.argsort(scores, kind='stable').astype(np.int32)          # sorted(..., key=entropy_score), stable (:138-141)
    pimg = np.fromiter((names[p['name']] for p in predictions), np.int32, n)
    pboxes = np.asarray([[float(z) for z in p['bbox']] for p in predictions], np.float64).reshape(n, 4)
    rank_img = pimg[order]
    by_image = np.argsort(rank_img, kind='stable').astype(np.int32)     # ranks grouped by image, ascending inside
    img_off = np.zeros(n_images + 1, np.int32)
    np.cumsum(np.bincount(rank_img, minlength=n_images), out=img_off[1:])
    thr = np.ascontiguousarray(thresholds, np.float64).reshape(-1)
    mn, at = C.c_double(), C.c_int64()
    ptr = lambda a: a.ctypes.data        # noqa: E731
    _check(_cabi.load().bod_mu_error(int(device), n, ptr(pboxes), ptr(pimg), ptr(order), n_images, ptr(img_off), ptr(by_image),
                                     ptr(gt_off), ptr(gt_boxes) if len(gt_boxes) else None, len(thr), ptr(thr),
                                     C.byref(mn), C.byref(at), None))
    ranked_scores = np.asarray([predictions[i]['entropy_score'] for i in order])
    return mn.value, ranked_scores[at.value]         # the reference indexes the ranked scores with the FLAT arg-min (:211-213)


def evaluate_u_error(gt, pred, iou_thresholds=[0.5], device=0):      # noqa: B006  (the reference's signature, :236)
    cat_gt = group_by_key(gt, 'category')
    cat_pred = group_by_key(pred, 'category')
    cat_list = sorted(cat_gt.keys())
    min_u_errors = np.zeros((len(iou_thresholds), len(cat_list)))
    scores_at_min_u_errors = np.zeros((len(iou_thresholds), len(cat_list)))
    for i, cat in enumerate(cat_list):
        if cat in cat_pred:
            min_u_errors[:, i], scores_at_min_u_errors[:, i] = compute_mu_error(cat_gt[cat], cat_pred[cat], iou_thresholds, device)
    min_u_error = np.mean(min_u_errors)
    return min_u_errors.flatten().tolist(), min_u_error, cat_list, scores_at_min_u_errors.flatten().tolist()
