"""ctypes binding of oracle/pdq_oracle.c — the CPU restatement of the PDQ spatial-quality path
(pdq_data_holders.py:92-247, pdq.py:199-230).  TEST INFRASTRUCTURE ONLY, like the rest of oracle/."""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from . import build as _build

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "_build", "libpdq_oracle.so")
_lib = None


def lib():
    global _lib
    if _lib is None:
        _build()
        _lib = C.CDLL(_LIB_PATH)
        _lib.orc_pdq_phi.restype = C.c_double
        _lib.orc_pdq_phi.argtypes = [C.c_double]
        _lib.orc_pdq_bvn_cdf.restype = C.c_double
        _lib.orc_pdq_bvn_cdf.argtypes = [C.c_double] * 3
    return _lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def bvn_cdf(h, k, r) -> float:
    return lib().orc_pdq_bvn_cdf(float(h), float(k), float(r))


def find_roi(img_size, mean, cov):
    """find_roi(img_size, mean=[y, x], cov=[[vy, c], [c, vx]]) -> [x1, y1, x2, y2]"""
    roi = np.zeros(4, np.int32)
    m = np.ascontiguousarray(mean, np.float64)
    c = np.ascontiguousarray(cov, np.float64)
    rc = lib().orc_pdq_find_roi(int(img_size[0]), int(img_size[1]), _p(m), _p(c), _p(roi))
    if rc:
        raise ValueError("find_roi: the reference raises for this input")
    return roi.tolist()


def single_heatmap(img_size, mean, cov) -> np.ndarray:
    hm = np.empty(tuple(img_size), np.float32)
    m = np.ascontiguousarray(mean, np.float64)
    c = np.ascontiguousarray(cov, np.float64)
    rc = lib().orc_pdq_single_heatmap(int(img_size[0]), int(img_size[1]), _p(m), _p(c), _p(hm))
    if rc:
        raise ValueError(f"gen_single_heatmap: rc={rc}")
    return hm


def heatmap(img_size, box, covs) -> np.ndarray:
    """PBoxDetInst(class_list, box, covs).calc_heatmap(img_size)"""
    hm = np.empty(tuple(img_size), np.float32)
    b = np.ascontiguousarray(box, np.int32)
    c = np.ascontiguousarray(covs, np.float64).reshape(8)
    rc = lib().orc_pdq_heatmap(int(img_size[0]), int(img_size[1]), _p(b), _p(c), _p(hm))
    if rc:
        raise ValueError(f"calc_heatmap: rc={rc}")
    return hm


def heatmaps(img_size, boxes, covs) -> np.ndarray:
    return np.stack([heatmap(img_size, b, c) for b, c in zip(boxes, covs)]) if len(boxes) else \
        np.zeros((0,) + tuple(img_size), np.float32)


def losses(heatmaps_dhw, gt_boxes):
    """-> fg_loss [G,D], bg_loss [G,D], bg_total [D] (binary64)"""
    hm = np.ascontiguousarray(heatmaps_dhw, np.float32)
    D, H, W = hm.shape
    gt = np.ascontiguousarray(gt_boxes, np.int32).reshape(-1, 4)
    G = gt.shape[0]
    fg = np.zeros((G, D)); bg = np.zeros((G, D)); tot = np.zeros(D)
    lib().orc_pdq_losses(H, W, G, _p(gt), D, _p(hm), _p(fg), _p(bg), _p(tot))
    return fg, bg, tot
