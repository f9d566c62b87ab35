"""PDQ evaluation on the GPU: a drop-in for the reference's ``src/retina_net/offline_eval/pdq.py`` whose
heavy half — one dense Gaussian-corner heat map per detection (``pdq_data_holders.py:92-247``) contracted
with one dense mask per ground-truth object (``pdq.py:199-230``), in a ``multiprocessing.Pool`` of
CPU workers (``pdq.py:76-77``) — runs as four CUDA kernels over a whole batch of images
(``csrc/kp_pdq.cu`` behind ``bod_pdq_*`` of ``include/bayesod.h``).

Same names, arguments and results as the reference:

    evaluator = PDQ();  evaluator.score(match_list);  evaluator.get_assignment_counts() ...

``match_list`` holds ``(gt_instances, det_instances)`` per image, exactly what
``bdd/compute_pdq.py:93-124`` builds: any objects with the attributes of
``pdq_data_holders.GroundTruthInstance`` (``bounding_box``, ``class_label``, ``segmentation_mask`` or
``num_pixels``) and ``pdq_data_holders.PBoxDetInst`` (``box``, ``covs``, ``class_list``) work —
``GroundTruthBox`` and ``PBoxDet`` below are light versions that do not carry a dense mask.  Ground truth
must be box-shaped (it is, in both ``compute_pdq.py`` scripts): the mask of an instance is checked
against its box when it has one.

What stays on the host is the reference's own small-matrix logic, restated from ``pdq.py:233-446``:
qualities from the loss sums, the Hungarian assignment (``scipy.optimize.linear_sum_assignment``, as
the reference) and the TP / FP / FN bookkeeping.  There is no CPU fallback for the heat maps.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _cabi
from ._cabi import BodError

_SMALL_VAL = 1e-14          # pdq.py:8


class GroundTruthBox:
    """``pdq_data_holders.GroundTruthInstance`` for a box-shaped object without the dense mask:
    ``bounding_box = [x1, y1, x2, y2]`` (ints); foreground = rows ``[y1, y2)`` x columns ``[x1, x2)``
    (``compute_pdq.py:108-110``)."""

    def __init__(self, bounding_box, true_class_label, img_size=None):
        self.bounding_box = np.asarray(bounding_box, np.int32)
        self.class_label = int(true_class_label)
        x1, y1, x2, y2 = (int(v) for v in self.bounding_box)
        if img_size is not None:
            x2, y2 = min(x2, int(img_size[1])), min(y2, int(img_size[0]))
        self.num_pixels = max(x2 - max(x1, 0), 0) * max(y2 - max(y1, 0), 0)
        self.segmentation_mask = None


class PBoxDet:
    """``pdq_data_holders.PBoxDetInst`` without ``calc_heatmap`` (``box = [x1, y1, x2, y2]``,
    ``covs = [cov_top_left, cov_bottom_right]``, each ``[[var_x, c], [c, var_y]]``)."""

    def __init__(self, class_list, box, covs):
        self.class_list = np.asarray(class_list)
        self.box = np.asarray(box)
        self.covs = covs


_CORNER_TRANSFORM = np.array([[0, 1, 0, -0.5], [1, 0, -0.5, 0], [0, 1, 0, 0.5], [1, 0, 0.5, 0]])     # compute_pdq.py:98-101


def det_instances_from_arrays(means_vuhw, covs_vuhw, cat_params, min_score=0.5445, cov_scale=70.0):
    """The detection list of one image as ``bdd/compute_pdq.py:93-124`` builds it from the saved ``mean`` /
    ``cov`` / ``cat_param`` arrays (``[D,4]``, ``[D,4,4]``, ``[D,K]``): corner covariances ``T.Sigma.T^T * cov_scale``,
    corners of ``vuhw_to_vuvu_np`` truncated to int32, detections whose best class score is below ``min_score``
    dropped.  ``kitti/compute_pdq.py:119-148`` is the same with ``min_score=0.5`` (its x70 is applied when loading)."""
    means = np.asarray(means_vuhw)
    if np.asarray(covs_vuhw).size == 0:
        return []
    means = means.reshape(-1, 4)
    cov_t = np.matmul(np.matmul(_CORNER_TRANSFORM, np.asarray(covs_vuhw)), _CORNER_TRANSFORM.T) * cov_scale
    v, u, h, w = means[:, 0], means[:, 1], means[:, 2], means[:, 3]
    vuvu = np.stack((v - h / 2.0, u - w / 2.0, v + h / 2.0, u + w / 2.0), axis=1)            # box_utils.py:70-88
    dets = []
    for cat, b, cv in zip(np.asarray(cat_params), vuvu, cov_t):
        if np.max(cat) >= min_score:
            dets.append(PBoxDet(cat, np.array([b[1], b[0], b[3], b[2]]).astype(np.int32), [cv[0:2, 0:2], cv[2:4, 2:4]]))
    return dets


def _num_pixels(gt, img_size) -> int:
    mask = getattr(gt, "segmentation_mask", None)
    if mask is None:
        return int(gt.num_pixels)
    x1, y1, x2, y2 = (int(v) for v in gt.bounding_box)
    n = int(np.count_nonzero(mask))
    if tuple(mask.shape) != tuple(img_size) or min(x1, y1) < 0 or n != int(np.count_nonzero(mask[y1:y2, x1:x2])) or \
            n != max(min(x2, mask.shape[1]) - x1, 0) * max(min(y2, mask.shape[0]) - y1, 0):
        raise ValueError("ground-truth mask is not the box [x1:x2, y1:y2] of its bounding_box: only box-shaped "
                         "ground truth (compute_pdq.py:107-113) is supported")
    return n


def _is_gt_included(gt, num_pixels) -> bool:
    """pdq.py:455-471"""
    bb = gt.bounding_box
    return bool(bb[2] - bb[0] > 10 and bb[3] - bb[1] > 10 and num_pixels > 100)


class PdqEngine:
    """One ``bod_pdq_ctx``: bound to a device and an image size."""

    def __init__(self, img_size, device: int = 0):
        self.img_size = (int(img_size[0]), int(img_size[1]))
        self._lib = _cabi.load()
        self._h = C.c_void_p()
        rc = self._lib.bod_pdq_create(C.byref(self._h), int(device), *self.img_size)
        if rc != _cabi.BOD_OK:
            self._h = None
            raise BodError(rc, (self._lib.bod_pdq_last_error(None) or b"").decode())

    def close(self):
        if getattr(self, "_h", None):
            self._lib.bod_pdq_destroy(self._h)
            self._h = None

    __del__ = close

    def _check(self, rc):
        if rc != _cabi.BOD_OK:
            raise BodError(rc, (self._lib.bod_pdq_last_error(self._h) or b"").decode())

    @staticmethod
    def _dets(boxes, covs):
        b = np.ascontiguousarray(np.asarray(boxes).reshape(-1, 4), np.int32)
        c = np.ascontiguousarray(np.asarray(covs, np.float64).reshape(-1, 8))
        if len(b) != len(c):
            raise ValueError("boxes and covs disagree on the number of detections")
        return b, c

    def heatmaps(self, boxes, covs, out=None) -> np.ndarray:
        """``np.stack([PBoxDetInst(_, box, covs).calc_heatmap(img_size) ...])`` -> ``[D, H, W]`` float32.
        ``out``: an optional CUDA ``torch.Tensor`` of that shape to fill in place (no host copy)."""
        b, c = self._dets(boxes, covs)
        D = len(b)
        if out is not None:
            if tuple(out.shape) != (D,) + self.img_size or not out.is_cuda or not out.is_contiguous() or out.element_size() != 4:
                raise ValueError("out must be a contiguous CUDA float32 tensor of shape [D, H, W]")
            self._check(self._lib.bod_pdq_heatmaps(self._h, D, b.ctypes.data, c.ctypes.data, out.data_ptr(), 1))
            return out
        hm = np.empty((D,) + self.img_size, np.float32)
        self._check(self._lib.bod_pdq_heatmaps(self._h, D, b.ctypes.data, c.ctypes.data, hm.ctypes.data, 0))
        return hm

    def losses(self, det_offsets, boxes, covs, gt_offsets, gt_boxes):
        """Batched ``_calc_fg_loss`` / ``_calc_bg_loss``: lists of ``[G_b, D_b]`` float64 matrices, and the
        whole-image background term ``[D]``."""
        b, c = self._dets(boxes, covs)
        do = np.ascontiguousarray(det_offsets, np.int32)
        go = np.ascontiguousarray(gt_offsets, np.int32)
        gt = np.ascontiguousarray(np.asarray(gt_boxes).reshape(-1, 4), np.int32)
        n = len(do) - 1
        if len(go) != n + 1 or do[-1] != len(b) or go[-1] != len(gt):
            raise ValueError("offsets do not match the arrays")
        sizes = [(int(go[i + 1] - go[i]), int(do[i + 1] - do[i])) for i in range(n)]
        total = sum(g * d for g, d in sizes)
        fg = np.zeros(max(total, 1)); bg = np.zeros(max(total, 1)); tot = np.zeros(max(len(b), 1))
        self._check(self._lib.bod_pdq_losses(self._h, n, do.ctypes.data, b.ctypes.data, c.ctypes.data, go.ctypes.data,
                                             gt.ctypes.data, fg.ctypes.data, bg.ctypes.data, tot.ctypes.data))
        fgs, bgs, o = [], [], 0
        for g, d in sizes:
            fgs.append(fg[o:o + g * d].reshape(g, d)); bgs.append(bg[o:o + g * d].reshape(g, d)); o += g * d
        return fgs, bgs, tot[:len(b)]

    def last_ms(self):
        ms = (C.c_float * 3)(); nf = C.c_int64(); nl = C.c_int64()
        self._lib.bod_pdq_last_ms(self._h, ms, C.byref(nf), C.byref(nl))
        return dict(roi=ms[0], tables=ms[1], sums_or_maps=ms[2], table_floats=nf.value, launches=nl.value)

    def bvn_cdf(self, h, k, r) -> np.ndarray:
        h, k, r = (np.ascontiguousarray(np.broadcast_to(np.asarray(a, np.float64), np.broadcast(h, k, r).shape).ravel())
                   for a in (h, k, r))
        out = np.empty_like(h)
        self._check(self._lib.bod_pdq_bvn_cdf(self._h, len(h), h.ctypes.data, k.ctypes.data, r.ctypes.data, out.ctypes.data))
        return out


def _calc_spatial_qual(fg_loss_sum, bg_loss_sum, num_fg_pixels_vec):
    """pdq.py:233-256"""
    spatial_quality = np.exp((fg_loss_sum + bg_loss_sum) / num_fg_pixels_vec)
    spatial_quality[np.isclose(spatial_quality, 0)] = 0
    spatial_quality[np.isclose(spatial_quality, 1)] = 1
    return spatial_quality


def _gmean2(a, b):
    """scipy.stats.gmean of two stacked arrays (pdq.py:273-281): exp(mean(log))."""
    with np.errstate(divide='ignore'):
        return np.exp((np.log(np.asarray(a, np.float64)) + np.log(np.asarray(b, np.float64))) / 2)


def _qual_img(gts, dets, num_fg, fg, bg, bg_total):
    """pdq.py:328-446 (_calc_qual_img) with the loss sums already computed; also returns the cost tables of
    pdq.py:283-325 (_gen_cost_tables)."""
    from scipy.optimize import linear_sum_assignment

    G, D = len(gts), len(dets)
    if G == 0 or D == 0:                                                                    # :351-368
        fn = sum(1 for g, n in zip(gts, num_fg) if _is_gt_included(g, n))
        return {'overall': 0.0, 'spatial': 0.0, 'label': 0.0, 'TP': 0, 'FP': D, 'FN': fn}, None
    n_pairs = max(G, D)
    tables = {k: np.ones((n_pairs, n_pairs), np.float32) for k in ('overall', 'spatial', 'label')}
    label_prob = np.stack([np.asarray(d.class_list) for d in dets], axis=0)                 # d x c, :190
    gt_labels = np.array([g.class_label for g in gts], dtype=int)
    label_qual = label_prob[:, gt_labels].T.astype(np.float32)                              # :268-270
    spatial_qual = _calc_spatial_qual(fg, bg, np.asarray(num_fg, np.int64).reshape(G, 1))
    tables['overall'][:G, :D] -= _gmean2(label_qual, spatial_qual)                          # :315-316
    tables['spatial'][:G, :D] -= spatial_qual
    tables['label'][:G, :D] -= label_qual
    row_idxs, col_idxs = linear_sum_assignment(tables['overall'])                           # :377
    overall_q, spatial_q, label_q = (1 - tables[k] for k in ('overall', 'spatial', 'label'))
    # :390-406, vectorised over the assignment: a pair with positive quality is a TP if its object counts (else its
    # quality is zeroed); a zero-quality pair is an FN if its row is a counted object and an FP if its column is a detection
    included = np.zeros(n_pairs, bool)
    included[:G] = [_is_gt_included(g, n) for g, n in zip(gts, num_fg)]
    pos = overall_q[row_idxs, col_idxs] > 0
    inc = included[row_idxs]
    tp = int(np.count_nonzero(pos & inc))
    overall_q[row_idxs[pos & ~inc], col_idxs[pos & ~inc]] = 0.0
    fn = int(np.count_nonzero(~pos & inc))
    fp_cols = col_idxs[~pos & (col_idxs < D)]
    fp = int(fp_cols.size)
    tot_tp_overall = np.sum(overall_q[row_idxs, col_idxs])
    spatial_q[overall_q == 0] = 0.0
    label_q[overall_q == 0] = 0.0
    tot_tp_spatial = np.sum(spatial_q[row_idxs, col_idxs])
    tot_tp_label = np.sum(label_q[row_idxs, col_idxs])
    fp_cols = np.asarray(fp_cols, dtype=np.int64)
    fp_label = 1.0 - label_prob[fp_cols].max(axis=1) if fp_cols.size else np.zeros(0)       # :417-419
    if fp_label.size:                                                                       # :421-432
        boxes = np.stack([np.asarray(d.box) for d in dets])[fp_cols]
        area = (boxes[:, 3] - boxes[:, 1]) * (boxes[:, 2] - boxes[:, 0])                     # _compute_bb_area, :449-452
        with np.errstate(divide='ignore', invalid='ignore'):
            fp_spatial = np.exp(np.asarray(bg_total, np.float64)[fp_cols].astype(np.float32) / area)
        tot_fp_spatial = np.sum(fp_spatial)
        tot_fp_overall = np.sum(_gmean2(fp_spatial, fp_label))
    else:
        tot_fp_spatial = 0.0
        tot_fp_overall = 0.0
    res = {'overall': tot_tp_overall + tot_fp_overall, 'spatial': tot_tp_spatial + tot_fp_spatial,
           'label': tot_tp_label + np.sum(fp_label), 'TP': tp, 'FP': fp, 'FN': fn}
    return res, tables


class PDQ:
    """Drop-in for ``offline_eval.pdq.PDQ`` (same methods).  ``score`` evaluates ``images_per_call`` images per
    GPU call instead of one image per pool worker."""

    def __init__(self, img_size=(720, 1280), device: int = 0, images_per_call: int = 64):
        self._engine = PdqEngine(img_size, device)
        self._images_per_call = int(images_per_call)
        self.reset()

    def reset(self):                                                                        # pdq.py:47-57
        self._tot_overall_quality = 0.0
        self._tot_spatial_quality = 0.0
        self._tot_label_quality = 0.0
        self._tot_TP = 0
        self._tot_FP = 0
        self._tot_FN = 0

    def _accumulate(self, results):
        self._tot_overall_quality += results['overall']
        self._tot_spatial_quality += results['spatial']
        self._tot_label_quality += results['label']
        self._tot_TP += results['TP']
        self._tot_FP += results['FP']
        self._tot_FN += results['FN']

    def evaluate_images(self, matches):
        """``[_calc_qual_img(gts, dets) for gts, dets in matches]`` and the cost tables of every image."""
        img_size = self._engine.img_size
        do, go, boxes, covs, gtb, nfg = [0], [0], [], [], [], []
        for gts, dets in matches:
            # the reference never builds heat maps for an image without objects or without detections (pdq.py:351)
            use = len(gts) > 0 and len(dets) > 0
            for d in (dets if use else []):
                boxes.append(np.asarray(d.box).reshape(4)); covs.append(np.asarray(d.covs, np.float64).reshape(8))
            for g in (gts if use else []):
                gtb.append(np.asarray(g.bounding_box).reshape(4))
            nfg.append([_num_pixels(g, img_size) for g in gts])
            do.append(len(boxes)); go.append(len(gtb))
        fgs, bgs, tot = self._engine.losses(do, np.array(boxes).reshape(-1, 4), np.array(covs).reshape(-1, 8), go,
                                            np.array(gtb).reshape(-1, 4))
        out = []
        for i, (gts, dets) in enumerate(matches):
            out.append(_qual_img(gts, dets, nfg[i], fgs[i], bgs[i], tot[do[i]:do[i + 1]]))
        return out

    def add_img_eval(self, gt_instances, det_instances):                                    # pdq.py:25-39
        self._accumulate(self.evaluate_images([(gt_instances, det_instances)])[0][0])

    def score(self, matches):                                                               # pdq.py:59-88
        self.reset()
        matches = list(matches)
        for i in range(0, len(matches), self._images_per_call):
            for res, _ in self.evaluate_images(matches[i:i + self._images_per_call]):
                self._accumulate(res)
        return self.get_pdq_score()

    def get_pdq_score(self):                                                                # pdq.py:41-45
        return self._tot_overall_quality / (self._tot_TP + self._tot_FP + self._tot_FN)

    def get_avg_spatial_score(self):                                                        # pdq.py:90-100
        n = self._tot_TP + self._tot_FP
        return self._tot_spatial_quality / float(n) if n > 0.0 else 0.0

    def get_avg_label_score(self):                                                          # pdq.py:102-111
        n = self._tot_TP + self._tot_FP
        return self._tot_label_quality / float(n) if n > 0.0 else 0.0

    def get_avg_overall_quality_score(self):                                                # pdq.py:113-124
        n = self._tot_TP + self._tot_FP
        return self._tot_overall_quality / float(n) if n > 0.0 else 0.0

    def get_assignment_counts(self):                                                        # pdq.py:126-131
        return self._tot_TP, self._tot_FP, self._tot_FN
