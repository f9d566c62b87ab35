#!/usr/bin/env python
"""Mint tests/golden/pdq_*.npz (AUTHORING container only, needs /root/reference and scipy).

Executes the reference's PDQ code verbatim — no shim of its arithmetic:

    pdq_data_holders.PBoxDetInst(...).calc_heatmap(img_size)      # pdq_data_holders.py:92-117 (find_roi, gen_single_heatmap)
    pdq._gen_cost_tables(gt_instances, det_instances)             # pdq.py:283-325 (fg/bg loss, spatial/label/overall quality)
    pdq._calc_qual_img(gt_instances, det_instances)               # pdq.py:328-446

with ground truth and detections built the way bdd/compute_pdq.py:93-124 builds them (box-shaped masks, int32-truncated
corner means, corner covariances = 2x2 blocks of T·Σ·Tᵀ·70).  The only environment patch: `np.int`, removed from numpy >= 1.24,
is restored as the builtin it aliased (`np.bool` exists again in numpy 2).

    python tests/golden/make_pdq_golden.py
"""
import importlib
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)

import tf_numpy_shim as shim  # noqa: E402  (box_utils.py imports tensorflow at module level; vuhw_to_vuvu_np is pure numpy)

np.int = int      # noqa: the alias the reference still uses (pdq_data_holders.py:73-74, pdq.py:163-167)

TRANSFORM = np.array([[0, 1, 0, -0.5], [1, 0, -0.5, 0], [0, 1, 0, 0.5], [1, 0, 0.5, 0]])     # compute_pdq.py:98-101

# name: (H, W, seed, n_gt, n_det, covariance scale range of the vuhw posterior (before x70), detections hugging the borders)
CASES = {
    "pdq_small":   (96, 160, 1, 4, 6, (0.02, 0.3), False),
    "pdq_borders": (80, 120, 2, 3, 6, (0.05, 1.0), True),
    "pdq_tight":   (64, 96, 3, 2, 4, (0.001, 0.01), False),
    "pdq_wide":    (120, 200, 4, 5, 5, (0.5, 3.0), True),
    "pdq_full":    (720, 1280, 5, 3, 4, (0.05, 0.6), False),      # the size bdd/compute_pdq.py evaluates
}


def make_case(H, W, seed, G, D, cov_range, borders):
    rng = np.random.default_rng(seed)
    gts = []
    for _ in range(G):
        h, w = rng.uniform(14, H * 0.5), rng.uniform(14, W * 0.5)
        y, x = rng.uniform(0, H - h), rng.uniform(0, W - w)
        gts.append([x, y, x + w, y + h])
    gt_boxes = np.array(gts).astype(np.int32)                                  # compute_pdq.py:109 box_inds
    gt_labels = rng.integers(0, 7, G)
    means = []
    for d in range(D):                                                         # detections: jittered GT or random, vuhw
        if d < G:
            x1, y1, x2, y2 = gts[d] + rng.normal(0, 2.0, 4)
        else:
            h, w = rng.uniform(10, H * 0.6), rng.uniform(10, W * 0.6)
            y1, x1 = rng.uniform(0, H - h), rng.uniform(0, W - w)
            x2, y2 = x1 + w, y1 + h
        if borders and d % 2 == 0:
            x1, y1 = rng.uniform(0, 3), rng.uniform(0, 3)
        if borders and d % 3 == 0:
            x2, y2 = W - 1 - rng.uniform(0, 3), H - 1 - rng.uniform(0, 3)
        means.append([(y1 + y2) / 2, (x1 + x2) / 2, y2 - y1, x2 - x1])
    means = np.array(means, np.float32)
    covs = []
    for _ in range(D):
        a = rng.normal(size=(4, 4))
        covs.append((a @ a.T / 4 + np.eye(4)) * rng.uniform(*cov_range))
    covs = np.array(covs, np.float32)                                          # what the .npy files hold (already x70, :361)
    cat = rng.dirichlet(np.ones(8) * 0.3, D).astype(np.float32)
    cat[np.arange(D), rng.integers(0, 7, D)] += 1.0
    cat /= cat.sum(1, keepdims=True)
    for d in range(min(G, D)):                                                 # make a few labels agree with their object
        cat[d] = 0.02; cat[d, gt_labels[d]] = 0.86
    return gt_boxes, gt_labels, means, covs, cat


def main():
    shim.load_reference()
    dh = importlib.import_module("src.retina_net.offline_eval.pdq_data_holders")
    pdq = importlib.import_module("src.retina_net.offline_eval.pdq")
    bu = importlib.import_module("src.retina_net.anchor_generator.box_utils")
    for name, (H, W, seed, G, D, cov_range, borders) in CASES.items():
        gt_boxes, gt_labels, means, covs4, cat = make_case(H, W, seed, G, D, cov_range, borders)
        # --- compute_pdq.py:93-124, with [720, 1280] -> [H, W] and without the score gate (every detection is kept)
        cov_t = np.matmul(np.matmul(TRANSFORM, covs4), TRANSFORM.T) * 70
        boxes_vuvu = bu.vuhw_to_vuvu_np(means)
        gt_instances = []
        for label, b in zip(gt_labels, gt_boxes):
            seg = np.zeros([H, W], dtype=bool)
            seg[b[1]:b[3], b[0]:b[2]] = True
            gt_instances.append(dh.GroundTruthInstance(seg, int(label), 0, 0, bounding_box=b))
        dets, boxes, covs = [], [], []
        for c, bx, cv in zip(cat, boxes_vuvu, cov_t):
            box = np.array([bx[1], bx[0], bx[3], bx[2]]).astype(np.int32)
            cp = [cv[0:2, 0:2], cv[2:4, 2:4]]
            dets.append(dh.PBoxDetInst(c, box, cp))
            boxes.append(box); covs.append(np.stack(cp))
        heatmaps = np.stack([d.calc_heatmap((H, W)) for d in dets])
        seg_mat, bg_mat, nfg, _ = pdq._vectorize_img_gts(gt_instances, (H, W))
        hm_hwd = np.ascontiguousarray(heatmaps.transpose(1, 2, 0))
        fg = pdq._calc_fg_loss(seg_mat, hm_hwd)
        bg = pdq._calc_bg_loss(bg_mat, hm_hwd)
        spatial = pdq._calc_spatial_qual(fg, bg, nfg)
        tables, _ = pdq._gen_cost_tables(gt_instances, dets)
        res = pdq._calc_qual_img(gt_instances, dets)
        rois = np.array([[dh.find_roi((H, W), [b[1], b[0]], np.flipud(np.fliplr(c[0]))),
                          dh.find_roi((H, W), [H - (b[3] + 1), W - (b[2] + 1)], np.flipud(np.fliplr(c[1])).T)]
                         for b, c in zip(boxes, covs)], np.int32)
        out = dict(img_size=np.array([H, W], np.int32), gt_boxes=gt_boxes, gt_labels=gt_labels.astype(np.int32),
                   means_vuhw=means, covs_vuhw=covs4, cat_param=cat,
                   boxes=np.array(boxes, np.int32), covs=np.array(covs, np.float64), rois=rois,
                   heatmaps=heatmaps.astype(np.float32), fg_loss=np.asarray(fg, np.float64), bg_loss=np.asarray(bg, np.float64),
                   num_fg=nfg.astype(np.int64), spatial=np.asarray(spatial, np.float64),
                   cost_overall=tables['overall'], cost_spatial=tables['spatial'], cost_label=tables['label'],
                   res_overall=np.float64(res['overall']), res_spatial=np.float64(res['spatial']),
                   res_label=np.float64(res['label']), res_counts=np.array([res['TP'], res['FP'], res['FN']], np.int32))
        np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
        print(f"{name:12s} {H}x{W} G={G} D={D} nonzero px/det={np.count_nonzero(heatmaps) / D:.0f} "
              f"spatial max={spatial.max():.3f} TP/FP/FN={res['TP']}/{res['FP']}/{res['FN']} overall={res['overall']:.4f}")


if __name__ == "__main__":
    main()
